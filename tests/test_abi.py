"""CPU tests of the drop-in boundary: the shared library builds for sm_100a, loads, exports every
symbol include/tvg_b200.h declares, the ctypes structs mirror the C structs, and the product has no
CPU fallback (no oracle import, loud failure without a GPU)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tvg_b200.h")


def _declared():
    src = open(HEADER).read()
    return re.findall(r"^TVG_API [\w \*]+?\b(tvg_\w+)\(", src, flags=re.M)


def test_library_builds_and_exports_every_declared_symbol(lib_built):
    from openwork_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(names) == sorted(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert lib.tvg_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", lib_built], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    assert set(names) <= exported
    assert all(e.startswith("tvg_") for e in exported), exported - set(names)   # nothing else leaks


def test_ctypes_structs_mirror_the_c_structs(lib_built, tmp_path):
    from openwork_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "tvg_b200.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(tvg_ransac_io), sizeof(tvg_pairs_io),'
                    ' offsetof(tvg_ransac_io, thr), offsetof(tvg_ransac_io, n_cand), offsetof(tvg_pairs_io, kinv),'
                    ' offsetof(tvg_pairs_io, pose_votes)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(_lib.RansacIO)
    assert vals[1] == ctypes.sizeof(_lib.PairsIO)
    assert vals[2] == _lib.RansacIO.thr.offset and vals[3] == _lib.RansacIO.n_cand.offset
    assert vals[4] == _lib.PairsIO.kinv.offset and vals[5] == _lib.PairsIO.pose_votes.offset


def test_sass_shows_the_blackwell_units(lib_built):
    """SASS evidence of what the kernels run on: the tensor-core matcher issues tcgen05.mma kind::f8f6f4 (UTCQMMA)
    with accumulators read back from tensor memory (LDTM), every staged tile arrives by a bulk async copy (UBLKCP:
    the 1-D bulk form -- the tiles are contiguous, no tensor map) signalled on mbarriers (SYNCS), the integer-pipe
    matcher is LOP3 + POPC, and the scoring kernel counts with PRMT."""
    sass = subprocess.run(["cuobjdump", "-sass", lib_built], capture_output=True, text=True).stdout
    for mnemonic in ("UTCQMMA", "LDTM", "UBLKCP", "SYNCS", "POPC", "LOP3", "PRMT", "DFMA"):
        assert mnemonic in sass, mnemonic
    assert "HMMA" not in sass and "HGMMA" not in sass          # no legacy tensor path
    assert "sm_100a" in sass or "SM100a" in sass or "EF_CUDA_SM100" in sass


def test_huge_pair_struct_layout_matches_the_header(tmp_path):
    import ctypes
    from openwork_b200 import _lib
    prog = tmp_path / "hp.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "tvg_b200.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %d\\n", sizeof(tvg_huge_pair_io), offsetof(tvg_huge_pair_io, kinv),'
                    ' offsetof(tvg_huge_pair_io, counts), offsetof(tvg_huge_pair_io, status), TVG_COMM_ID_BYTES); return 0;}\n')
    exe = tmp_path / "hp"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    vals = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(_lib.HugePairIO)
    assert vals[1] == _lib.HugePairIO.kinv.offset and vals[2] == _lib.HugePairIO.counts.offset
    assert vals[3] == _lib.HugePairIO.status.offset and vals[4] == _lib.COMM_ID_BYTES


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "openwork_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+(cv2|skimage)\b", text, flags=re.M), f
    code = "import sys; import openwork_b200, openwork_b200.slam, openwork_b200.ops, openwork_b200.dist; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules); " \
           "assert 'cv2' not in sys.modules and 'skimage' not in sys.modules"
    subprocess.run([sys.executable, "-c", code], cwd=ROOT, check=True)


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a GPU-less box")
def test_fails_loudly_without_a_gpu(lib_built):
    from openwork_b200 import _lib, ops
    import numpy as np
    with pytest.raises(_lib.TvgError):
        _lib.Context(0)
    with pytest.raises(Exception):
        ops.knn2_hamming256(np.zeros((4, 32), np.uint8), np.zeros((4, 32), np.uint8))
