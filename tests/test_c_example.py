"""The C ABI is usable from plain C: examples/knn_demo.c compiles and links against libtvg_b200.so with gcc
(CPU check), and on a B200 its output equals the numpy oracle's on the same LCG-generated descriptors."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = "/usr/local/cuda"


def _build(lib_built, out_dir):
    exe = os.path.join(str(out_dir), "knn_demo")
    libdir = os.path.dirname(lib_built)
    cudart = os.path.join(CUDA, "lib64")
    cmd = ["gcc", "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "examples", "knn_demo.c"), "-o", exe, "-L", libdir, "-ltvg_b200", "-L", cudart,
           "-lcudart", "-Wl,-rpath," + libdir, "-Wl,-rpath," + cudart]
    subprocess.run(cmd, check=True)
    return exe


def test_c_example_compiles_and_links(lib_built, tmp_path):
    exe = _build(lib_built, tmp_path)
    out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libtvg_b200.so" in out


def _lcg_bytes(n, seed):
    out = np.empty(n, dtype=np.uint8)
    s = np.uint32(seed)
    with np.errstate(over="ignore"):
        for i in range(n):
            s = np.uint32(s * np.uint32(1664525) + np.uint32(1013904223))
            out[i] = s >> np.uint32(24)
    return out, int(s)


@pytest.mark.gpu
def test_c_example_matches_oracle(lib_built, tmp_path):
    from oracle import two_view as tv
    exe = _build(lib_built, tmp_path)
    nq, nt, seed = 300, 500, 7
    r = subprocess.run([exe, str(nq), str(nt), str(seed)], capture_output=True, text=True, check=True)
    got = r.stdout.split()
    qb, s = _lcg_bytes(nq * 32, seed)
    tb, _ = _lcg_bytes(nt * 32, s)
    q, t = qb.reshape(nq, 32), tb.reshape(nt, 32).copy()
    for i in range(nt // 4):
        t[nt - 1 - i] = t[i]
    idx, dist = tv.knn2_cv2(q, t)
    ci = cd = 0
    for i in range(nq):
        ci = (ci * 1000003 + (int(idx[i, 0]) & 0xFFFFFFFF) * 65537 + (int(idx[i, 1]) & 0xFFFFFFFF)) % (1 << 64)
        cd = (cd * 1000003 + (int(dist[i, 0]) & 0xFFFFFFFF) * 65537 + (int(dist[i, 1]) & 0xFFFFFFFF)) % (1 << 64)
    kept = len(tv.ratio_filter(idx, dist)[0])
    assert [int(got[0]), int(got[1])] == [nq, nt]
    assert int(got[2]) == ci and int(got[3]) == cd and int(got[4]) == kept and int(got[5]) == 1
