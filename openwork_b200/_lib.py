"""ctypes binding of ``libtvg_b200.so`` (C ABI: ``include/tvg_b200.h``).

There is no fallback: if the library is missing or no B200 is visible the import / context creation
raises.  PyTorch is used only for device memory and streams; signatures carry plain pointers.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "libtvg_b200.so")

c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_u8p = C.POINTER(C.c_uint8)
c_f64p = C.POINTER(C.c_double)
c_f32p = C.POINTER(C.c_float)
c_vp = C.c_void_p


class RansacIO(C.Structure):
    """tvg_ransac_io"""
    _fields_ = [
        ("src", c_vp), ("dst", c_vp), ("corr32", c_vp), ("m_count", c_vp),
        ("batch", C.c_int), ("n_max", C.c_int), ("m", C.c_int), ("H", C.c_int),
        ("samples", c_vp), ("words", c_vp), ("sample_stride", C.c_longlong), ("thr", C.c_double),
        ("qidx", c_vp), ("tidx", c_vp), ("pair_q", c_vp), ("pair_t", c_vp), ("kps_q", c_vp), ("kps_t", c_vp),
        ("E", c_vp), ("mask", c_vp), ("n_inl", c_vp), ("best", c_vp), ("E_best", c_vp), ("counts", c_vp),
        ("sse32", c_vp), ("models", c_vp), ("Rt", c_vp), ("status", c_vp), ("inl_idx", c_vp),
        ("samples_out", c_vp), ("pix_last", c_vp), ("pix_curr", c_vp), ("band", c_vp), ("n_cand", c_vp), ("best_sse", c_vp),
    ]


class PairsIO(C.Structure):
    """tvg_pairs_io"""
    _fields_ = [
        ("des", c_vp), ("kps", c_vp), ("frame_count", c_vp), ("pair_q", c_vp), ("pair_t", c_vp),
        ("batch", C.c_int), ("n_max", C.c_int), ("H", C.c_int),
        ("seq_off", c_vp), ("nseq", C.c_int), ("init_pose", c_vp), ("words", c_vp), ("samples", c_vp),
        ("sample_stride", C.c_longlong), ("kinv", C.c_double * 9), ("ratio", C.c_double), ("thr", C.c_double),
        ("idx", c_vp), ("dist", c_vp), ("qidx", c_vp), ("tidx", c_vp), ("m_count", c_vp), ("E", c_vp),
        ("E_best", c_vp), ("best", c_vp), ("counts", c_vp), ("mask", c_vp), ("n_inl", c_vp), ("inl_idx", c_vp),
        ("Rt", c_vp), ("poses", c_vp), ("points4d", c_vp), ("good", c_vp), ("status", c_vp),
        ("pose_mode", C.c_int), ("pose_votes", c_vp),
    ]


class HugePairIO(C.Structure):
    """tvg_huge_pair_io"""
    _fields_ = [
        ("curr_des", c_vp), ("last_des", c_vp), ("curr_kps", c_vp), ("last_kps", c_vp),
        ("nq", C.c_int), ("nt", C.c_int), ("H", C.c_int), ("words", c_vp), ("samples", c_vp),
        ("kinv", C.c_double * 9), ("ratio", C.c_double), ("thr", C.c_double),
        ("idx", c_vp), ("dist", c_vp), ("qidx", c_vp), ("tidx", c_vp), ("m_count", c_vp), ("E", c_vp), ("E_best", c_vp),
        ("best", c_vp), ("best_sse", c_vp), ("counts", c_vp), ("n_cand", c_vp), ("mask", c_vp), ("n_inl", c_vp),
        ("inl_idx", c_vp), ("Rt", c_vp), ("poses", c_vp), ("points4d", c_vp), ("good", c_vp), ("status", c_vp),
    ]


COMM_ID_BYTES = 128

# name -> (restype, argtypes); also the list tests check against include/tvg_b200.h
SIGNATURES = {
    "tvg_ctx_create": (C.c_int, [C.POINTER(c_vp), C.c_int]),
    "tvg_ctx_destroy": (None, [c_vp]),
    "tvg_last_error": (C.c_char_p, []),
    "tvg_abi_version": (C.c_int, []),
    "tvg_launch_count": (C.c_longlong, [c_vp]),
    "tvg_set_matcher_config": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int]),
    "tvg_knn2_hamming256": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp,
                                      c_vp, C.c_double, c_vp, c_vp, c_vp, c_vp]),
    "tvg_ratio_compact": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                    c_f64p, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tvg_ransac_essential": (C.c_int, [c_vp, C.POINTER(RansacIO), c_vp]),
    "tvg_extract_rt": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_vp]),
    "tvg_pose_chain": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp]),
    "tvg_invert_poses": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_vp]),
    "tvg_triangulate": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp, C.c_longlong, c_vp,
                                  C.c_longlong, c_f64p, c_vp, c_vp, c_vp]),
    "tvg_point_quality": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp, C.c_longlong, c_vp,
                                    C.c_longlong, c_f64p, C.c_double, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tvg_map_accumulate": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_longlong, C.c_int, c_vp,
                                     c_vp, c_vp, c_vp]),
    "tvg_extract_points": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                     c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tvg_cheirality": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp,
                                 c_vp, c_vp]),
    "tvg_bundle_adjust_pairs": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp,
                                          C.c_longlong, c_vp, C.c_longlong, c_vp, C.c_int, C.c_double, c_vp, c_vp, c_vp]),
    "tvg_bundle_adjust_window": (C.c_int, [c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, C.c_int, c_vp, c_vp,
                                           c_vp, C.c_int, C.c_double, c_vp, c_vp, c_vp]),
    "tvg_pairs_device": (C.c_int, [c_vp, C.POINTER(PairsIO), c_vp]),
    "tvg_set_host_chunk": (C.c_int, [c_vp, C.c_int]),
    "tvg_set_matcher_l2_budget": (C.c_int, [c_vp, C.c_longlong]),
    "tvg_set_score_config": (C.c_int, [c_vp, C.c_int, C.c_int]),
    "tvg_set_profile": (C.c_int, [c_vp, C.c_int]),
    "tvg_get_stage_ms": (C.c_int, [c_vp, c_f32p, C.c_int]),
    "tvg_get_exchange_ms": (C.c_int, [c_vp, c_f32p]),
    "tvg_pairs_host": (C.c_int, [c_vp, C.POINTER(PairsIO), C.c_int, C.POINTER(C.c_longlong),
                                 C.POINTER(C.c_longlong), c_vp]),
    "tvg_pairs_host_submit": (C.c_int, [c_vp, C.POINTER(PairsIO), C.c_int, C.POINTER(C.c_longlong),
                                        C.POINTER(C.c_longlong), c_vp]),
    "tvg_pairs_host_wait": (C.c_int, [c_vp]),
    "tvg_measure_fp8_peak": (C.c_int, [c_vp, C.c_int, c_f64p, c_f64p]),
    "tvg_comm_unique_id": (C.c_int, [c_vp]),
    "tvg_comm_init": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int]),
    "tvg_comm_destroy": (C.c_int, [c_vp]),
    "tvg_comm_info": (C.c_int, [c_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "tvg_shard_range": (None, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "tvg_huge_pair_rows": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "tvg_huge_pair": (C.c_int, [c_vp, C.POINTER(HugePairIO), c_vp]),
}

_lib = None


class TvgError(RuntimeError):
    pass


def load():
    """dlopen the library (no GPU needed for this step). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    # Build (or rebuild after a source change) with nvcc; this compiles the extension, it is not a fallback.
    try:
        from . import build as _build
        _build.build()
    except Exception as e:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                "openwork_b200: %s is missing and could not be built (%s) -- run `python -m openwork_b200.build`. "
                "There is no CPU fallback." % (LIB_PATH, e))
    # TVG_B200_LIB: a differently-configured build of the SAME library (kernel-knob sweeps, tools/sweep_tc2_knobs.sh)
    lib = C.CDLL(os.environ.get("TVG_B200_LIB") or LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().tvg_last_error()
        raise TvgError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def comm_unique_id():
    """128 bytes identifying a new communicator (rank 0 creates them and hands them to every rank)."""
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    check(load().tvg_comm_unique_id(buf), "tvg_comm_unique_id")
    return bytes(buf)


class Context:
    """One tvg_ctx (device + scratch arena). Not thread-safe; use one per stream."""

    def __init__(self, device=0):
        self.lib = load()
        h = c_vp()
        check(self.lib.tvg_ctx_create(C.byref(h), int(device)), "tvg_ctx_create")
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.tvg_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(self.lib.tvg_launch_count(self.handle))

    STAGES = ("knn2", "compact", "fit", "score", "reduce", "rescore", "refit_solve", "pose", "triangulate")

    def set_profile(self, on=True):
        check(self.lib.tvg_set_profile(self.handle, 1 if on else 0), "tvg_set_profile")

    def stage_ms(self):
        """Durations (ms) of the 9 stages of the last profiled tvg_pairs_device call."""
        buf = (C.c_float * 9)()
        check(self.lib.tvg_get_stage_ms(self.handle, buf, 9), "tvg_get_stage_ms")
        return dict(zip(self.STAGES, [float(v) for v in buf]))

    def exchange_ms(self):
        """(exchange 1, exchange 2) durations in ms of the last profiled tvg_huge_pair call on this rank; each includes
        the wait for the slowest rank."""
        buf = (C.c_float * 2)()
        check(self.lib.tvg_get_exchange_ms(self.handle, buf), "tvg_get_exchange_ms")
        return float(buf[0]), float(buf[1])

    def set_matcher_l2_budget(self, nbytes=0):
        check(self.lib.tvg_set_matcher_l2_budget(self.handle, int(nbytes)), "tvg_set_matcher_l2_budget")

    def measure_fp8_peak(self, iters=20000):
        """Measured e4m3 tcgen05.mma peak of this device: (TFLOP/s, ms of the timed launch)."""
        t, ms = C.c_double(0), C.c_double(0)
        check(self.lib.tvg_measure_fp8_peak(self.handle, int(iters), C.byref(t), C.byref(ms)), "tvg_measure_fp8_peak")
        return t.value, ms.value

    def comm_init(self, id_bytes, rank, world):
        """Bind an NCCL communicator to this context (collective over all ranks; id from :func:`comm_unique_id`)."""
        buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(bytes(id_bytes))
        check(self.lib.tvg_comm_init(self.handle, buf, int(rank), int(world)), "tvg_comm_init")

    def comm_destroy(self):
        check(self.lib.tvg_comm_destroy(self.handle), "tvg_comm_destroy")

    def comm_info(self):
        r, w = C.c_int(0), C.c_int(1)
        check(self.lib.tvg_comm_info(self.handle, C.byref(r), C.byref(w)), "tvg_comm_info")
        return r.value, w.value

    def set_score_config(self, hyp_per_thread=0, ctas_per_sm=0):
        """Scoring kernel tuning: hypotheses per thread (0 = default, 2, 4), resident CTAs per SM (0 = automatic)."""
        check(self.lib.tvg_set_score_config(self.handle, int(hyp_per_thread), int(ctas_per_sm)), "tvg_set_score_config")

    def set_host_chunk(self, pairs=256):
        check(self.lib.tvg_set_host_chunk(self.handle, int(pairs)), "tvg_set_host_chunk")

    def set_matcher_config(self, r=4, variant=-1, ctas_per_sm=6):
        """variant -1 = automatic (tensor cores for large calls), 16 = tensor cores, 0..11 = integer-pipe variants."""
        check(self.lib.tvg_set_matcher_config(self.handle, r, variant, ctas_per_sm), "tvg_set_matcher_config")
