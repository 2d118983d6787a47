"""ctypes binding of libclc_b200.so (the C ABI declared in include/clc_b200.h).

There is no fallback: if the shared library is missing the import of any op fails loudly
with the build instruction.  Every call goes through `call()`, which turns a negative
status into a RuntimeError carrying clc_strerror / clc_last_cuda_error.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLC_B200_LIB") or os.path.join(_HERE, "libclc_b200.so")

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f = C.c_float
_sz = C.c_size_t


class PatchView(C.Structure):
    """struct clc_patch_view (include/clc_b200.h)."""
    _fields_ = [("q", _p), ("q_sn", _i64), ("q_spy", _i64), ("q_spx", _i64), ("q_sc", _i64),
                ("q_sy", _i64), ("npx", _i32), ("q_repeat", _i32)]


_PTR5 = _p * 5
_PTR4 = _p * 4

# name -> (restype, argtypes); mirrors include/clc_b200.h one to one.
PROTOTYPES = {
    "clc_version": (C.c_int, []),
    "clc_strerror": (C.c_char_p, [C.c_int]),
    "clc_last_cuda_error": (C.c_char_p, []),
    "clc_kernel_launch_count": (C.c_uint64, []),
    "clc_trace_start": (C.c_int, [_p]),
    "clc_trace_mark": (C.c_int, []),
    "clc_trace_stop": (C.c_int, []),
    "clc_trace_count": (C.c_int, []),
    "clc_trace_get": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]),
    "clc_gc_fwd": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _p,
                             _i64, _i64, _f, _f, _p]),
    "clc_gc_bwd": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _f, _p, _i64,
                             _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _f, _f, _p]),
    "clc_gc_fwd_rng": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _p, C.c_uint64, _p, _i64, _p, _i64, _p, _i64, _p,
                                 _i64, _i64, _f, _f, _p]),
    "clc_gc_bwd_rng": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _p, C.c_uint64, _p, _i64, _p, _i64, _f, _p, _i64,
                                 _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _f, _f, _p]),
    "clc_rng_advance": (C.c_int, [_p, C.c_uint64, _p]),
    "clc_eb_fwd_rng": (C.c_int, [_p, _p, C.c_uint64, _PTR5, _PTR5, _PTR4, _p, _p, _p, _p, _p, _i64, _i64, _i64, _f, _p]),
    "clc_eb_bwd_rng": (C.c_int, [_p, _p, C.c_uint64, _PTR5, _PTR5, _PTR4, _p, _p, _p, _f, _p, _p, _p, _p, _p,
                                 _i64, _i64, _i64, _f, _p]),
    "clc_bpp_finalize": (C.c_int, [_p, _i32, C.c_double, _p, _p, C.c_uint64, _p]),
    "clc_zero": (C.c_int, [_p, _sz, _p]),
    "clc_lrp_add_fwd": (C.c_int, [_p, _i64, _p, _i64, _i64, _i64, _p]),
    "clc_lrp_add_bwd": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i64, _i64, _p]),
    "clc_gc_symbols_indexes": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _p, C.c_int, _p, _i64, _p, _i64,
                                         _i64, _i64, _f, _p]),
    "clc_eb_fwd": (C.c_int, [_p, _p, _PTR5, _PTR5, _PTR4, _p, _p, _p, _p, _p, _i64, _i64, _i64, _f, _p]),
    "clc_eb_bwd": (C.c_int, [_p, _p, _PTR5, _PTR5, _PTR4, _p, _p, _p, _f, _p, _p, _p, _p, _p,
                             _i64, _i64, _i64, _f, _p]),
    "clc_log2_sum_fwd": (C.c_int, [_p, _i64, _p, _p]),
    "clc_log2_sum_bwd": (C.c_int, [_p, _f, _p, _p, _i64, _p]),
    "clc_pearson_corr": (C.c_int, [C.POINTER(PatchView), _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _i32,
                                   _p, _sz, _p]),
    "clc_pearson_corr_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32, _i32, _i32]),
    "clc_topk_rows": (C.c_int, [_p, _i64, _i64, _i32, _p, _p, _p]),
    "clc_gaussian_mask": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p]),
    "clc_match_topk_tc": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p,
                                    _f, _p, _p, _p, _sz, _p]),
    "clc_match_clm_fwd": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p,
                                    _f, _p, _p, _p, _i64, _i64, _p, _p, _sz, _p]),
    "clc_match_topk_tc_ref_cl": (_p, [_p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "clc_match_topk_tc_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "clc_gather_blend_fwd": (C.c_int, [_p, _p, _p, _f, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                       _i32, _p]),
    "clc_gather_blend_bwd": (C.c_int, [_p, _p, _p, _f, _p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _i32,
                                       _i32, _i32, _p]),
    "clc_pearson_topk_bwd": (C.c_int, [C.POINTER(PatchView), _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32,
                                       _i32, _i32, _i32, _i32, _p]),
    "clc_match_bwd": (C.c_int, [C.POINTER(PatchView), _p, _p, _p, _p, _p, _f, _p, _p, _p, _p, _i64, _i32, _i32, _i32,
                                _i32, _i32, _i32, _i32, _i32, _p, _sz, _p]),
    "clc_match_clm_bwd": (C.c_int, [C.POINTER(PatchView), _p, _p, _p, _p, _f, _p, _p, _i64, _i64, _p, _p, _p, _p, _p,
                                    _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _sz, _p]),
    "clc_match_bwd_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "clc_match_bwd_zero_workspace": (C.c_int, [_p, _sz, _i64, _i32, _i32, _i32, _p]),
    "clc_pmf_to_quantized_cdf": (C.c_int, [_p, _i32, _i32, _p]),
    "clc_rans_encode": (C.c_int, [_p, _p, _i64, _p, _i32, _i32, _p, _p, _p, _sz, C.POINTER(_sz)]),
    "clc_rans_encode_capacity": (_sz, [_i64]),
    "clc_rans_decode": (C.c_int, [_p, _sz, _p, _p, _i64, _p, _i32, _i32, _p, _p, _p]),
    "clc_peer_alloc": (C.c_int, [_sz, C.POINTER(_p)]),
    "clc_peer_free": (C.c_int, [_p]),
    "clc_peer_export": (C.c_int, [_p, _p]),
    "clc_peer_open": (C.c_int, [_p, C.POINTER(_p)]),
    "clc_peer_close": (C.c_int, [_p]),
    "clc_peer_allreduce_bytes": (_sz, [_i32, _i32, _i32]),
    "clc_peer_allreduce": (C.c_int, [_p, _i32, _i32, _p, _i32, _p, _i32, _f, _p, _p]),
    "clc_clm_fuse_fwd": (C.c_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _p, _i32, _i64, _i32, _i64, _p]),
    "clc_clm_fuse_bwd": (C.c_int, [_p, _i64, _i64, _p, _i64, _i64, _p, _p, _p, _i32, _i64, _i32, _i64, _p]),
    "clc_clm_sim_colsum_workspace_bytes": (C.c_size_t, [_i64, _i64]),
    "clc_clm_sim_colsum": (C.c_int, [_p, _p, _i64, _i64, _i32, _i64, C.c_float, _p, _p, C.c_size_t, _p]),
    "clc_clm_weighted_concat": (C.c_int, [_p, _p, _p, _i64, _i32, _i64, _p]),
    "clc_clm_deform_fwd": (C.c_int, [_p, _p, _p, _i32, _p, _i64, _i32, _i32, _i32, _p]),
    "clc_clm_attention_sum_fwd": (C.c_int, [_p, _p, _p, _p, _i32, _i64, _i32, _i64, _p]),
    "clc_knn_neg_sqdist": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p]),
    "clc_window_attention_fwd": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _f, _p]),
}

# Bring-up entry points: only in libclc_b200_dbg.so (the -DCLC_DEBUG_ABI build), see debug_lib().
DEBUG_PROTOTYPES = {
    "clc_debug_set_stage_mask": (None, [C.c_int]),
    "clc_debug_rescore_stamps": (C.c_int, [_p]),
    "clc_debug_bwd_stamps": (C.c_int, [_p]),
    "clc_debug_match_tc_xy": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p,
                                        _p, _sz, _p]),
    "clc_debug_match_tc_timing": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p,
                                            _p, _sz, _p]),
}
DEBUG_LIB_PATH = os.path.join(_HERE, "libclc_b200_dbg.so")

_lib = None
_dbg_lib = None
_launches = 0  # number of C-ABI kernel-enqueueing calls made by this process (bench.py reads it)


def lib():
    """Load (once) and return the ctypes handle; fail loudly if the extension is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"clc_b200: {LIB_PATH} is missing -- the CUDA extension is not built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback.")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)  # AttributeError if the library does not export the symbol
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def debug_lib():
    """Bring-up build of the same sources (clc_debug_* hooks, stage masks, in-kernel stamps).  For tests/ and
    scripts/ only -- nothing under clc_b200/ calls this."""
    global _dbg_lib
    if _dbg_lib is None:
        if not os.path.exists(DEBUG_LIB_PATH):
            raise RuntimeError(f"clc_b200: {DEBUG_LIB_PATH} is missing -- run __graft_entry__.build()")
        h = C.CDLL(DEBUG_LIB_PATH)
        for table in (PROTOTYPES, DEBUG_PROTOTYPES):
            for name, (res, args) in table.items():
                fn = getattr(h, name)
                fn.restype = res
                fn.argtypes = args
        _dbg_lib = h
    return _dbg_lib


TRACE = None  # bench.py sets this to a list to collect (name, start_event, end_event) per call


def call(name, *args):
    """Invoke a status-returning entry point; raise RuntimeError on a negative status."""
    global _launches
    h = lib()
    if TRACE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(h, name)(*args)
        e1.record()
        TRACE.append((name, e0, e1))
    else:
        rc = getattr(h, name)(*args)
    _launches += 1
    if rc != 0:
        msg = h.clc_strerror(rc).decode()
        cuda = h.clc_last_cuda_error().decode()
        raise RuntimeError(f"{name} failed: {msg}" + (f" [{cuda}]" if cuda and rc == -4 else ""))
    return rc


def kernel_trace(fn, stream):
    """Run fn() with per-kernel tracing on `stream`; returns [(label, ms), ...] in launch order."""
    h = lib()
    if h.clc_trace_start(stream) != 0:
        raise RuntimeError("clc_trace_start failed")
    try:
        fn()
    finally:
        h.clc_trace_stop()
    out = []
    name, ms = C.c_char_p(), C.c_float()
    for i in range(h.clc_trace_count()):
        if h.clc_trace_get(i, C.byref(name), C.byref(ms)) == 0 and name.value != b"(mark)":
            out.append((name.value.decode(), ms.value))
    return out


def launches():
    """Kernels enqueued by libclc_b200.so in this process (counted inside the library)."""
    return int(lib().clc_kernel_launch_count())


def ptr(t):
    """Device pointer of a tensor (or NULL for None)."""
    return None if t is None else t.data_ptr()


def ptr_array(tensors, n):
    arr = (_p * n)()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
