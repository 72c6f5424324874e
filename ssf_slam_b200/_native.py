"""ctypes binding of the C ABI declared in ``include/ssf_b200.h`` (``libssf_b200.so``).

PyTorch is only plumbing here: it owns device memory and the stream; every kernel is ours and is called
through the C ABI with raw device pointers.  There is no fallback: a missing library, a non-CUDA tensor or
a non-sm_100 device raises.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssf_b200.so")

_P, _I, _F, _U64, _I64, _D = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_ulonglong, ctypes.c_longlong, ctypes.c_double
_CODES = {"p": _P, "i": _I, "f": _F, "Q": _U64, "q": _I64, "d": _D, "H": ctypes.c_void_p, "I": ctypes.c_uint}   # "H": host pointer (no device bookkeeping)

# name -> (argument codes, restype); mirrors include/ssf_b200.h one to one
SIGNATURES = {
    "ssf_abi_version": ("", _I),
    "ssf_last_error": ("", ctypes.c_char_p),
    "ssf_launch_count": ("", _U64),
    "ssf_require_device": ("", _I),
    "ssf_furthest_point_sample": ("piiipp", _I),
    "ssf_gather_operation": ("ppiiiipp", _I),
    "ssf_knn": ("ippiiippp", _I),
    "ssf_knn_offset": ("ipppiiippp", _I),
    "ssf_knn_blocks_workspace_floats": ("ii", _I64),
    "ssf_knn_blocks_build": ("piipp", _I),
    "ssf_knn_blocks_search": ("ipppiiippp", _I),
    "ssf_knn_warp_scan": ("ipppiiippp", _I),
    "ssf_ball_query_blocks": ("fippiiippp", _I),
    "ssf_three_nn": ("ppiiippp", _I),
    "ssf_ball_query": ("fippiiippp", _I),
    "ssf_grouping_operation": ("ppiiiiipp", _I),
    "ssf_three_interpolate": ("pppiiiipp", _I),
    "ssf_csr_workspace_ints": ("iii", _I64),
    "ssf_build_csr_i64": ("piiipp", _I),
    "ssf_build_csr_i32": ("piiipp", _I),
    "ssf_segment_softmax": ("ppiiiipp", _I),
    "ssf_segment_sum": ("ppiiiipp", _I),
    "ssf_segment_softmax_sum": ("pppiiiipp", _I),
    "ssf_linear": ("piipiipiiipiiifpifpip", _I),
    "ssf_gather_rows": ("ppiiiipp", _I),
    "ssf_transpose": ("piiipp", _I),
    "ssf_interpolate": ("ppppiiiiiiifpp", _I),
    "ssf_group_mlp_max": ("ppppppp" + "ppi" + "ppi" + "iiiiii" + "pp", _I),
    "ssf_cost_volume": ("p" * 16 + "f" + "pppp" + "iiii" + "pppp" + "p", _I),
    "ssf_cost_volume_tc": ("p" * 9 + "iiii" + "pppp" + "ip", _I),
    "ssf_cost_volume_tc_blob_bytes": ("", _I64),
    "ssf_cost_volume_tc_param_floats": ("", _I),
    "ssf_attention_mix": ("ppqippp", _I),
    "ssf_softmax_pool": ("ppiiippp", _I),
    "ssf_dense_tc": ("pp", _I),
    "ssf_dense_args_bytes": ("", _I),
    "ssf_dense_set_variant": ("i", _I),
    "ssf_dense_set_tma": ("i", _I),
    "ssf_frontend": ("ppiiippQpifpppp", _I),
    "ssf_solve_rt_f64": ("ppiippp", _I),
    "ssf_gmm_mask": ("ppiiidppp", _I),
    "ssf_dataset_select": ("pippiifippp", _I),
    "ssf_index_compose": ("pipippp", _I),
    "ssf_gather_u8": ("pipipp", _I),
    "ssf_plane_features_workspace_bytes": ("ii", _I64),
    "ssf_plane_features": ("piiiiifipppp", _I),
}

# developer library (csrc/dev/ssf_b200_dev.h -> libssf_b200_dev.so): bring-up / probe kernels kept out of the product ABI
DEV_LIB_PATH = os.path.join(_HERE, "libssf_b200_dev.so")
DEV_SIGNATURES = {
    "ssf_last_error": ("", ctypes.c_char_p),
    "ssf_tc_gemm_test": ("pppiiiipp", _I),
    "ssf_tc_mma_rate": ("iiiipp", _I),
    "ssf_dev_hilbert30": ("HiH", _I),
    "ssf_dev_fastdiv": ("HiIH", _I),
}



class DenseArgs(ctypes.Structure):
    """Mirror of ``ssf_dense_args`` (include/ssf_dense.h), field for field."""
    _fields_ = [("a_mode", _I), ("K", _I), ("rows", _I64),
                ("x1", _P), ("c1", _I), ("ld1", _I), ("x2", _P), ("c2", _I), ("ld2", _I),
                ("G", _P), ("ldG", _I), ("offG", _I), ("H", _P), ("ldH", _I), ("offH", _I),
                ("b1", _P), ("Wd1", _P), ("act1", _I),
                ("idx", _P), ("S", _I), ("Nq", _I), ("Nsrc", _I), ("pos_src", _P), ("pos_q", _P),
                ("wimg", _P), ("N", _I),
                ("bias", _P), ("Hq", _P), ("ldHq", _I), ("Wd2", _P), ("act", _I), ("epi_mode", _I),
                ("wvec", _P), ("b0", _F), ("y", _P), ("ldy", _I)]


_lib = None
_tls = threading.local()


def _wrap(fn):
    """Launch on the device that owns the tensors of the call (recorded by ``ptr``), not on whatever device happens to be
    current: kernels, their stream and the per-device function attributes all follow the data (``device='cuda:1'`` arguments,
    nn.DataParallel replicas).  One device per call; mixing devices raises."""
    def call(*args):
        devs = getattr(_tls, "devs", None)
        _tls.devs = None
        if devs:
            if len(devs) > 1:
                raise SsfError("tensors of one call live on different devices: %s" % sorted(devs))
            dev = next(iter(devs))
            if dev != torch.cuda.current_device():
                with torch.cuda.device(dev):
                    return fn(*args)
        return fn(*args)
    call.__name__ = fn.__name__
    return call


class _Lib:
    pass


def lib():
    """Loads libssf_b200.so (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -m ssf_slam_b200.build` (or __graft_entry__.build()); "
                               "ssf_slam_b200 has no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        ns = _Lib()
        ns.raw = L      # the ctypes handle itself (developer scripts reach trace-build symbols through it)
        for name, (codes, res) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = [_CODES[c] for c in codes]
            fn.restype = res
            setattr(ns, name, _wrap(fn) if "p" in codes else fn)
        _lib = ns
    return _lib


_dev_lib = None


def dev_lib():
    """The developer library (tests / scripts only; the product never loads it)."""
    global _dev_lib
    if _dev_lib is None:
        L = ctypes.CDLL(DEV_LIB_PATH)
        ns = _Lib()
        for name, (codes, res) in DEV_SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = [_CODES[c] for c in codes]
            fn.restype = res
            setattr(ns, name, _wrap(fn) if "p" in codes else fn)
        _dev_lib = ns
    return _dev_lib


class SsfError(RuntimeError):
    pass


def check(rc, library=None):
    if rc != 0:
        raise SsfError((library or lib()).ssf_last_error().decode())


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL).  Notes the tensor's device for the launch (see _wrap)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SsfError("ssf_slam_b200 kernels need CUDA tensors (no CPU fallback); got a %s tensor" % t.device)
    if not t.is_contiguous():
        raise SsfError("tensor must be contiguous")
    devs = getattr(_tls, "devs", None)
    if devs is None:
        devs = _tls.devs = set()
    devs.add(t.device.index)
    return t.data_ptr()


def stream():
    """The current torch stream of the device the call's tensors live on (arguments are evaluated left to right and the stream
    comes last in every signature, so every ``ptr`` of the call has run)."""
    devs = getattr(_tls, "devs", None)
    if devs:
        return torch.cuda.current_stream(next(iter(devs))).cuda_stream
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().ssf_launch_count())


_devices_ok = set()


def require_device(device=None):
    """Raises unless ``device`` (default: the current device) is an sm_100 GPU; checked once per device."""
    if not torch.cuda.is_available():
        raise SsfError("no CUDA device: ssf_slam_b200 runs on B200 (sm_100a) only")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _devices_ok:
        with torch.cuda.device(idx):
            check(lib().ssf_require_device())
        _devices_ok.add(idx)
