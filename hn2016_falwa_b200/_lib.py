"""ctypes binding of libfalwa_b200.so (the C ABI declared in include/falwa_b200.h).

There is deliberately no fallback: if the shared library is missing or CUDA is unavailable, importing
the compute entry points raises.  PyTorch is used only for device memory, streams and events.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfalwa_b200.so")

c_dp = ctypes.c_void_p   # device (or host) pointer
c_i = ctypes.c_int
c_d = ctypes.c_double

_SIGNATURES = {
    "falwa_b200_compute_qgpv": [c_i] * 3 + [c_dp] * 6 + [c_d] * 6 + [c_dp] * 2 + [c_dp],
    "falwa_b200_compute_qgpv_direct_inv": [c_i] * 4 + [c_dp] * 8 + [c_d] * 6 + [c_dp] * 2 + [c_dp],
    "falwa_b200_compute_reference_states": [c_dp] * 4 + [c_i] * 6 + [c_d] * 10 + [c_dp] * 3 + [c_dp, c_dp, c_dp],
    "falwa_b200_compute_qref_and_fawa_first": [c_dp] * 5 + [c_i] * 7 + [c_d] * 8 + [c_dp] * 7 + [c_dp, c_dp, c_dp],
    "falwa_b200_matrix_b4_inversion": [c_i] * 6 + [c_dp] * 5 + [c_d] * 6 + [c_dp] * 4 + [c_dp],
    "falwa_b200_matrix_after_inversion": [c_i] * 3 + [c_dp] * 6 + [c_dp],
    "falwa_b200_upward_sweep": [c_i] * 5 + [c_dp] * 8 + [c_d] * 6 + [c_dp],
    "falwa_b200_compute_flux_dirinv_nshem": [c_dp] * 9 + [c_i] * 7 + [c_d] * 7 + [c_dp] * 9 + [c_dp, c_i, c_dp],
    "falwa_b200_barotropic_eqvlat_lwa": [c_i] * 4 + [c_dp] * 4 + [c_d] + [c_dp] * 2 + [c_i, c_dp],
    "falwa_b200_eqvlat_2d": [c_i] * 5 + [c_dp] * 4 + [c_d] + [c_dp] * 3,
    "falwa_b200_lwa_2d": [c_i] * 3 + [c_dp] * 6 + [c_i, c_dp],
    "falwa_b200_ingest_field": [c_i] * 5 + [c_dp, c_i, c_d, c_d, c_i, c_d, c_i, c_i, c_dp, c_dp],
    "falwa_b200_compute_lwa_only_nhn22": [c_dp] * 3 + [c_i] * 6 + [c_d] * 7 + [c_dp] * 4 + [c_i, c_dp],
}

_lib = None


class FalwaB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FalwaB200Error(
            f"{LIB_PATH} not found. Build it with `python -m hn2016_falwa_b200.build` "
            "(nvcc, sm_100a). falwa_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.falwa_b200_version.restype = ctypes.c_char_p
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_i
    lib.falwa_b200_launch_count.restype = ctypes.c_longlong
    lib.falwa_b200_launch_count.argtypes = []
    lib.falwa_b200_profile_enable.argtypes = [c_i]
    lib.falwa_b200_profile_enable.restype = c_i
    lib.falwa_b200_profile_read.argtypes = [ctypes.c_char_p, c_i]
    lib.falwa_b200_profile_read.restype = c_i
    _lib = lib
    return lib


def launch_count():
    """Number of CUDA kernels this library has launched so far in this process."""
    return int(load().falwa_b200_launch_count())


def profile_enable(on=True):
    check(load().falwa_b200_profile_enable(int(bool(on))), "profile_enable")


def profile_read():
    """-> {kernel_name: (launches, total_ms)} since profile_enable(True); synchronises the device."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(load().falwa_b200_profile_read(buf, len(buf)), "profile_read")
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.split()
        out[name] = (int(n), float(ms))
    return out


def register(name, argtypes):
    """Register the signature of a Section-B entry point (called by the modules that use them)."""
    lib = load()
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = c_i
    _SIGNATURES[name] = argtypes
    return fn


def declared_symbols():
    """Every `int falwa_b200_*(` / `const char* falwa_b200_*(` symbol declared in include/falwa_b200.h."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "falwa_b200.h")
    text = open(hdr).read()
    return sorted(set(re.findall(r"\b(falwa_b200_[a-z0-9_]+)\s*\(", text)))


def require_cuda():
    if not torch.cuda.is_available():
        raise FalwaB200Error("falwa_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        msg = {-1: "invalid dimension/index argument", -2: "NULL pointer argument",
               -3: "size not supported by the kernels"}.get(rc, "invalid argument")
        raise ValueError(f"{what}: {msg} (falwa_b200 status {rc})")
    raise FalwaB200Error(f"{what}: CUDA error {rc}")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t):
    """Device pointer of a contiguous float64 / int32 / int64 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "falwa_b200 needs contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def hptr(a):
    """Host pointer of a contiguous float64 numpy array."""
    return a.ctypes.data_as(ctypes.c_void_p)
