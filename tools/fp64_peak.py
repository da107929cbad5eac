"""FP64 DFMA rate of the device: the denominator of every FP64-pipe utilisation quoted in DESIGN.md.
Prints one JSON line: achieved DFMA/s, TFLOP/s (2 flops per DFMA) and DFMA per clock per SM at the sampled SM clock."""
import ctypes
import json
import sys

sys.path.insert(0, ".")
import torch
from hn2016_falwa_b200 import _lib

lib = _lib.load()
_lib.require_cuda()
sms = torch.cuda.get_device_properties(0).multi_processor_count
blocks, iters = sms * 16, 1 << 15
sink = torch.zeros(1, dtype=torch.float64, device="cuda")
ms = ctypes.c_float(0)
lib.falwa_b200_debug_dfma_rate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.c_void_p]
best = None
for _ in range(5):
    rc = lib.falwa_b200_debug_dfma_rate(blocks, iters, sink.data_ptr(), ctypes.byref(ms), None)
    assert rc == 0, rc
    best = ms.value if best is None else min(best, ms.value)
dfma = blocks * 256 * 8.0 * iters
try:
    import pynvml
    pynvml.nvmlInit()
    mhz = pynvml.nvmlDeviceGetClockInfo(pynvml.nvmlDeviceGetHandleByIndex(0), pynvml.NVML_CLOCK_SM)
except Exception:
    mhz = None
rate = dfma / (best * 1e-3)
print(json.dumps({"what": "fp64 DFMA micro-benchmark", "sms": sms, "blocks": blocks, "threads": 256, "dfma_per_thread": 8 * iters,
                  "ms": best, "dfma_per_s": rate, "tflops": 2 * rate / 1e12, "sm_mhz_after": mhz,
                  "dfma_per_clk_per_sm": (rate / sms / (mhz * 1e6)) if mhz else None}))
