"""Parity bookkeeping shared by the GPU tests: every comparison is appended to
gpurun_out/parity_report.jsonl so a failed run on the GPU box can be diagnosed from the artefact."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def compare(name, got, want, rtol=1e-8, atol=0.0, mask=None, record=True):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    if mask is not None:
        got, want = got[mask], want[mask]
    diff = np.abs(got - want)
    scale = np.abs(want)
    ok = diff <= atol + rtol * scale
    nbad = int((~ok).sum()) + int(np.isnan(got).sum() != np.isnan(want).sum())
    denom = float(np.nanmax(scale)) if scale.size else 0.0
    rec = dict(name=name, n=int(got.size), bit_identical=int((got == want).sum()),
               max_abs=float(np.nanmax(diff)) if diff.size else 0.0,
               max_abs_over_maxwant=float(np.nanmax(diff) / denom) if denom > 0 else 0.0,
               max_want=denom, nbad=nbad, rtol=rtol, atol=atol)
    if record:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(rec) + "\n")
    assert nbad == 0, f"{name}: {nbad}/{got.size} outside rtol={rtol} atol={atol}; max_abs={rec['max_abs']:.3e} " \
                      f"(max|want|={denom:.3e})"
    return rec
