"""Measured on-chip roofline denominators of the device at hand: L2 -> SM read bandwidth and shared-memory read
bandwidth (csrc/peaks.cu through rrtk_peak_l2_read / rrtk_peak_smem_read), timed with CUDA events.  SURVEY.md section 8(d)
bounds the nearest / radius scan by shared memory and the cfg2 collision walk by L2, and MEASURED_PEAKS.json (driver-
written) carries HBM and bf16 only, so bench.py measures these two on the box it runs on."""
from __future__ import annotations

from . import _lib


def measure(device=0, l2_mb=64, reps=5):
    """{'l2_read_GBps', 'smem_read_GBps', ...}: best of ``reps`` timed launches each (burst figures, like the driver's
    copy bandwidth).  The L2 figure is read from a buffer of ``l2_mb`` MB: above every L1 (256 KB x SMs = 37 MB), below
    the 126 MB L2; the other sizes are reported beside it."""
    import torch
    L = _lib.lib()
    torch.cuda.set_device(device)
    st = torch.cuda.current_stream().cuda_stream
    sink = torch.zeros(4, dtype=torch.int32, device="cuda")
    import ctypes as C
    out = {}
    nb = C.c_int64(0)
    for mb in sorted({16, 32, l2_mb, 64, 96}):
        buf = torch.empty(mb << 20, dtype=torch.uint8, device="cuda").random_(0, 255)
        _lib.check(L.rrtk_peak_l2_read(buf.data_ptr(), buf.numel(), 2, sink.data_ptr(), C.byref(nb), st), "l2 warm")
        best = 0.0
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(L.rrtk_peak_l2_read(buf.data_ptr(), buf.numel(), 20, sink.data_ptr(), C.byref(nb), st), "l2")
            e1.record(); torch.cuda.synchronize()
            best = max(best, nb.value / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        out[f"l2_read_GBps_{mb}MB"] = best
    out["l2_read_GBps"] = out[f"l2_read_GBps_{l2_mb}MB"]
    best = 0.0
    _lib.check(L.rrtk_peak_smem_read(0, 50, sink.data_ptr(), C.byref(nb), st), "smem warm")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.rrtk_peak_smem_read(0, 2000, sink.data_ptr(), C.byref(nb), st), "smem")
        e1.record(); torch.cuda.synchronize()
        best = max(best, nb.value / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    out["smem_read_GBps"] = best
    return out


