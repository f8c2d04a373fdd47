"""Measured on-chip roofline denominators of the device at hand: L2 -> SM read bandwidth and shared-memory read
bandwidth (csrc/peaks.cu), timed with CUDA events.  bench.py uses the same function."""
import sys
sys.path.insert(0, ".")
import torch
from rrtplanner_b200 import _lib


def measure(device=0, l2_mb=48, reps=5):
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


if __name__ == "__main__":
    import json
    print(json.dumps(measure()))
