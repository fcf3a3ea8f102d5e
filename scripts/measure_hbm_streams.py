"""HBM stream rates on this GPU (CUDA events, 2 GiB buffers >> 126 MB L2): write-only, read-only, copy.
The roofline denominator stays MEASURED_PEAKS.json (copy: read + write bytes); this records how far a one-directional
stream -- the grouped-tensor writes and reads of the inter conv -- can get."""
import json, sys, torch
dev = torch.device("cuda:0")
n = (2 << 30) // 4
x = torch.empty(n, dtype=torch.float32, device=dev)
y = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
gb = n * 4 / 1e9
out = {"write_only_GBps": gb / t(lambda: x.zero_()), "read_only_GBps": gb / t(lambda: x.sum()),
       "copy_GBps_read_plus_write": 2 * gb / t(lambda: y.copy_(x)), "buffer_GiB": 2}
print(json.dumps(out))
