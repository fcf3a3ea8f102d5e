"""Measure the TF32 dense-matmul rate of this GPU the way MEASURED_PEAKS.json measures bf16
(torch.matmul 8192^3, best of 10 = burst; back to back for 3 s = sustained)."""
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
for _ in range(3): a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
t0 = time.time(); k = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(10): a @ b
    k += 10
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / k
print(json.dumps({"tf32_tflops": 2 * n ** 3 / best / 1e9, "tf32_tflops_sustained": 2 * n ** 3 / sus / 1e9,
                  "how": "torch.matmul fp32 inputs, allow_tf32=True, 8192^3"}))
