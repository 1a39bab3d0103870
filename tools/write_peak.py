"""pure-write / pure-read / copy bandwidth of the device with simple grid-stride kernels (torch ops), for context next to
the roofline fractions of the store-bound assembly kernel"""
import json, torch
n = 10_700_000_000 // 8
x = torch.empty(n, dtype=torch.float64, device="cuda")
y = torch.empty(n, dtype=torch.float64, device="cuda")
def timed(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
out = {}
t = timed(lambda: x.fill_(1.0)); out["fill (write only)"] = 8 * n / t / 1e9
t = timed(lambda: torch.cuda.memset if False else x.zero_()); out["zero_ (memset)"] = 8 * n / t / 1e9
t = timed(lambda: y.copy_(x)); out["copy (1R:1W)"] = 16 * n / t / 1e9
t = timed(lambda: x.sum()); out["sum (read only)"] = 8 * n / t / 1e9
print(json.dumps(out))
