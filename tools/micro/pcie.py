"""Raw PCIe rates of the box, for the e2e number of bench.py: pinned host <-> device copies of 512 MiB, one direction at a time and
both at once on two streams (what the chunked host pipeline of ab_tensor_op_host does)."""
import torch
n = 512 << 20
h0, h1 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d0, d1 = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s0.synchronize(); s1.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def h2d():
    with torch.cuda.stream(s0): d0.copy_(h0, non_blocking=True)
def d2h():
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
def both():
    h2d(); d2h()
import time
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(name, "%.1f GB/s per direction" % (n / dt / 1e9))
