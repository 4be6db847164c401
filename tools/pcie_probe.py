"""Measure pinned H2D / D2H bandwidth (alone and concurrently) on the box: the floor of the host-buffer API."""
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def up():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
def down():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    up(); down()
for name, fn in (("h2d", up), ("d2h", down), ("both", both)):
    dt = t(fn)
    print("%s: %.2f ms per GiB -> %.1f GB/s%s" % (name, dt * 1e3, n / dt / 1e9, " each way" if name == "both" else ""))
