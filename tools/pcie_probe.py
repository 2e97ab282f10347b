import torch, time
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n // 2, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
a = t(lambda: h.copy_(d, non_blocking=True))
b = t(lambda: d.copy_(h, non_blocking=True))
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
c = t(both)
print(f"D2H 1 GiB alone {n / a / 1e9:.1f} GB/s; H2D alone {n / b / 1e9:.1f} GB/s; D2H 1 GiB with 0.5 GiB H2D concurrently: {c * 1e3:.1f} ms = {n / c / 1e9:.1f} GB/s D2H-equivalent")

# the host path's traffic pattern without any compute: 18 uploads of 32 MiB queued on one
# stream, a 61 MB download per chunk on another as soon as "its" upload has landed
nin, nout, K = 588_538_415, n, 18
hin = torch.empty(nin, dtype=torch.uint8, pin_memory=True)
din = torch.empty(nin, dtype=torch.uint8, device="cuda")
ci, co = (nin + K - 1) // K, (nout + K - 1) // K
def pipe():
    evs = []
    with torch.cuda.stream(s1):
        for k in range(K):
            din[k * ci:(k + 1) * ci].copy_(hin[k * ci:(k + 1) * ci], non_blocking=True)
            e = torch.cuda.Event(); e.record(s1); evs.append(e)
    with torch.cuda.stream(s2):
        for k in range(K):
            s2.wait_event(evs[k])
            h[k * co:(k + 1) * co].copy_(d[k * co:(k + 1) * co], non_blocking=True)
p = t(pipe, 3)
print(f"chunked pipeline pattern, copies only: {p * 1e3:.1f} ms = {nout / p / 1e9:.1f} GB/s decoded-equivalent")
