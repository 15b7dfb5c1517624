"""k-NN kernel timing at the tracking shape (56 pairs of 2000 x 2000, k = 10) and on a 10^6-row map; prints ms and the fraction
of the popc roof.  UCO_B200_LIB selects a build variant."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ucoslam-cv3_b200", "python"))
import numpy as np, torch
import ucoslam_b200
ctx = ucoslam_b200.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
rng = np.random.default_rng(1)
P, N, K = 56, 2000, 10
base = rng.integers(0, 256, (P + 1, N, 32), dtype=np.uint8)
for p in range(1, P + 1):          # frame p = frame p-1 with a few bits flipped per row (real neighbours exist)
    base[p] = base[p - 1]
    flips = rng.integers(0, 256, (N, 20))
    for j in range(20):
        base[p, np.arange(N), flips[:, j] >> 3] ^= (1 << (flips[:, j] & 7)).astype(np.uint8)
d = torch.from_numpy(base).cuda()
idx = torch.empty((P, N, K), dtype=torch.int32, device="cuda"); dist = torch.empty_like(idx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run():
    ctx.hamming_knn_batch_dev(P, d[1].data_ptr(), N * 32, N, None, d[0].data_ptr(), N * 32, N, None, K, 0, idx.data_ptr(), dist.data_ptr())
def timed(fn, reps=10):
    fn(); ctx.sync()
    ts = []
    with torch.cuda.stream(stream):
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            ts.append((a, b))
    ctx.sync(); torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ts]))
peak = 16 * 148 * 1.965e9
ms = timed(run)
out = {"tracking_56x2000x2000_ms": ms, "frac_popc": P * N * N * 8 / (ms * 1e-3) / peak}
nt = 1_000_000
t = torch.from_numpy(rng.integers(0, 256, (nt, 32), dtype=np.uint8)).cuda()
q = t[torch.randint(0, nt, (N,), device="cuda")].clone(); q[:, :2] ^= 0x5A
i2 = torch.empty((N, K), dtype=torch.int32, device="cuda"); d2 = torch.empty_like(i2)
ms2 = timed(lambda: ctx.hamming_knn_dev(q.data_ptr(), N, t.data_ptr(), nt, K, 0, i2.data_ptr(), d2.data_ptr()), 5)
out.update({"map_2000x1e6_ms": ms2, "frac_popc_map": N * nt * 8 / (ms2 * 1e-3) / peak})
print(json.dumps(out))
del stream, flush
torch.cuda.synchronize()
