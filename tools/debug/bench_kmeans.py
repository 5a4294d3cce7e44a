"""Time the k-means update and the query self-attention at the cfg2 sizes (32 clips, OS8 = 13122 pixels per clip)."""
import torch
from axial_vs_b200 import ops

N, L = 32, 128
for M in (882, 3362, 13122):
    lg = torch.randn(N, L, M, device="cuda")
    pv = torch.randn(N, 256, M, device="cuda")
    for _ in range(3):
        ops.kmeans_update(lg, pv)
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(10):
        ops.kmeans_update(lg, pv)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"kmeans M={M}: {ms*1e3:.1f} us, {N*M*(L+256)*4/ms/1e6:.0f} GB/s")
q = torch.randn(N, 8, 16, L, device="cuda"); v = torch.randn(N, 8, 32, L, device="cuda")
s = torch.ones(8, 2, device="cuda"); va = torch.ones(256, 2, device="cuda")
for _ in range(3):
    ops.query_self_attn(q, q, v, s, va)
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record()
for _ in range(10):
    ops.query_self_attn(q, q, v, s, va)
b.record(); torch.cuda.synchronize()
print(f"query_self_attn N={N}: {a.elapsed_time(b)/10*1e3:.1f} us")
