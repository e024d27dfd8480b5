"""Fused attention kernel (hyvae_attn_block_causal) vs the GEMM -> softmax -> GEMM schedule at the canonical mid-block
shape (L = 17 x 32 x 32, D = 512): max abs / relative difference and CUDA-event time of each.  usage: bench_attn.py [T hw D]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hunyuanvideo_efficiency_b200 import _native as N
from hunyuanvideo_efficiency_b200.vae.blocks import _gemm_nt
from hunyuanvideo_efficiency_b200._native import Vol

dev = torch.device("cuda:0")
T, hw, D = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (17, 1024, 512)
only_fused = len(sys.argv) > 4 and sys.argv[4] == "fused"
L = T * hw
dt = torch.float16
torch.manual_seed(0)
q = torch.randn(L, D, device=dev).to(dt)
k = torch.randn(L, D, device=dev).to(dt)
vt = torch.randn(D, L, device=dev).to(dt)
bv = torch.randn(D, device=dev)
scale = D ** -0.5


def unfused():
    qv = Vol(1, 1, 1, L, D, dt, dev, tensor=q.reshape(1, 1, 1, L, D))
    s = _gemm_nt(qv, k, None, L, out_dtype=torch.float32)
    p = N.softmax_frame_causal(s.t.reshape(1, L, L), hw, scale, dt)
    pv = Vol(1, 1, 1, L, L, dt, dev, tensor=p.reshape(1, 1, 1, L, L))
    return _gemm_nt(pv, vt, bv, D).t.reshape(L, D)


def fused():
    return N.attn_block_causal(q, k, vt, bv, hw, scale)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


of = fused()
torch.cuda.synchronize()
print(f"fused ran: L={L} n_hw={hw} D={D} finite={bool(torch.isfinite(of.float()).all())}")
if not only_fused:
    ou = unfused()
    d = (of.float() - ou.float())
    print(f"fused vs unfused: max abs {d.abs().max().item():.3e}  rel {(d.norm() / ou.float().norm()).item():.3e}")
    tu = timeit(unfused)
tf = timeit(fused)
vis = sum((f + 1) * hw for f in range(T)) * hw  # visible (query, key) pairs
flops_min = 4.0 * vis * D
print(f"fused   {tf:8.3f} ms  {flops_min / tf / 1e9:8.1f} TFLOP/s of the visible QK^T + PV work (dense {4.0 * L * L * D / tf / 1e9:.1f})")
if not only_fused:
    print(f"unfused {tu:8.3f} ms  -> fused is {tu / tf:.2f}x")
