import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from hunyuanvideo_efficiency_b200 import _native as N
from oracle import weights as W, vae_oracle as O
dev = torch.device("cuda:0")
torch.manual_seed(0)

def conv_check(B, Cin, Cout, T, H, W_, k=3, stride=(1,1,1), out_dtype=None, res=False):
    pad = (k-1, k//2, k//2)
    x = N.Vol(B, T, H, W_, Cin, torch.bfloat16, dev, (0,0,0)); x.t.normal_()
    xp = N.pad_upsample(x, (1,1,1), pad)
    w = (torch.randn(k**3, Cout, Cin, device=dev) / (k**3*Cin)**0.5).bfloat16()
    b = torch.randn(Cout, device=dev)
    To, Ho, Wo = N.conv_out_dims(T, H, W_, stride)
    r = None
    if res:
        r = N.Vol(B, To, Ho, Wo, Cout, torch.bfloat16, dev); r.t.normal_()
    y1 = N.conv3d_tc(xp, w, b, k, stride, Cout, residual=r, out_dtype=out_dtype, round_like_ref=False)
    y2 = N.conv3d_tc(xp, w, b, k, stride, Cout, residual=r, out_dtype=out_dtype, round_like_ref=False)
    yd = N.conv3d_direct(x, w, b, k, stride, Cout, residual=r, out_dtype=out_dtype, round_like_ref=False)
    torch.cuda.synchronize()
    det = torch.equal(y1.t, y2.t)
    err = O.rel_err(yd.t.float().cpu(), y1.t.float().cpu())
    mx = (yd.t.float() - y1.t.float()).abs().max().item()
    print(f"conv B={B} Cin={Cin} Cout={Cout} T={T} H={H} W={W_} k={k} s={stride} out={out_dtype} res={res}: deterministic={det} rel_vs_direct={err:.2e} maxabs={mx:.3e}", flush=True)

conv_check(1, 64, 64, 3, 8, 16)
conv_check(1, 128, 128, 4, 64, 64)
conv_check(1, 128, 128, 4, 64, 64, out_dtype=torch.float32)
conv_check(1, 256, 256, 3, 64, 64)
conv_check(1, 512, 512, 3, 32, 32, res=True)
conv_check(1, 128, 256, 5, 72, 40)
conv_check(1, 512, 512, 1, 1, 5120, k=1)
conv_check(1, 512, 5120, 1, 1, 5120, k=1, out_dtype=torch.float32)
conv_check(1, 5120, 512, 1, 1, 5120, k=1)
conv_check(2, 128, 128, 9, 64, 64, stride=(2,2,2))

# GN determinism + whole-decoder determinism on both paths
from hunyuanvideo_efficiency_b200.vae import AutoencoderKLCausal3D
cfg = W.HY_VAE_CONFIG
m = AutoencoderKLCausal3D.from_config(cfg); m.load_state_dict(W.make_state_dict(cfg)); m = m.to(torch.bfloat16).to(dev).eval()
z = W.make_latent((1, 16, 3, 16, 16)).to(dev, torch.bfloat16)
with torch.no_grad():
    for force in ("0", "1"):
        os.environ["HYVAE_FORCE_DIRECT"] = force
        d1 = m.decode(z).sample; d2 = m.decode(z).sample
        torch.cuda.synchronize()
        print("decoder force_direct=", force, "deterministic:", torch.equal(d1, d2), "maxdiff", (d1.float()-d2.float()).abs().max().item(), flush=True)
        if force == "0": dtc = d1
        else: print("tc vs direct rel:", O.rel_err(d1.float().cpu(), dtc.float().cpu()))
    os.environ["HYVAE_FORCE_DIRECT"] = "0"
    # layer-by-layer: find the first non-deterministic op on the tc path
    from hunyuanvideo_efficiency_b200._native import Vol
    v = Vol.from_ncthw(z, dtype=torch.bfloat16)
    def twice(name, fn, inp):
        a = fn(inp); b = fn(inp); torch.cuda.synchronize()
        print(f"  {name}: det={torch.equal(a.t, b.t)} shape={a.dims}", flush=True)
        return a
    v = twice("post_quant", m.post_quant_conv.forward_vol, v)
    v = twice("conv_in", m.decoder.conv_in.forward_vol, v)
    mb = m.decoder.mid_block
    v = twice("mid.res0", mb.resnets[0].forward_vol, v)
    v = twice("mid.attn", mb.attentions[0].forward_vol, v)
    v = twice("mid.res1", mb.resnets[1].forward_vol, v)
    for i, blk in enumerate(m.decoder.up_blocks):
        for j, r in enumerate(blk.resnets):
            v = twice(f"up{i}.res{j}", r.forward_vol, v)
        if blk.upsamplers is not None:
            v = twice(f"up{i}.upsample", blk.upsamplers[0].forward_vol, v)
