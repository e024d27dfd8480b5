"""CPU study: which storage / operand rounding points dominate the 16-bit error of the VAE decoder.
Emulates the GPU pipeline (fp32 accumulation everywhere) with explicit rounding hooks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
from oracle import vae_oracle as O, weights as W
torch.set_grad_enabled(False)

def q(x, dt):
    return x if dt is None else x.to(dt).float()

class P:  # rounding policy
    def __init__(s, w, a, c1, st, g=None, attn=None): s.w, s.a, s.c1, s.st, s.g, s.attn = w, a, c1, st, g, attn or a

def conv(x, w, b, p, stride=(1,1,1)):
    return O.causal_conv3d(q(x, p.a), q(w, p.w), b, stride)

def gn_silu(x, w, b, g, p):
    return F.silu(q(O.group_norm(x, w, b, g), p.g))

def resnet(sd, pre, x, g, p):
    h = conv(gn_silu(x, sd[pre+"norm1.weight"], sd[pre+"norm1.bias"], g, p), sd[pre+"conv1.conv.weight"], sd[pre+"conv1.conv.bias"], p)
    h = q(h, p.c1)
    h = conv(gn_silu(h, sd[pre+"norm2.weight"], sd[pre+"norm2.bias"], g, p), sd[pre+"conv2.conv.weight"], sd[pre+"conv2.conv.bias"], p)
    if pre+"conv_shortcut.conv.weight" in sd:
        x = q(conv(x, sd[pre+"conv_shortcut.conv.weight"], sd[pre+"conv_shortcut.conv.bias"], p), p.st)
    return q(x + h, p.st)

def attn(sd, pre, x, g, p):
    B,C,T,H,Wd = x.shape
    seq = x.permute(0,2,3,4,1).reshape(B,T*H*Wd,C)
    hn = q(O.group_norm(seq.transpose(1,2), sd[pre+"group_norm.weight"], sd[pre+"group_norm.bias"], g).transpose(1,2), p.attn)
    lin = lambda t, n: F.linear(t, q(sd[pre+n+".weight"], p.w), sd[pre+n+".bias"])
    qq, kk, vv = q(lin(hn,"to_q"), p.attn), q(lin(hn,"to_k"), p.attn), q(lin(hn,"to_v"), p.attn)
    s = qq @ kk.transpose(1,2) * C**-0.5 + O.frame_causal_mask(T, H*Wd)[None]
    o = q(q(torch.softmax(s, -1), p.attn) @ vv, p.attn)
    o = F.linear(o, q(sd[pre+"to_out.0.weight"], p.w), sd[pre+"to_out.0.bias"]) + seq
    return q(o, p.st).reshape(B,T,H,Wd,C).permute(0,4,1,2,3)

def decoder(sd, cfg, z, p):
    g = cfg["norm_num_groups"]
    z = F.conv3d(q(z, p.a), q(sd["post_quant_conv.weight"], p.w), sd["post_quant_conv.bias"])
    x = q(conv(z, sd["decoder.conv_in.conv.weight"], sd["decoder.conv_in.conv.bias"], p), p.st)
    x = resnet(sd, "decoder.mid_block.resnets.0.", x, g, p)
    x = attn(sd, "decoder.mid_block.attentions.0.", x, g, p)
    x = resnet(sd, "decoder.mid_block.resnets.1.", x, g, p)
    for i, fac in enumerate(O.decoder_upfactors(cfg)):
        for j in range(3):
            x = resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}.", x, g, p)
        if fac is not None:
            pre = f"decoder.up_blocks.{i}.upsamplers.0.conv.conv."
            x = q(conv(O.upsample_nearest_causal(x, fac), sd[pre+"weight"], sd[pre+"bias"], p), p.st)
    x = gn_silu(x, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], g, p)
    return conv(x, sd["decoder.conv_out.conv.weight"], sd["decoder.conv_out.conv.bias"], p)

bf, hf = torch.bfloat16, torch.float16
policies = {
  "fp32 (check)": P(None, None, None, None),
  "all bf16 (reference-like)": P(bf, bf, bf, bf, bf),
  "all fp16": P(hf, hf, hf, hf, hf),
  "W bf16 only": P(bf, None, None, None),
  "A bf16 only": P(None, bf, None, None),
  "stream bf16 only": P(None, None, None, bf),
  "c1 bf16 only": P(None, None, bf, None),
  "W bf16, A fp16, c1 fp16, stream bf16": P(bf, hf, hf, bf),
  "W bf16, A fp16, c1 fp16, stream fp32": P(bf, hf, hf, None),
  "W bf16, A fp16, c1 bf16, stream bf16": P(bf, hf, bf, bf),
  "W bf16, A bf16, c1 fp32, stream fp32": P(bf, bf, None, None),
  "W fp16(of fp32), A fp16, c1 fp16, stream bf16": P(hf, hf, hf, bf),
  "W fp16, A fp16, c1 fp16, stream fp32": P(hf, hf, hf, None),
}
for cfgname, zshape in (("SMALL_CONFIG", (1,16,3,4,4)), ("HY_VAE_CONFIG", (1,16,2,4,4))):
    cfg = getattr(W, cfgname); sd = W.make_state_dict(cfg)
    z = W.make_latent(zshape)
    ref = O.decode(sd, cfg, z, O.Tiling())
    for name, p in policies.items():
        d = decoder(sd, cfg, z, p)
        print(f"{cfgname:14s} {name:48s} rel={O.rel_err(ref, d):.4f} psnr={O.psnr(ref, d):.1f}", flush=True)
