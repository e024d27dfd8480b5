"""Generate tests/golden/*.npz by running the UNMODIFIED reference VAE on CPU (fp32).

TEST INFRASTRUCTURE.  Runs only in the authoring container, where /root/reference exists:

    python oracle/make_golden.py            # rewrites tests/golden/

The reference is imported from /root/reference/hyvideo/vae through the `diffusers` shim in
oracle/_refshim (diffusers is not installed here).  Weights come from oracle/weights.py via
load_state_dict(strict=True); inputs are seeded (weights.make_video / make_latent), so fixtures hold
outputs only.  Every case records the arguments needed to replay it.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_refshim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from oracle import weights as W  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def build_reference(cfg, seed=0):
    from hyvideo.vae.autoencoder_kl_causal_3d import AutoencoderKLCausal3D
    m = AutoencoderKLCausal3D.from_config(cfg)
    m.load_state_dict(W.make_state_dict(cfg, seed), strict=True)
    return m.eval().requires_grad_(False)


def apply_t_ops(m, t_ops):
    from hyvideo.vae import _apply_t_ops_config_to_vae
    _apply_t_ops_config_to_vae(m, t_ops)


def base_t_ops():
    with open("/root/reference/t_ops_config.json") as f:
        return json.load(f)


def save(name, meta, **arrays):
    os.makedirs(OUT, exist_ok=True)
    np.savez(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), **{k: v.numpy() for k, v in arrays.items()})
    print(name, {k: tuple(v.shape) for k, v in arrays.items()})


def model_case(name, cfg_name, shape, spatial=False, temporal=False, t_ops=None, decode_latent=True):
    cfg = getattr(W, cfg_name)
    m = build_reference(cfg)
    if t_ops is not None:
        apply_t_ops(m, t_ops)
    m.enable_spatial_tiling(spatial)
    m.enable_temporal_tiling(temporal)
    x = W.make_video(shape)
    post = m.encode(x).latent_dist
    arrays = dict(moments=post.parameters)
    if decode_latent:
        arrays["dec"] = m.decode(post.mode()).sample
    save(name, dict(cfg=cfg_name, shape=list(shape), spatial=spatial, temporal=temporal, t_ops=t_ops,
                    video_seed=1234, weight_seed=0), **arrays)


def op_cases():
    """Per-op fixtures from the reference's own block classes."""
    from hyvideo.vae.unet_causal_3d_blocks import (CausalConv3d, ResnetBlockCausal3D, UpsampleCausal3D,
                                                   UNetMidBlockCausal3D, prepare_causal_attention_mask)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 32, 5, 12, 10, generator=g)
    arrays = {"x": x}
    arrays["conv_w"] = torch.randn(64, 32, 3, 3, 3, generator=g) * 0.05
    arrays["conv_b"] = torch.randn(64, generator=g) * 0.1
    arrays["conv1_w"] = torch.randn(48, 32, 1, 1, 1, generator=g) * 0.2
    for tag, stride in (("s111", (1, 1, 1)), ("s122", (1, 2, 2)), ("s222", (2, 2, 2)), ("s422", (4, 2, 2))):
        c = CausalConv3d(32, 64, 3, stride=stride)
        with torch.no_grad():
            c.conv.weight.copy_(arrays["conv_w"])
            c.conv.bias.copy_(arrays["conv_b"])
            arrays[f"conv_{tag}_y"] = c(x)
    c1 = CausalConv3d(32, 48, 1, bias=False)
    with torch.no_grad():
        c1.conv.weight.copy_(arrays["conv1_w"])
        arrays["conv1_y"] = c1(x)
    with torch.no_grad():
        for tag, fac in (("u222", (2, 2, 2)), ("u122", (1, 2, 2))):
            arrays[f"up_{tag}_y"] = UpsampleCausal3D(32, use_conv=False, upsample_factor=fac)(x)
        arrays["up_u222_T1_y"] = UpsampleCausal3D(32, use_conv=False, upsample_factor=(2, 2, 2))(x[:, :, :1])
        r = ResnetBlockCausal3D(in_channels=32, out_channels=64, temb_channels=None, groups=32, eps=1e-6)
        rsd = {k: torch.randn(v.shape, generator=g) * (0.05 if v.ndim > 1 else 0.2) + (1.0 if "norm" in k and k.endswith("weight") else 0.0)
               for k, v in r.state_dict().items()}
        r.load_state_dict(rsd)
        for k, v in rsd.items():
            arrays["res_sd." + k] = v
        arrays["res_y"] = r(x, temb=None)
        mb = UNetMidBlockCausal3D(in_channels=64, temb_channels=None, resnet_groups=32, attention_head_dim=64,
                                  resnet_eps=1e-6, resnet_act_fn="silu", output_scale_factor=1, add_attention=True)
        msd = {k: torch.randn(v.shape, generator=g) * (0.05 if v.ndim > 1 else 0.2) + (1.0 if "norm" in k and k.endswith("weight") else 0.0)
               for k, v in mb.state_dict().items()}
        mb.load_state_dict(msd)
        for k, v in msd.items():
            arrays["mid_sd." + k] = v
        xm = torch.randn(1, 64, 3, 6, 5, generator=g)
        arrays["mid_x"] = xm
        arrays["mid_y"] = mb(xm)
        arrays["mask_3_4"] = prepare_causal_attention_mask(3, 4, torch.float32, "cpu")
    save("ops", dict(note="per-op outputs of the reference block classes, fp32 CPU"), **arrays)


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    op_cases()
    model_case("small_untiled", "SMALL_CONFIG", (1, 3, 9, 32, 32))
    model_case("small_untiled_b2", "SMALL_CONFIG", (2, 3, 5, 24, 40))
    model_case("small_spatial", "SMALL_CONFIG", (1, 3, 5, 72, 88), spatial=True)
    model_case("small_temporal", "SMALL_CONFIG", (1, 3, 33, 32, 32), temporal=True)
    model_case("small_tiled", "SMALL_CONFIG", (1, 3, 29, 56, 40), spatial=True, temporal=True)
    model_case("hy_untiled", "HY_VAE_CONFIG", (1, 3, 5, 32, 32))
    # t-ops (config 5): pool in the encoder, interp in the decoder, and a run-time T-stride of 4
    t1 = base_t_ops()
    t1["encoder"]["down_blocks"][1]["enable_t_pool_before_block"] = [True, False]
    t1["decoder"]["up_blocks"][2]["enable_t_interp_after_block"] = [False, True, False]
    model_case("small_tops_pool_interp", "SMALL_CONFIG", (1, 3, 17, 32, 32), t_ops=t1)
    t2 = base_t_ops()
    t2["encoder"]["down_blocks"][1]["downsample_stride"] = [4, 2, 2]
    t2["encoder"]["mid_block"]["enable_t_pool_after_block"] = [True, False]
    t2["decoder"]["up_blocks"][0]["enable_t_interp_before_block"] = [True, False, False]
    t2["decoder"]["up_blocks"][0]["interp_t_scale_factor"] = 2
    model_case("small_tops_stride4", "SMALL_CONFIG", (1, 3, 17, 32, 32), t_ops=t2)


if __name__ == "__main__":
    main()
