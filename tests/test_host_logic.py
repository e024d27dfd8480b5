"""Host-side mirror of the reference interface (runs without a GPU)."""
import copy
import json

import pytest
import torch

from hunyuanvideo_efficiency_b200.vae import (AutoencoderKLCausal3D, DiagonalGaussianDistribution,
                                              _apply_t_ops_config_to_vae, load_vae)
from oracle import weights as W


def test_state_dict_matches_reference_keys():
    m = AutoencoderKLCausal3D.from_config(W.HY_VAE_CONFIG)
    spec = W.state_dict_spec(W.HY_VAE_CONFIG)
    sd = m.state_dict()
    assert list(sd) and set(sd) == set(spec) and len(sd) == 248
    assert all(tuple(sd[k].shape) == spec[k] for k in spec)
    m.load_state_dict(W.make_state_dict(W.HY_VAE_CONFIG), strict=True)


def test_config_and_tiling_attributes():
    m = AutoencoderKLCausal3D.from_config(W.HY_VAE_CONFIG)
    assert m.config.block_out_channels == [128, 256, 512, 512] and m.config.scaling_factor == 0.476986
    assert not hasattr(m.config, "shift_factor")
    assert (m.tile_sample_min_size, m.tile_latent_min_size, m.tile_sample_min_tsize, m.tile_latent_min_tsize) == (256, 32, 64, 16)
    assert m.tile_overlap_factor == 0.25
    m.enable_tiling()
    assert m.use_spatial_tiling and m.use_temporal_tiling
    m.disable_tiling()
    assert not (m.use_spatial_tiling or m.use_temporal_tiling)
    m.enable_slicing()
    assert m.use_slicing
    assert m.to(torch.bfloat16).dtype == torch.bfloat16


def test_downsample_strides_follow_reference_plan():
    m = AutoencoderKLCausal3D.from_config(W.HY_VAE_CONFIG)
    s = [b.downsamplers[0].conv.conv.stride if b.downsamplers is not None else None for b in m.encoder.down_blocks]
    assert s == [(1, 2, 2), (2, 2, 2), (2, 2, 2), None]
    u = [b.upsamplers[0].upsample_factor if b.upsamplers is not None else None for b in m.decoder.up_blocks]
    assert u == [(1, 2, 2), (2, 2, 2), (2, 2, 2), None]


def test_t_ops_injection_and_errors():
    m = AutoencoderKLCausal3D.from_config(W.SMALL_CONFIG)
    base = {
        "encoder": {"down_blocks": [{"block_index": 1, "pool_t_kernel": 3, "pool_t_stride": 2,
                                     "enable_t_pool_before_block": [True, False], "enable_t_pool_after_block": [False, False],
                                     "downsample_stride": [4, 2, 2]}],
                    "mid_block": {"enable_t_pool_before_block": [False, False], "enable_t_pool_after_block": [False, True]}},
        "decoder": {"up_blocks": [{"block_index": 0, "enable_t_interp_before_block": [False, True, False],
                                   "enable_t_interp_after_block": [False, False, False], "interp_t_scale_factor": 2}],
                    "mid_block": {"enable_t_pool_before_block": [False, False], "enable_t_pool_after_block": [False, False]}},
    }
    _apply_t_ops_config_to_vae(m, base)
    assert m.encoder.down_blocks[1].downsamplers[0].conv.conv.stride == (4, 2, 2)
    assert m.encoder.down_blocks[1].resnet_pool_configs[0] == {"enable_before": True, "enable_after": False, "kernel": 3, "stride": 2}
    assert m.decoder.up_blocks[0].resnet_interp_configs[1]["enable_before"] is True
    bad = copy.deepcopy(base)
    bad["encoder"]["down_blocks"][0]["enable_t_pool_before_block"] = [True]
    with pytest.raises(ValueError):
        _apply_t_ops_config_to_vae(m, bad)
    bad = copy.deepcopy(base)
    del bad["decoder"]["mid_block"]  # the reference raises on the empty default too (unet_causal_3d_blocks.py:629-633)
    with pytest.raises(ValueError):
        _apply_t_ops_config_to_vae(m, bad)


def test_load_vae_roundtrip(tmp_path):
    cfg = dict(W.SMALL_CONFIG)
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    with pytest.raises(AssertionError):
        load_vae(vae_path=str(tmp_path))
    sd = W.make_state_dict(cfg)
    torch.save({"state_dict": {"vae." + k: v for k, v in sd.items()}}, tmp_path / "pytorch_model.pt")
    vae, path, sr, tr = load_vae(vae_path=str(tmp_path), vae_precision="bf16")
    assert (sr, tr) == (8, 4) and path == str(tmp_path) and vae.dtype == torch.bfloat16 and not vae.training
    assert all(not p.requires_grad for p in vae.parameters())
    assert torch.equal(vae.state_dict()["quant_conv.weight"].float(), sd["quant_conv.weight"].bfloat16().float())


def test_posterior_matches_oracle():
    from oracle import vae_oracle as O
    mom = torch.randn(2, 32, 3, 4, 4) * 20
    p = DiagonalGaussianDistribution(mom)
    mean, logvar = O.posterior_mean_logvar(mom)
    assert torch.equal(p.mode(), mean) and torch.equal(p.logvar, logvar)
    g1, g2 = torch.Generator().manual_seed(3), torch.Generator().manual_seed(3)
    s = p.sample(g1)
    assert torch.equal(s, mean + torch.exp(0.5 * logvar) * torch.randn(mean.shape, generator=g2))
    assert p.kl().shape == (2,) and p.nll(s, dims=[1, 2, 3, 4]).shape == (2,)


def test_run_tiles_keeps_order_and_degrades_to_sequential_without_cuda():
    """run_tiles returns results in thunk order; with one stream (or no CUDA device, as here) it is a plain loop."""
    from hunyuanvideo_efficiency_b200.vae.model import run_tiles
    calls = []

    def mk(i):
        def f():
            calls.append(i)
            return torch.full((2,), float(i))
        return f

    for n_streams in (1, 2, 3):
        calls.clear()
        outs = run_tiles([mk(i) for i in range(5)], n_streams)
        assert [int(o[0]) for o in outs] == list(range(5)) and calls == list(range(5))
    assert run_tiles([], 2) == []
