"""GPU parity at BASELINE.json's OWN shapes: the CUDA path (through the C ABI) against the CPU oracle at the sizes the
headline bench runs, not at reduced ones.

What is covered (SURVEY.md §8d shape list; reference loops autoencoder_kl_causal_3d.py:422-469,510-541, vae.py:230-294):
  * the canonical decoder tile 1x16x17x32x32 -> 1x3x65x256x256 and encoder tile 1x3x65x256x256 -> 1x32x17x32x32 of the
    720p split (36 of the 84 sub-model calls per direction), HY config, bf16 and fp16 models;
  * one ragged tile per direction (latent 17x18x16, video 33x144x128: the right / bottom edge tiles);
  * the mid-block attention at L = 17 x 32 x 32 = 17408, D = 512 against the oracle's SDPA restatement;
  * BASELINE config 1 (1x3x17x256x256, fp32, untiled) at <= 1e-4;
  * a 2 x 2 spatial x 2 temporal tile grid at the real tile sizes (65 x 264 x 264 video: tiles of 256 / 72 pixels and
    65 / 17 frames, blends in all three directions) through enable_tiling().
The oracle evaluates, in fp32, the SAME parameters and inputs the 16-bit model holds (rounded to bf16; every bf16 value
is an fp16 value up to fp16's subnormal range, so one oracle run serves both 16-bit models).  Gates are BASELINE.json's:
16-bit <= 2e-2 relative, decode PSNR >= 45 dB, fp32 <= 1e-4.  About 3-4 minutes of CPU oracle on the box's 16 cores.
"""
import functools
import os

import pytest
import torch

from oracle import vae_oracle as O
from oracle import weights as W

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2
PSNR_MIN = 45.0
CFG = W.HY_VAE_CONFIG


def _dev():
    return torch.device("cuda:0")


@functools.lru_cache(maxsize=None)
def _sd_bf16():
    """HY-config parameters rounded to bf16 (what `vae.to(torch.bfloat16)` holds), as fp32 tensors for the oracle."""
    torch.set_num_threads(os.cpu_count() or 1)
    return {k: v.bfloat16().float() for k, v in W.make_state_dict(CFG).items()}


@functools.lru_cache(maxsize=None)
def _model(dtype):
    from hunyuanvideo_efficiency_b200.vae import AutoencoderKLCausal3D
    m = AutoencoderKLCausal3D.from_config(CFG)
    sd = _sd_bf16() if dtype != torch.float32 else W.make_state_dict(CFG)
    m.load_state_dict(sd)
    return m.to(dtype).to(_dev()).eval().requires_grad_(False)


@functools.lru_cache(maxsize=None)
def _oracle_decoder_tile(shape):
    z = W.make_latent(shape).bfloat16().float()
    with torch.no_grad():
        return z, O._dec_tile(_sd_bf16(), CFG, z, None)


@functools.lru_cache(maxsize=None)
def _oracle_encoder_tile(shape):
    x = W.make_video(shape).bfloat16().float()
    with torch.no_grad():
        return x, O._enc_tile(_sd_bf16(), CFG, x, None)


def _check_decoder_tile(shape, dtype):
    z, ref = _oracle_decoder_tile(shape)
    m = _model(dtype)
    m.disable_tiling()
    dec = m.decode(z.to(_dev(), dtype)).sample.float().cpu()
    assert dec.shape == ref.shape
    err, psnr = O.rel_err(ref, dec), O.psnr(ref, dec)
    print(f"decoder tile {shape} {dtype}: rel err {err:.3e}, PSNR {psnr:.1f} dB")
    assert err < BF16_TOL and psnr > PSNR_MIN, (err, psnr)


def _check_encoder_tile(shape, dtype):
    x, ref = _oracle_encoder_tile(shape)
    m = _model(dtype)
    m.disable_tiling()
    mom = m.encode(x.to(_dev(), dtype)).latent_dist.parameters.float().cpu()
    assert mom.shape == ref.shape
    lc = CFG["latent_channels"]
    err_mean, err_all = O.rel_err(ref[:, :lc], mom[:, :lc]), O.rel_err(ref, mom)
    print(f"encoder tile {shape} {dtype}: latent rel err {err_mean:.3e}, moments rel err {err_all:.3e}")
    assert err_mean < BF16_TOL and err_all < BF16_TOL, (err_mean, err_all)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_canonical_decoder_tile_vs_oracle(dtype):
    """post_quant_conv + DecoderCausal3D on 1x16x17x32x32 (autoencoder_kl_causal_3d.py:443-450 at the 720p split): the
    pair halo kernel with hundreds of m-tiles, the kh-trick kernel at 136 m-tiles, the sub-pixel phases, conv_stack, the
    first-frame fold at T = 65 and the fused attention at L = 17408, all at the shapes the bench dispatches."""
    _check_decoder_tile((1, 16, 17, 32, 32), dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_canonical_encoder_tile_vs_oracle(dtype):
    """EncoderCausal3D + quant_conv on 1x3x65x256x256 (autoencoder_kl_causal_3d.py:389-396)."""
    _check_encoder_tile((1, 3, 65, 256, 256), dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_ragged_edge_tiles_vs_oracle(dtype):
    """The bottom-right tiles of the 720p split: latent 17x18x16 (decode) and video 33x144x128 (encode)."""
    _check_decoder_tile((1, 16, 17, 18, 16), dtype)
    _check_encoder_tile((1, 3, 33, 144, 128), dtype)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention_at_L17408_vs_oracle(dtype):
    """hyvae_attn_block_causal at the canonical mid-block shape (17 frames of 32x32, one 512-wide head) against
    O.sdpa_frame_causal (unet_causal_3d_blocks.py:38-46,661) in fp32 on the same 16-bit-rounded q, k, v."""
    from hunyuanvideo_efficiency_b200 import _native as N
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    torch.set_num_threads(os.cpu_count() or 1)
    T, hw, D = 17, 1024, 512
    L = T * hw
    g = torch.Generator().manual_seed(17408)
    q, k, v = (torch.randn(L, D, generator=g).to(dtype) for _ in range(3))
    bv = torch.randn(D, generator=g)
    with torch.no_grad():
        ref = O.sdpa_frame_causal(q.float(), k.float(), v.float(), T, hw, D ** -0.5) + bv
    o = N.attn_block_causal(q.to(_dev()), k.to(_dev()), v.t().contiguous().to(_dev()), bv.to(_dev()), hw, D ** -0.5)
    err = O.rel_err(ref, o.float().cpu())
    print(f"attention L={L} {dtype}: rel err {err:.3e}")
    assert err < (2e-3 if dtype == torch.float16 else 1e-2), err


def test_config1_fp32_vs_oracle():
    """BASELINE config 1: encode+decode of 1x3x17x256x256 in fp32, untiled, HY config; latents and reconstruction
    within 1e-4 relative of the oracle (which is pinned to the unmodified reference by tests/test_oracle_golden.py)."""
    torch.set_num_threads(os.cpu_count() or 1)
    sd = W.make_state_dict(CFG)
    x = W.make_video((1, 3, 17, 256, 256))
    with torch.no_grad():
        ref_dec, ref_mean, _ = O.forward(sd, CFG, x, O.Tiling.from_cfg(CFG))
    m = _model(torch.float32)
    m.disable_tiling()
    dec, post = m(x.to(_dev()), return_dict=False, return_posterior=True)
    e_lat, e_dec = O.rel_err(ref_mean, post.mode().cpu()), O.rel_err(ref_dec, dec.cpu())
    print(f"config 1 fp32: latent rel err {e_lat:.3e}, decode rel err {e_dec:.3e}")
    assert e_lat < FP32_TOL and e_dec < FP32_TOL, (e_lat, e_dec)


@pytest.mark.parametrize("dtype", [torch.bfloat16])
def test_tiled_subgrid_with_blends_vs_oracle(dtype):
    """enable_tiling() on a 65 x 264 x 264 clip: 2 temporal tiles (65 and 17 frames) x 2 x 2 spatial tiles (256 and 72
    pixels; 32 and 9 latent) with blend_v / blend_h / blend_t, first-frame drops and crops, at the HY tile sizes
    (autoencoder_kl_causal_3d.py:362-541).  Encode is compared on the blended latents, decode on the reconstruction of the
    oracle's latents (rounded to the model dtype, as the model receives them)."""
    torch.set_num_threads(os.cpu_count() or 1)
    sd = _sd_bf16()
    tl = O.Tiling.from_cfg(CFG, True, True)
    x = W.make_video((1, 3, 65, 264, 264)).bfloat16().float()
    with torch.no_grad():
        mom = O.encode_moments(sd, CFG, x, tl)
        mean, _ = O.posterior_mean_logvar(mom)
        zin = mean.to(dtype).float()
        ref_dec = O.decode(sd, CFG, zin.clone(), tl)
    m = _model(dtype)
    m.enable_tiling()
    try:
        lat = m.encode(x.to(_dev(), dtype)).latent_dist.mode().float().cpu()
        dec = m.decode(zin.to(_dev(), dtype)).sample.float().cpu()
    finally:
        m.disable_tiling()
    assert lat.shape == mean.shape == (1, 16, 17, 33, 33) and dec.shape == ref_dec.shape == (1, 3, 65, 264, 264)
    e_lat, e_dec, psnr = O.rel_err(mean, lat), O.rel_err(ref_dec, dec), O.psnr(ref_dec, dec)
    print(f"tiled 65x264x264 {dtype}: latent rel err {e_lat:.3e}, decode rel err {e_dec:.3e}, PSNR {psnr:.1f} dB")
    assert e_lat < BF16_TOL and e_dec < BF16_TOL and psnr > PSNR_MIN, (e_lat, e_dec, psnr)
