"""Pin the CPU oracle (oracle/vae_oracle.py) to outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by oracle/make_golden.py from /root/reference)."""
import pytest
import torch

from oracle import vae_oracle as O
from oracle import weights as W

from conftest import load_golden

TOL = 2e-5  # fp32 CPU vs fp32 CPU; only summation order differs


def _check(ref, out, tol=TOL):
    assert ref.shape == out.shape
    assert O.rel_err(ref, out) < tol


def test_ops_conv_stride_variants():
    _, a = load_golden("ops")
    for tag, s in (("s111", (1, 1, 1)), ("s122", (1, 2, 2)), ("s222", (2, 2, 2)), ("s422", (4, 2, 2))):
        _check(a[f"conv_{tag}_y"], O.causal_conv3d(a["x"], a["conv_w"], a["conv_b"], s))
    _check(a["conv1_y"], O.causal_conv3d(a["x"], a["conv1_w"], None))


def test_ops_upsample_first_frame_rule():
    _, a = load_golden("ops")
    assert torch.equal(a["up_u222_y"], O.upsample_nearest_causal(a["x"], (2, 2, 2)))
    assert torch.equal(a["up_u122_y"], O.upsample_nearest_causal(a["x"], (1, 2, 2)))
    assert torch.equal(a["up_u222_T1_y"], O.upsample_nearest_causal(a["x"][:, :, :1], (2, 2, 2)))


def test_ops_mask():
    _, a = load_golden("ops")
    assert torch.equal(a["mask_3_4"], O.frame_causal_mask(3, 4))


def test_ops_resnet_and_midblock():
    _, a = load_golden("ops")
    rsd = {k[len("res_sd."):]: v for k, v in a.items() if k.startswith("res_sd.")}
    _check(a["res_y"], O.resnet_block(rsd, "", a["x"], 32))
    msd = {k[len("mid_sd."):]: v for k, v in a.items() if k.startswith("mid_sd.")}
    _check(a["mid_y"], O.mid_block(msd, "", a["mid_x"], 32))


@pytest.mark.parametrize("name", ["small_untiled", "small_untiled_b2", "small_spatial", "small_temporal",
                                  "small_tiled", "hy_untiled", "small_tops_pool_interp", "small_tops_stride4"])
def test_model_cases(name):
    meta, a = load_golden(name)
    cfg = getattr(W, meta["cfg"])
    sd = W.make_state_dict(cfg, meta["weight_seed"])
    tl = O.Tiling.from_cfg(cfg, meta["spatial"], meta["temporal"])
    x = W.make_video(tuple(meta["shape"]), meta["video_seed"])
    mom = O.encode_moments(sd, cfg, x, tl, meta["t_ops"])
    _check(a["moments"], mom)
    mean, _ = O.posterior_mean_logvar(a["moments"])
    dec = O.decode(sd, cfg, mean.clone(), tl, meta["t_ops"])
    _check(a["dec"], dec)
    assert O.psnr(a["dec"], dec) > 80


def test_spec_has_248_keys():
    assert len(W.state_dict_spec(W.HY_VAE_CONFIG)) == 248
    n = sum(torch.Size(s).numel() for s in W.state_dict_spec(W.HY_VAE_CONFIG).values())
    assert n == 246_478_803  # SURVEY.md §8c [measured]


def test_sdpa_restatement_matches_torch_sdpa_with_the_reference_mask():
    """diffusers' AttnProcessor2_0 (third-party, not vendored) calls torch's F.scaled_dot_product_attention with the additive
    mask of prepare_causal_attention_mask (unet_causal_3d_blocks.py:38-46, call site :661).  The oracle's restatement is
    pinned here against that torch function itself, with the mask built the reference's way (row by row)."""
    import torch
    T, hw, D = 4, 6, 32
    L = T * hw
    g = torch.Generator().manual_seed(3)
    q, k, v = (torch.randn(L, D, generator=g) for _ in range(3))
    mask = torch.full((L, L), float("-inf"))
    for i in range(L):                      # the reference's loop: mask[i, :(i // n_hw + 1) * n_hw] = 0
        mask[i, :(i // hw + 1) * hw] = 0
    ref = torch.nn.functional.scaled_dot_product_attention(q[None, None], k[None, None], v[None, None], attn_mask=mask[None, None])[0, 0]
    out = O.sdpa_frame_causal(q, k, v, T, hw, D ** -0.5)
    assert torch.allclose(out, ref, atol=1e-6, rtol=1e-5)
    assert torch.equal(O.frame_causal_mask(T, hw), mask)


def test_winograd_t_algebra_matches_the_causal_conv():
    """oracle/winograd.py (the F(2,3)-along-T form a round-2 kernel has to follow) == the plain causal conv, in fp64, for even,
    odd and single-frame T; with fp16-rounded transformed operands it stays within 2x of the direct fp16 evaluation."""
    import torch
    from oracle import winograd as WG
    torch.manual_seed(0)
    for T in (1, 2, 5, 8):
        x = torch.randn(2, 6, T, 7, 9, dtype=torch.float64)
        w = torch.randn(5, 6, 3, 3, 3, dtype=torch.float64) / 10
        b = torch.randn(5, dtype=torch.float64)
        ref = O.causal_conv3d(x, w, b)
        assert torch.allclose(WG.causal_conv3d_winograd_t(x, w, b), ref, atol=1e-12, rtol=1e-10)
    x = torch.nn.functional.silu(torch.randn(1, 32, 6, 10, 10))
    w = torch.randn(32, 32, 3, 3, 3) / (27 * 32) ** 0.5
    ref = O.causal_conv3d(x, w, None)
    h = lambda t: t.half().float()
    e_direct = O.rel_err(ref, O.causal_conv3d(h(x), h(w), None))
    e_wino = O.rel_err(ref, WG.causal_conv3d_winograd_t(x, w, None, rnd=h))
    assert e_wino < 2 * e_direct + 1e-4, (e_wino, e_direct)
