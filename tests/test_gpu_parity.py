"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed golden
fixtures of the unmodified reference.  Tolerances are BASELINE.json's: fp32 <= 1e-4 relative,
bf16 <= 2e-2 relative, decode PSNR >= 45 dB."""
import os

import pytest
import torch

from conftest import load_golden
from oracle import vae_oracle as O
from oracle import weights as W

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2
PSNR_MIN = 45.0


def _dev():
    return torch.device("cuda:0")


def _N():
    from hunyuanvideo_efficiency_b200 import _native as N
    return N


def _vol(x, dtype=None, pad=(0, 0, 0)):
    return _N().Vol.from_ncthw(x.to(_dev()), dtype=dtype, pad=pad)


def _pack(w, dtype):
    co, ci, k = w.shape[0], w.shape[1], w.shape[2]
    return w.permute(2, 3, 4, 0, 1).reshape(k ** 3, co, ci).to(_dev(), dtype).contiguous()


# ----------------------------------------------------------------------------------------- kernels
def test_layout_roundtrip_and_halo():
    N = _N()
    x = torch.randn(2, 5, 3, 6, 7)
    v = _vol(x, pad=(2, 1, 1))
    assert torch.equal(v.to_ncthw().cpu(), x)
    ref = torch.nn.functional.pad(x, (1, 1, 1, 1, 2, 0), mode="replicate").permute(0, 2, 3, 4, 1)
    assert torch.equal(v.t.cpu(), ref)
    sl = x[:, :, 1:, 2:5, ::2]  # strided source view
    assert torch.equal(N.Vol.from_ncthw(sl.to(_dev())[:, :, :, :, :]).to_ncthw().cpu(), sl)


@pytest.mark.parametrize("tag,stride", [("s111", (1, 1, 1)), ("s122", (1, 2, 2)), ("s222", (2, 2, 2)), ("s422", (4, 2, 2))])
def test_direct_conv_fp32_vs_reference(tag, stride):
    N = _N()
    _, a = load_golden("ops")
    y = N.conv3d_direct(_vol(a["x"]), _pack(a["conv_w"], torch.float32), a["conv_b"].to(_dev()), 3, stride, 64)
    assert O.rel_err(a[f"conv_{tag}_y"], y.to_ncthw().cpu()) < 1e-5


def test_direct_conv_k1_nobias_and_upsample_fold():
    N = _N()
    _, a = load_golden("ops")
    y = N.conv3d_direct(_vol(a["x"]), _pack(a["conv1_w"], torch.float32), None, 1, (1, 1, 1), 48)
    assert O.rel_err(a["conv1_y"], y.to_ncthw().cpu()) < 1e-5
    for up in ((2, 2, 2), (1, 2, 2)):
        ref = O.causal_conv3d(O.upsample_nearest_causal(a["x"], up), a["conv_w"], a["conv_b"])
        y = N.conv3d_direct(_vol(a["x"]), _pack(a["conv_w"], torch.float32), a["conv_b"].to(_dev()), 3, (1, 1, 1), 64, up=up)
        assert O.rel_err(ref, y.to_ncthw().cpu()) < 1e-5


def test_pad_upsample_matches_reference():
    N = _N()
    _, a = load_golden("ops")
    for key, up, x in (("up_u222_y", (2, 2, 2), a["x"]), ("up_u122_y", (1, 2, 2), a["x"]), ("up_u222_T1_y", (2, 2, 2), a["x"][:, :, :1])):
        y = N.pad_upsample(_vol(x), up, pad=(2, 1, 1))
        assert torch.equal(y.to_ncthw().cpu(), a[key])
        ref = torch.nn.functional.pad(a[key], (1, 1, 1, 1, 2, 0), mode="replicate").permute(0, 2, 3, 4, 1)
        assert torch.equal(y.t.cpu(), ref)


@pytest.mark.parametrize("pad", [(2, 1, 1), (1, 1, 1), (2, 0, 0)])
def test_halo_fill_replicates_the_interior(pad):
    """hyvae_halo_fill on a volume whose interior was written in place == F.pad(replicate) of that interior; and a conv that
    writes into the interior of a padded output followed by halo_fill == the same conv followed by the pad pass."""
    N = _N()
    x = torch.randn(2, 16, 3, 6, 7).half()
    v = N.Vol(2, 3, 6, 7, 16, torch.float16, _dev(), pad)
    v.t.fill_(float("nan"))
    v.interior().copy_(x.permute(0, 2, 3, 4, 1).to(_dev()))
    N.halo_fill(v)
    ref = torch.nn.functional.pad(x.float(), (pad[2], pad[2], pad[1], pad[1], pad[0], 0), mode="replicate").permute(0, 2, 3, 4, 1).half()
    assert torch.equal(v.t.cpu(), ref)
    if pad == (2, 1, 1) and N.device_supports_tc():
        from hunyuanvideo_efficiency_b200.vae.blocks import CausalConv3d
        torch.manual_seed(1)
        conv = CausalConv3d(64, 128, 3).half().to(_dev())
        xin = _vol(torch.randn(1, 64, 4, 20, 24), torch.float16, pad=(2, 1, 1))
        y_pad = conv.forward_vol(xin, out_pad=pad)
        y_ref = N.pad_upsample(conv.forward_vol(xin), (1, 1, 1), pad)
        assert y_pad.pad == pad and torch.equal(y_pad.t, y_ref.t)


def test_groupnorm_silu_fp32_and_halo():
    N = _N()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 64, 3, 9, 7, generator=g) * 3 + 1.5
    gamma, beta = torch.randn(64, generator=g), torch.randn(64, generator=g)
    ref = torch.nn.functional.silu(O.group_norm(x, gamma, beta, 32))
    y = N.groupnorm(_vol(x), gamma.to(_dev()), beta.to(_dev()), 32, 1e-6, True, pad=(2, 1, 1), round_like_ref=False)
    assert O.rel_err(ref, y.to_ncthw().cpu()) < 1e-5
    refp = torch.nn.functional.pad(y.to_ncthw().cpu(), (1, 1, 1, 1, 2, 0), mode="replicate").permute(0, 2, 3, 4, 1)
    assert torch.equal(y.t.cpu(), refp)


def test_softmax_frame_causal():
    N = _N()
    T, hw = 3, 20
    L = T * hw
    S = torch.randn(2, L, L)
    ref = torch.softmax(S * 0.125 + O.frame_causal_mask(T, hw)[None], dim=-1)
    P = N.softmax_frame_causal(S.to(_dev()).contiguous(), hw, 0.125, torch.float32)
    assert torch.allclose(P.cpu(), ref, atol=1e-6, rtol=1e-5)
    assert torch.count_nonzero(P.cpu()[0, 0, hw:]) == 0


@pytest.mark.parametrize("T,hw,D,dtype,spread", [
    (3, 64, 512, torch.float16, False),    # L = 192: tiles cross frame boundaries, ragged last tile
    (5, 72, 512, torch.bfloat16, False),   # L = 360
    (9, 288, 512, torch.float16, True),    # the smallest 720p latent tile (9 x 18 x 16); score range forces O rescaling
    (4, 40, 256, torch.float16, True),
    (3, 48, 128, torch.bfloat16, False),
    (2, 1024, 512, torch.float16, False),  # n_hw = 8 key blocks per frame, every tile inside one frame
])
def test_attn_fused_matches_oracle(T, hw, D, dtype, spread):
    """hyvae_attn_block_causal (flash-style tcgen05 kernel) against the oracle's SDPA restatement, evaluated in fp32 on the
    same 16-bit-rounded q, k, v.  P is rounded to the operand type once: 2^-11 (fp16) / 2^-8 (bf16) relative per term."""
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    L = T * hw
    g = torch.Generator().manual_seed(7)
    q = torch.randn(L, D, generator=g)
    k = torch.randn(L, D, generator=g)
    v = torch.randn(L, D, generator=g)
    if spread:  # later key blocks score far higher: the running maximum jumps by more than 2^8 -> lazy rescale path
        k[128:] *= 6.0
        k[384:] *= 3.0
    bv = torch.randn(D, generator=g)
    q, k, v = q.to(dtype), k.to(dtype), v.to(dtype)
    ref = O.sdpa_frame_causal(q.float(), k.float(), v.float(), T, hw, D ** -0.5) + bv
    o = N.attn_block_causal(q.to(_dev()), k.to(_dev()), v.t().contiguous().to(_dev()), bv.to(_dev()), hw, D ** -0.5)
    err = O.rel_err(ref, o.float().cpu())
    assert err < (2e-3 if dtype == torch.float16 else 1e-2), err
    # rows of the first frame must not see any later key: recompute them from the first frame alone
    ref0 = O.sdpa_frame_causal(q[:hw].float(), k[:hw].float(), v[:hw].float(), 1, hw, D ** -0.5) + bv
    assert O.rel_err(ref0, o[:hw].float().cpu()) < (2e-3 if dtype == torch.float16 else 1e-2)


def test_attn_fused_full_size_matches_unfused_chain_and_is_reproducible():
    """BASELINE's canonical mid-block shape (L = 17 x 32 x 32, D = 512), too large for the CPU oracle in test time: the
    fused kernel against the GEMM -> softmax -> GEMM chain of the same library (itself oracle-checked at small sizes),
    bit-reproducibility, and the causal property (changing later frames' keys / values leaves earlier rows unchanged)."""
    from hunyuanvideo_efficiency_b200.vae.blocks import _gemm_nt
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    T, hw, D = 17, 1024, 512
    L = T * hw
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(L, D, device=_dev(), generator=g).half()
    k = torch.randn(L, D, device=_dev(), generator=g).half()
    vt = torch.randn(D, L, device=_dev(), generator=g).half()
    bv = torch.randn(D, device=_dev(), generator=g)
    o1 = N.attn_block_causal(q, k, vt, bv, hw, D ** -0.5)
    o2 = N.attn_block_causal(q, k, vt, bv, hw, D ** -0.5)
    assert torch.equal(o1, o2)
    qv = N.Vol(1, 1, 1, L, D, torch.float16, _dev(), tensor=q.reshape(1, 1, 1, L, D))
    s = _gemm_nt(qv, k, None, L, out_dtype=torch.float32)
    p = N.softmax_frame_causal(s.t.reshape(1, L, L), hw, D ** -0.5, torch.float16)
    pv = N.Vol(1, 1, 1, L, L, torch.float16, _dev(), tensor=p.reshape(1, 1, 1, L, L))
    ou = _gemm_nt(pv, vt, bv, D).t.reshape(L, D)
    assert O.rel_err(ou.float().cpu(), o1.float().cpu()) < 1e-3
    k2, vt2 = k.clone(), vt.clone()
    k2[9 * hw:] = 0
    vt2[:, 9 * hw:] = 1
    o3 = N.attn_block_causal(q, k2, vt2, bv, hw, D ** -0.5)
    assert torch.equal(o3[:9 * hw], o1[:9 * hw]) and not torch.equal(o3[9 * hw:], o1[9 * hw:])


def test_attn_fused_rejects_unsupported_shapes():
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    q = torch.zeros(90, 64, dtype=torch.float16, device=_dev())
    with pytest.raises(N.HyvaeUnsupported):
        N.attn_block_causal(q, q, q.t().contiguous(), None, 30, 0.125)


def test_midblock_fused_attention_matches_unfused_schedule(monkeypatch):
    """The mid block with the fused attention kernel against the GEMM -> softmax -> GEMM schedule of the same build, and
    both against the oracle evaluated in fp32 on the same parameters."""
    from hunyuanvideo_efficiency_b200.vae.blocks import UNetMidBlockCausal3D
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    torch.manual_seed(3)
    mb = UNetMidBlockCausal3D(in_channels=256, temb_channels=None, resnet_groups=32, attention_head_dim=256).half()
    x = torch.randn(1, 256, 3, 12, 16).half()
    sd = {k_: v_.float() for k_, v_ in mb.state_dict().items()}
    ref = O.mid_block(sd, "", x.float(), 32)
    mbd = mb.to(_dev())
    y_f = mbd(x.to(_dev())).float().cpu()
    monkeypatch.setenv("HYVAE_ATTN_UNFUSED", "1")
    y_u = mbd(x.to(_dev())).float().cpu()
    assert O.rel_err(ref, y_f) < 5e-3 and O.rel_err(ref, y_u) < 5e-3
    assert O.rel_err(y_u, y_f) < 2e-3


def test_temporal_pool_and_interp():
    N = _N()
    x = torch.randn(1, 16, 7, 4, 5)
    for k, s in ((3, 2), (2, 2), (2, 1)):
        assert torch.allclose(N.avgpool_t(_vol(x), k, s).to_ncthw().cpu(), O.t_avg_pool(x, k, s), atol=1e-6)
    for sc in (2, 3, 1.5):
        assert torch.equal(N.interp_t_nearest(_vol(x), sc).to_ncthw().cpu(), O.t_interp(x, sc, "nearest"))
        assert torch.equal(N.interp_t(_vol(x), sc, "nearest-exact").to_ncthw().cpu(), O.t_interp(x, sc, "nearest-exact"))
        for mode in ("trilinear", "area"):   # the other modes F.interpolate takes for a 5-D tensor (unet_causal_3d_blocks.py:889-897)
            assert torch.allclose(N.interp_t(_vol(x), sc, mode).to_ncthw().cpu(), O.t_interp(x, sc, mode), atol=1e-5), (sc, mode)


def test_peer_copy_is_a_stream_ordered_device_copy():
    """hyvae_peer_copy (the tile push of vae/tile_parallel.py; on one GPU the 'peer' is the same device)."""
    N = _N()
    src = torch.randn(3, 5, 7, device=_dev()).to(torch.bfloat16)
    dst = torch.zeros_like(src)
    N.peer_copy(dst, src)
    torch.cuda.synchronize()
    assert torch.equal(dst, src)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_blend_bit_exact(dtype):
    from hunyuanvideo_efficiency_b200.vae import AutoencoderKLCausal3D
    m = AutoencoderKLCausal3D.from_config(W.SMALL_CONFIG)
    g = torch.Generator().manual_seed(5)
    a = torch.randn(1, 3, 4, 10, 12, generator=g).to(dtype)
    for fn, ofn, b_shape in ((m.blend_v, O.blend_v, (1, 3, 4, 6, 12)), (m.blend_h, O.blend_h, (1, 3, 4, 10, 5)),
                             (m.blend_t, O.blend_t, (1, 3, 7, 10, 12))):
        b = torch.randn(b_shape, generator=g).to(dtype)
        ref = ofn(a.clone(), b.clone(), 4)
        out = fn(a.to(_dev()), b.to(_dev()).clone(), 4)
        assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_blend_vec8_path_and_assembly_chain_bit_exact(dtype):
    """The 16-byte blend kernel (rows / windows / extents that are multiples of 8, as in every production split) and the
    raster-order assembly built on it, against the oracle's blend loops in the reference's order of operations
    (autoencoder_kl_causal_3d.py:403-412 spatial, :529-541 temporal)."""
    from hunyuanvideo_efficiency_b200.vae import AutoencoderKLCausal3D
    m = AutoencoderKLCausal3D.from_config(W.SMALL_CONFIG)
    g = torch.Generator().manual_seed(11)
    a = torch.randn(1, 3, 6, 24, 32, generator=g).to(dtype)
    for fn, ofn, b_shape, e in ((m.blend_v, O.blend_v, (1, 3, 6, 16, 32), 8), (m.blend_h, O.blend_h, (1, 3, 6, 24, 24), 16),
                                (m.blend_t, O.blend_t, (1, 3, 9, 24, 32), 4)):
        b = torch.randn(b_shape, generator=g).to(dtype)
        ref = ofn(a.clone(), b.clone(), e)
        out = fn(a.to(_dev()), b.to(_dev()).clone(), e)
        assert torch.equal(out.cpu(), ref)
    # 3 x 3 grid of tiles (ragged last row / column), extent 8, limit 24
    hs, ws = [32, 32, 16], [32, 32, 24]
    tiles = [[torch.randn(1, 3, 5, h, w, generator=g).to(dtype) for w in ws] for h in hs]
    ref_rows = []
    work = [[t.clone() for t in row] for row in tiles]
    for i, row in enumerate(work):
        res = []
        for j, t in enumerate(row):
            if i > 0:
                t = O.blend_v(work[i - 1][j], t, 8)
            if j > 0:
                t = O.blend_h(row[j - 1], t, 8)
            res.append(t[:, :, :, :24, :24])
        ref_rows.append(torch.cat(res, dim=-1))
    ref = torch.cat(ref_rows, dim=-2)
    out = m._assemble_spatial([[t.to(_dev()) for t in row] for row in tiles], 8, 24)
    assert torch.equal(out.cpu(), ref)
    # temporal chain: three tiles, the later ones drop their first frame, extent 4, limit 12
    tt = [torch.randn(1, 3, n, 16, 24, generator=g).to(dtype) for n in (17, 17, 9)]
    row = [tt[0].clone()] + [t[:, :, 1:].clone() for t in tt[1:]]
    res = []
    for i, t in enumerate(row):
        if i > 0:
            t = O.blend_t(row[i - 1], t, 4)
            res.append(t[:, :, :12])
        else:
            res.append(t[:, :, :13])
    ref = torch.cat(res, dim=2)
    out = m._assemble_temporal([(tt[0].to(_dev()), 0)] + [(t.to(_dev()), 1) for t in tt[1:]], 4, 12)
    assert torch.equal(out.cpu(), ref)


# ----------------------------------------------------------------------------------------- tcgen05 conv
def _rand_case(B, Cin, Cout, T, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, T, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5
    b = torch.randn(Cout, generator=g)
    return x, w, b


@pytest.mark.parametrize("B,Cin,Cout,T,H,W,stride", [
    (1, 64, 64, 3, 8, 16, (1, 1, 1)),      # exactly one 8x16 tile, BN=64
    (1, 128, 128, 4, 16, 32, (1, 1, 1)),   # BN=128
    (2, 128, 256, 3, 18, 20, (1, 1, 1)),   # ragged H/W, batch 2, BN=256
    (1, 512, 512, 2, 8, 16, (1, 1, 1)),    # two n-tiles, deep K
    (1, 192, 96, 2, 9, 33, (1, 1, 1)),     # Cin/Cout not multiples of 64 -> TMA zero fill on both operands
    (1, 128, 128, 5, 16, 32, (1, 2, 2)),   # DownsampleCausal3D strides
    (1, 128, 128, 5, 16, 32, (2, 2, 2)),
    (1, 64, 32, 9, 12, 10, (4, 2, 2)),     # run-time T-stride 4 (t-ops), BN=32
])
def test_tc_conv_matches_direct_bf16(B, Cin, Cout, T, H, W, stride):
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    x, w, b = _rand_case(B, Cin, Cout, T, H, W, 11)
    xb, wb = x.bfloat16(), w.bfloat16()
    ref = O.causal_conv3d(xb.float(), wb.float(), b, stride)  # same bf16-rounded operands, fp32 math
    xv = _vol(xb, pad=(2, 1, 1))
    y = N.conv3d_tc(xv, _pack(wb, torch.bfloat16), b.to(_dev()), 3, stride, Cout, out_dtype=torch.float32)
    assert O.rel_err(ref, y.to_ncthw().cpu()) < 2e-5   # fp32 accumulation of exact bf16 products
    y16 = N.conv3d_tc(xv, _pack(wb, torch.bfloat16), b.to(_dev()), 3, stride, Cout)
    assert torch.equal(y16.to_ncthw().cpu(), y.to_ncthw().cpu().bfloat16()) or O.rel_err(ref, y16.to_ncthw().float().cpu()) < 4e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,T,H,W", [(2, 5, 20, 24), (1, 3, 33, 17), (1, 1, 16, 16)])
def test_tc_conv_in_kw_packed_matches_reference(dtype, B, T, H, W):
    """conv_in (3 -> 128) on the kw-packed operand (hyvae_ncthw_to_vol_kw3 + variant bit 9: 9 taps of K = 16) against the
    fp32 reference conv on the same 16-bit-rounded operands and against the unpacked thin kernel; the packed layout itself
    is checked element by element."""
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    from hunyuanvideo_efficiency_b200.vae.blocks import CausalConv3d
    g = torch.Generator().manual_seed(T * 100 + W)
    conv = CausalConv3d(3, 128, 3).to(_dev())
    conv.emit_gn_groups = 32
    assert conv.wants_kw_pack(dtype)
    x = torch.randn(B, 3, T, H, W, generator=g).to(dtype)
    pad, ch = conv.input_layout(dtype)
    vp = N.Vol.from_ncthw(x.to(_dev()), pad=pad, channels=ch, kw_pack=True)
    xi = vp.interior().float().cpu()                                   # [B][T][H][W][16]
    xl = x.float().permute(0, 2, 3, 4, 1)                              # [B][T][H][W][3]
    idx = torch.arange(W)
    for kw in range(3):
        assert torch.equal(xi[..., 3 * kw:3 * kw + 3], xl[:, :, :, (idx + kw - 1).clamp(0, W - 1)])
    assert torch.count_nonzero(xi[..., 9:]) == 0
    y_p = conv.forward_vol(vp)
    y_u = conv.forward_vol(N.Vol.from_ncthw(x.to(_dev()), pad=pad, channels=ch))
    ref = O.causal_conv3d(x.float(), conv.conv.weight.detach().to(dtype).float().cpu(), conv.conv.bias.detach().float().cpu())
    tol = 2e-3 if dtype == torch.float16 else 8e-3                      # one rounding of the stored output
    assert O.rel_err(ref, y_p.to_ncthw().float().cpu()) < tol and O.rel_err(ref, y_u.to_ncthw().float().cpu()) < tol
    assert torch.allclose(y_p.gn_sums, y_u.gn_sums, rtol=1e-3, atol=0.5)


def test_tc_thin_layers_and_fused_gn_stats():
    """conv_in (3 -> C, channels zero-padded to 8), conv_out (C -> 3, Cout padded to 8) and the GroupNorm partial
    statistics emitted by the conv epilogue."""
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    from hunyuanvideo_efficiency_b200.vae.blocks import CausalConv3d
    g = torch.Generator().manual_seed(9)
    cin3 = CausalConv3d(3, 128, 3).to(_dev())
    x = torch.randn(2, 3, 5, 20, 24, generator=g).half()
    cin3.emit_gn_groups = 32
    pad, ch = cin3.input_layout(torch.float16)
    assert (pad, ch) == ((2, 1, 1), 16)     # thin-Cin form: 3 channels stored as 16 (one K = 16 slice per 32-byte row)
    y = cin3.forward_vol(N.Vol.from_ncthw(x.to(_dev()), pad=pad, channels=ch))
    ref = O.causal_conv3d(x.float(), cin3.conv.weight.detach().half().float().cpu(), cin3.conv.bias.detach().float().cpu())
    assert O.rel_err(ref, y.to_ncthw().float().cpu()) < 2e-3
    # fused statistics vs the stats kernel on the stored tensor
    assert y.gn_sums is not None and y.gn_groups == 32
    n = 4 * 5 * 20 * 24
    s_ref = ref.reshape(2, 32, -1)
    assert torch.allclose(y.gn_sums[:, :, 0].cpu() / n, s_ref.mean(-1).double(), atol=2e-4)
    assert torch.allclose(y.gn_sums[:, :, 1].cpu() / n, (s_ref ** 2).mean(-1).double(), rtol=2e-3)
    cout3 = CausalConv3d(128, 3, 3).to(_dev())
    h = torch.randn(1, 128, 3, 18, 20, generator=g).half()
    y3 = cout3.forward_vol(N.Vol.from_ncthw(h.to(_dev())))
    assert y3.c_valid == 3 and y3.to_ncthw().shape == (1, 3, 3, 18, 20)
    ref3 = O.causal_conv3d(h.float(), cout3.conv.weight.detach().half().float().cpu(), cout3.conv.bias.detach().float().cpu())
    assert O.rel_err(ref3, y3.to_ncthw().float().cpu()) < 2e-3


@pytest.mark.parametrize("Cin,Cout,T,H,W,stride", [(128, 128, 9, 64, 64, (1, 1, 1)), (256, 256, 5, 40, 72, (1, 1, 1)),
                                                   (512, 512, 3, 18, 32, (1, 1, 1)), (128, 256, 9, 36, 64, (2, 2, 2)),
                                                   (128, 64, 3, 8, 24, (1, 1, 1))])
def test_tc_cta_pair_kernel_matches_single_cta_kernel(Cin, Cout, T, H, W, stride):
    """variant 4 (cta_group::2 pair kernel, many tiles per pair, odd tile counts) vs variant 2 (1-CTA kernel)."""
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    x, w, b = _rand_case(1, Cin, Cout, T, H, W, 21)
    xv = _vol(x.half(), pad=(2, 1, 1))
    wp = _pack(w.half(), torch.float16)
    r = N.Vol(1, *N.conv_out_dims(T, H, W, stride), Cout, torch.float16, _dev())
    r.t.normal_()
    # variant 4 = CTA-pair kernel with the kh trick where it applies (auto would pick the halo kernel for Cout <= 128)
    ys = [N.conv3d_tc(xv, wp, b.to(_dev()), 3, stride, Cout, residual=r, variant=v, gn_groups=32) for v in (4, 2, 4, 3)]
    assert torch.equal(ys[0].t, ys[2].t)                                    # reproducible
    for other in (ys[1], ys[3]):                                            # pair+kh-trick vs 1-CTA vs pair without kh-trick
        assert O.rel_err(other.t.float().cpu(), ys[0].t.float().cpu()) < 1e-3
        assert torch.allclose(ys[0].gn_sums, other.gn_sums, rtol=1e-5, atol=1e-3)
    ref = O.causal_conv3d(x.half().float(), w.half().float(), b, stride) + r.to_ncthw().float().cpu()
    assert O.rel_err(ref, ys[0].to_ncthw().float().cpu()) < 2e-3


@pytest.mark.parametrize("B,Cin,Cout,T,H,W,res,gn", [
    (2, 128, 128, 3, 40, 24, True, 32),    # two 16x16 groups in W (second m-tile of the last group half empty), batch 2
    (2, 64, 128, 2, 21, 19, False, 32),    # ragged H and W, Cin = one chunk
    (1, 256, 128, 3, 32, 48, True, 32),    # four K chunks (decoder 256 -> 128)
    (1, 128, 8, 3, 40, 40, False, 0),      # conv_out: Cout padded 3 -> 8 (halo kernel BN = 32; auto = stacked-tap kernel)
    (2, 128, 8, 2, 21, 35, False, 0),      # conv_out, ragged tile edges, batch 2
    (1, 8, 128, 3, 40, 40, False, 32),     # conv_in: Cin padded 3 -> 8, one K = 16 slice per tap
    (1, 64, 64, 3, 32, 32, True, 32),      # BN = 64
    (1, 128, 128, 5, 6, 5, True, 32),      # smaller than one tile
])
def test_tc_halo_kernel_matches_oracle_and_single_cta_kernel(B, Cin, Cout, T, H, W, res, gn):
    """conv_halo.cu (variant 5 / auto for stride-1 3x3x3, Cout <= 128): full 2-D halo stage read at unaligned
    row offsets, TMA-store epilogue with the residual tile fetched by TMA, halving-tree GroupNorm statistics."""
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    x, w, b = _rand_case(B, Cin, Cout, T, H, W, 31)
    xv = _vol(x.half(), pad=(2, 1, 1))
    wp = _pack(w.half(), torch.float16)
    r = None
    if res:
        r = N.Vol(B, T, H, W, Cout, torch.float16, _dev())
        r.t.normal_()
    ys = [N.conv3d_tc(xv, wp, b.to(_dev()), 3, (1, 1, 1), Cout, residual=r, variant=v, gn_groups=gn, round_like_ref=False)
          for v in (5, 2, 5, 0)]
    assert torch.equal(ys[0].t, ys[2].t)                                        # reproducible
    if Cout == 8 and Cin % 64 == 0:   # auto = the stacked-tap kernel (conv_stack.cu): other summation order, also reproducible
        y7 = N.conv3d_tc(xv, wp, b.to(_dev()), 3, (1, 1, 1), Cout, variant=7, round_like_ref=False)
        assert torch.equal(ys[3].t, y7.t) and O.rel_err(ys[0].t.float().cpu(), ys[3].t.float().cpu()) < 1e-3
    else:
        assert torch.equal(ys[0].t, ys[3].t)                                    # auto picks the halo kernel
    if Cout > 64:   # CTA-pair form (cta_group::2, half of the weight tile per CTA): same MMA order, same result
        yp = N.conv3d_tc(xv, wp, b.to(_dev()), 3, (1, 1, 1), Cout, residual=r, variant=6, gn_groups=gn, round_like_ref=False)
        assert torch.equal(ys[0].t, yp.t)
        if gn:
            assert torch.allclose(ys[0].gn_sums, yp.gn_sums, rtol=1e-9, atol=1e-6)
    assert O.rel_err(ys[1].t.float().cpu(), ys[0].t.float().cpu()) < 1e-3
    if gn:
        assert torch.equal(ys[0].gn_sums, ys[2].gn_sums)
        assert torch.allclose(ys[0].gn_sums, ys[1].gn_sums, rtol=1e-5, atol=5e-2)   # different fp32 summation trees
    ref = O.causal_conv3d(x.half().float(), w.half().float(), b)
    if res:
        ref = ref + r.to_ncthw().float().cpu()
    assert O.rel_err(ref, ys[0].to_ncthw().float().cpu()) < 2e-3
    if gn:   # statistics of the fp32 result, per (batch, group)
        n = (Cout // gn) * T * H * W
        g_ref = ref.reshape(B, gn, -1).double()
        assert torch.allclose(ys[0].gn_sums[:, :, 0].cpu() / n, g_ref.mean(-1), atol=5e-4)
        assert torch.allclose(ys[0].gn_sums[:, :, 1].cpu() / n, (g_ref ** 2).mean(-1), rtol=2e-3)


@pytest.mark.parametrize("B,Cin,Cout,T,H,W,up", [
    (1, 128, 128, 3, 10, 12, (2, 2, 2)),   # ragged low-res tile (10 x 12 < 16 x 8 grid), BN = 128
    (2, 64, 256, 2, 16, 8, (1, 2, 2)),     # T not upsampled: 3 x 2 x 2 taps, batch 2, BN = 256
    (1, 256, 256, 1, 20, 24, (2, 2, 2)),   # single frame: the odd temporal phase is empty
    (1, 512, 512, 5, 18, 32, (2, 2, 2)),   # two n-tiles, deep K, the decoder's ragged 18 x 32 tile
])
def test_tc_upsample_phase_decomposition_matches_reference(B, Cin, Cout, T, H, W, up):
    """UpsampleCausal3D (unet_causal_3d_blocks.py:130-183) as sub-pixel phase convolutions over the low-res tensor
    (hyvae_conv3d_upphase_tc) vs the reference order of operations (upsample, then 27-tap conv) in fp32."""
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    from hunyuanvideo_efficiency_b200.vae.blocks import UpsampleCausal3D
    torch.manual_seed(5)
    m = UpsampleCausal3D(Cin, use_conv=True, out_channels=Cout, upsample_factor=up).to(_dev())
    x = torch.randn(B, Cin, T, H, W).half()
    w, b = m.conv.conv.weight.detach().half().float().cpu(), m.conv.conv.bias.detach().float().cpu()
    ref = O.causal_conv3d(O.upsample_nearest_causal(x.float(), up), w, b)
    assert m.phase_decomposition
    y = m.forward_vol(_vol(x))
    assert O.rel_err(ref, y.to_ncthw().float().cpu()) < 2e-3
    m.phase_decomposition = False
    y2 = m.forward_vol(_vol(x))
    assert O.rel_err(y2.to_ncthw().float().cpu(), y.to_ncthw().float().cpu()) < 2e-3
    assert y.gn_sums is not None and y2.gn_sums is not None
    n = (Cout // 32) * ref.shape[2] * ref.shape[3] * ref.shape[4]
    g_ref = ref.reshape(B, 32, -1).double()
    assert torch.allclose(y.gn_sums[:, :, 0].cpu() / n, g_ref.mean(-1), atol=1e-3)
    assert torch.allclose(y.gn_sums[:, :, 1].cpu() / n, (g_ref ** 2).mean(-1), rtol=3e-3)


@pytest.mark.parametrize("Cin,Cout,T,H,W,variant", [
    (128, 128, 5, 32, 32, 5),     # halo kernel, 1 CTA
    (128, 128, 4, 24, 40, 6),     # halo kernel, CTA pair; 6 groups per frame
    (64, 128, 3, 16, 40, 6),      # CTA pair with 3 groups per frame: pairs straddle frames 0/1 and 1/2 -> unfolded for those
    (128, 64, 2, 20, 28, 5),      # BN = 64: three taps per weight stage
    (256, 256, 5, 32, 32, 4),     # kh-trick pair kernel
    (128, 512, 3, 16, 24, 4),     # kh-trick, 3 m-tiles per frame (odd): straddling pairs
    (256, 256, 1, 16, 16, 4),     # a single frame: every tile is class 0
])
def test_tc_first_frame_temporal_fold_matches_unfolded(Cin, Cout, T, H, W, variant):
    """Folded first-frame taps (variant bit 8, 45 weight slices) against the plain 27-tap schedule of the same kernel and
    against the fp32 reference conv: frames >= 2 are bit-identical, frames 0 and 1 agree to fp16 rounding of the folded
    weights, residual and GroupNorm statistics included."""
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    from hunyuanvideo_efficiency_b200.vae.blocks import CausalConv3d
    torch.manual_seed(Cin + T)
    conv = CausalConv3d(Cin, Cout, 3).to(_dev())
    x = torch.randn(1, Cin, T, H, W)
    r = torch.randn(1, Cout, T, H, W)
    xv, rv = _vol(x, torch.float16, pad=(2, 1, 1)), _vol(r, torch.float16)
    w27, b = conv.conv.packed(torch.float16, pad8=True)
    w27 = w27.clone()
    w45, _ = conv.conv.packed(torch.float16, pad8=True, tfold=True)
    assert w45.shape[0] == 45 and torch.equal(w45[:27], w27)
    y0 = N.conv3d_tc(xv, w27, b, 3, (1, 1, 1), Cout, residual=rv, gn_groups=32, variant=variant)
    y1 = N.conv3d_tc(xv, w45, b, 3, (1, 1, 1), Cout, residual=rv, gn_groups=32, variant=variant | N.VARIANT_TFOLD)
    a0, a1 = y0.to_ncthw().float().cpu(), y1.to_ncthw().float().cpu()
    assert torch.equal(a0[:, :, 2:], a1[:, :, 2:])
    ref = torch.nn.functional.conv3d(torch.nn.functional.pad(x.half().float(), (1, 1, 1, 1, 2, 0), mode="replicate"),
                                     conv.conv.weight.detach().half().float().cpu(), conv.conv.bias.detach().float().cpu()) + r.half().float()
    assert O.rel_err(ref, a0) < 2e-3 and O.rel_err(ref, a1) < 2e-3
    assert O.rel_err(a0[:, :, :2], a1[:, :, :2]) < 2e-3
    assert torch.allclose(y0.gn_sums, y1.gn_sums, rtol=2e-3, atol=1.0)


def _wino_planes_ref(f):
    """Planes of conv_wino.cu's operand from f = SiLU(GroupNorm(x)) [B][C][T][H][W] (oracle/winograd.py with the pairs
    starting at frame 1: plane 0 = f[0]; pair p = frames (2p+1, 2p+2); an even T ends with three planes)."""
    T = f.shape[2]
    xp = torch.cat([f[:, :, :1], f[:, :, :1], f], 2)
    out = [f[:, :, 0]]
    for p in range((T - 1) // 2):
        d = [xp[:, :, 2 * p + 1 + i] for i in range(4)]
        out += [d[0] - d[2], d[1] + d[2], d[2] - d[1], d[1] - d[3]]
    if T % 2 == 0:
        d = [xp[:, :, T - 1 + i] for i in range(3)]
        out += [d[0] - d[2], d[1] + d[2], d[2] - d[1]]
    return torch.stack(out, 2)


@pytest.mark.parametrize("B,Cin,Cout,T,H,W,res,gn", [
    (1, 64, 128, 1, 16, 16, False, False),     # a single frame: only the folded first-frame GEMM
    (1, 64, 128, 2, 9, 21, True, False),       # even T: frame 0 + the 3-GEMM tail, ragged tile, odd m-tile count
    (1, 128, 128, 9, 40, 24, True, True),      # four pairs, residual + GroupNorm statistics of the output
    (2, 64, 256, 4, 20, 18, False, True),      # batch 2, even T with a pair, two n-tiles
    (1, 256, 512, 3, 16, 24, True, True),      # four n-tiles, an odd number of m-tiles (half-empty CTA pair)
    (1, 512, 512, 5, 32, 32, True, True),      # the mid-block geometry
])
def test_tc_winograd_t_conv_matches_oracle(B, Cin, Cout, T, H, W, res, gn):
    """GroupNorm + SiLU written as Winograd-T planes (hyvae_groupnorm_apply_wino) followed by hyvae_conv3d_causal_wino, against
    the oracle's SiLU(GroupNorm) -> causal_conv3d (+ residual) in fp32 on the same fp16-rounded parameters: <= 2e-3 relative
    (the direct fp16 kernel is at ~3.6e-4; the transformed operands add one rounding).  Also the plane volume itself."""
    import torch.nn.functional as F
    from hunyuanvideo_efficiency_b200.vae.blocks import CausalConv3d, _GroupNorm
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    g = torch.Generator().manual_seed(B * 1000 + Cin + T)
    x = torch.randn(B, Cin, T, H, W, generator=g).half()
    conv = CausalConv3d(Cin, Cout, 3).to(_dev())
    norm = _GroupNorm(32, Cin).to(_dev())
    with torch.no_grad():
        norm.weight.copy_(1 + 0.1 * torch.randn(Cin, generator=g))
        norm.bias.copy_(0.1 * torch.randn(Cin, generator=g))
    conv.emit_gn_groups = 32 if gn else 0
    assert conv.wants_wino(torch.float16)
    f = F.silu(O.group_norm(x.float(), norm.weight.detach().cpu(), norm.bias.detach().cpu(), 32))
    ref = O.causal_conv3d(f, conv.conv.weight.detach().cpu().half().float(), conv.conv.bias.detach().cpu().float())
    r = torch.randn(B, Cout, T, H, W, generator=g).half() if res else None
    if res:
        ref = ref + r.float()
    pl = norm.forward_vol(_vol(x), True, wino=True)
    assert pl.wino_T == T and pl.T == N.wino_planes(T) and pl.pad == (0, 1, 1)
    got = pl.t[:, :, 1:-1, 1:-1, :].permute(0, 4, 1, 2, 3).float().cpu()
    assert (got - _wino_planes_ref(f)).abs().max() < 4e-3                      # fp16 rounding of values up to ~8, tanh.approx SiLU
    assert torch.equal(pl.t[:, :, 0], pl.t[:, :, 1]) and torch.equal(pl.t[:, :, :, -1], pl.t[:, :, :, -2])   # replicate halo
    y = conv.forward_vol(pl, residual=_vol(r) if res else None)
    out = y.to_ncthw().float().cpu()
    assert O.rel_err(ref, out) < 2e-3, O.rel_err(ref, out)
    if gn:
        o64 = out.double().reshape(B, 32, Cout // 32, -1)
        sq_ref = (o64 * o64).sum((2, 3))
        assert ((y.gn_sums.cpu()[..., 1] - sq_ref).abs() / sq_ref).max() < 2e-3
        assert (y.gn_sums.cpu()[..., 0] - o64.sum((2, 3))).abs().max() < 2e-3 * o64.abs().sum((2, 3)).max()
    y2 = conv.forward_vol(norm.forward_vol(_vol(x), True, wino=True), residual=_vol(r) if res else None)
    assert torch.equal(y.t, y2.t)                                              # static schedule: bit-reproducible
    if gn:   # the in-kernel finalize left the partial buffer and its ticket zeroed: the second launch finds the same sums
        assert torch.equal(y.gn_sums, y2.gn_sums)


def test_tc_winograd_t_matches_plain_path_in_a_resnet_block(monkeypatch):
    """A ResnetBlockCausal3D with and without the Winograd-T convs (HYVAE_WINO=0): same block, same input, results within
    the fp16 rounding of the transformed operands; out_pad (the halo'd destination a sampler conv asks for) is honoured."""
    from hunyuanvideo_efficiency_b200.vae.blocks import ResnetBlockCausal3D
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    blk = ResnetBlockCausal3D(in_channels=128, out_channels=128, temb_channels=None).to(_dev())
    x = torch.randn(1, 128, 7, 24, 40, generator=torch.Generator().manual_seed(3)).half()
    n0 = N.launch_count()
    y = blk.forward_vol(_vol(x), out_pad=(2, 1, 1))
    assert y.pad == (2, 1, 1) and torch.equal(y.t[:, 0], y.t[:, 2]) and torch.equal(y.t[:, :, 0], y.t[:, :, 1])
    yw = y.to_ncthw().float().cpu()
    monkeypatch.setenv("HYVAE_WINO", "0")
    yp = blk.forward_vol(_vol(x)).to_ncthw().float().cpu()
    assert N.launch_count() > n0 and O.rel_err(yp, yw) < 1e-3, O.rel_err(yp, yw)


def test_tc_gemm_k1_residual_and_fp16():
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    g = torch.Generator().manual_seed(2)
    for dt in (torch.bfloat16, torch.float16):
        x = torch.randn(1, 256, 2, 10, 12, generator=g).to(dt)
        w = (torch.randn(128, 256, 1, 1, 1, generator=g) / 16).to(dt)
        b = torch.randn(128, generator=g)
        r = torch.randn(1, 128, 2, 10, 12, generator=g).to(dt)
        conv = (O.causal_conv3d(x.float(), w.float(), b)).to(dt).float()   # reference rounds before the add
        ref = (conv + r.float()).to(dt)
        y = N.conv3d_tc(_vol(x), _pack(w, dt), b.to(_dev()), 1, (1, 1, 1), 128, residual=_vol(r))
        d = (y.to_ncthw().cpu().float() - ref.float()).abs().max().item()
        assert d <= 0.0625, d   # at most one 16-bit ulp at |v| ~ 8


def test_tc_big_gemm_matrix_mode():
    """Attention-shaped GEMM: S = Q K^T with L rows, fp32 out, tile = 1 x 128 rows."""
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    g = torch.Generator().manual_seed(4)
    L, Cn = 1000, 128   # L not a multiple of 128 or 64
    q = torch.randn(L, Cn, generator=g).bfloat16()
    k = torch.randn(L, Cn, generator=g).bfloat16()
    qv = N.Vol(1, 1, 1, L, Cn, torch.bfloat16, _dev(), tensor=q.to(_dev()).reshape(1, 1, 1, L, Cn).contiguous())
    s = N.conv3d_tc(qv, k.to(_dev()).contiguous(), None, 1, (1, 1, 1), L, out_dtype=torch.float32)
    ref = q.float() @ k.float().T
    assert O.rel_err(ref, s.t.reshape(L, L).cpu()) < 2e-5


# ----------------------------------------------------------------------------------------- blocks vs reference
def test_resnet_and_midblock_fp32_vs_reference():
    from hunyuanvideo_efficiency_b200.vae.blocks import ResnetBlockCausal3D, UNetMidBlockCausal3D
    _, a = load_golden("ops")
    r = ResnetBlockCausal3D(in_channels=32, out_channels=64, temb_channels=None, groups=32, eps=1e-6)
    r.load_state_dict({k[len("res_sd."):]: v for k, v in a.items() if k.startswith("res_sd.")})
    y = r.to(_dev())(a["x"].to(_dev()))
    assert O.rel_err(a["res_y"], y.cpu()) < FP32_TOL
    mb = UNetMidBlockCausal3D(in_channels=64, temb_channels=None, resnet_groups=32, attention_head_dim=64)
    mb.load_state_dict({k[len("mid_sd."):]: v for k, v in a.items() if k.startswith("mid_sd.")})
    y = mb.to(_dev())(a["mid_x"].to(_dev()))
    assert O.rel_err(a["mid_y"], y.cpu()) < FP32_TOL


@pytest.mark.parametrize("cin,cout,T,H,W", [(256, 128, 3, 24, 40), (64, 128, 2, 9, 21), (128, 256, 3, 20, 24), (512, 256, 2, 18, 32)])
def test_resnet_block_fp16_fused_shortcut_vs_reference(cin, cout, T, H, W):
    """ResnetBlockCausal3D with Cin != Cout (unet_causal_3d_blocks.py:338-348,407-415): the 1x1x1 conv_shortcut runs
    as extra K chunks inside conv2 (hyvae_conv3d_causal_tc_shortcut).  Checked against the fp32 oracle evaluated on the
    fp16-rounded parameters, and against the unfused schedule."""
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    from hunyuanvideo_efficiency_b200.vae.blocks import ResnetBlockCausal3D
    torch.manual_seed(3)
    r = ResnetBlockCausal3D(in_channels=cin, out_channels=cout, temb_channels=None, groups=32, eps=1e-6)
    with torch.no_grad():
        for n_, p_ in r.named_parameters():   # non-trivial norm affine parameters
            if "norm" in n_:
                p_.copy_(torch.randn_like(p_) * 0.3 + (1.0 if n_.endswith("weight") else 0.0))
    r = r.half().to(_dev())
    x = torch.randn(2, cin, T, H, W).half()
    sd = {"b." + k: v.float().cpu() for k, v in r.state_dict().items()}
    ref = O.resnet_block(sd, "b.", x.float(), 32)
    n0 = N.launch_count()
    y = r(x.to(_dev()))
    fused_launches = N.launch_count() - n0
    os.environ["HYVAE_FUSE_SHORTCUT"] = "0"
    try:
        n0 = N.launch_count()
        y2 = r(x.to(_dev()))
        unfused_launches = N.launch_count() - n0
    finally:
        os.environ.pop("HYVAE_FUSE_SHORTCUT")
    assert fused_launches <= unfused_launches              # the k=1 shortcut launch is gone when the tile shape allows
    if cout <= 128:   # ... and with it (Winograd-T path: statistics finished inside the conv) the GroupNorm finalize launches
        assert fused_launches < unfused_launches
    assert O.rel_err(ref, y.float().cpu()) < 3e-3
    assert O.rel_err(y2.float().cpu(), y.float().cpu()) < 2e-3


# ----------------------------------------------------------------------------------------- whole model
def _build(cfg, dtype, t_ops=None):
    from hunyuanvideo_efficiency_b200.vae import AutoencoderKLCausal3D, _apply_t_ops_config_to_vae
    m = AutoencoderKLCausal3D.from_config(cfg)
    m.load_state_dict(W.make_state_dict(cfg))
    m = m.to(dtype).to(_dev()).eval().requires_grad_(False)
    if t_ops is not None:
        _apply_t_ops_config_to_vae(m, t_ops)
    return m


MODEL_CASES = ["small_untiled", "small_untiled_b2", "small_spatial", "small_temporal", "small_tiled", "hy_untiled",
               "small_tops_pool_interp", "small_tops_stride4"]


@pytest.mark.parametrize("name", MODEL_CASES)
def test_model_fp32_vs_reference_golden(name):
    meta, a = load_golden(name)
    m = _build(getattr(W, meta["cfg"]), torch.float32, meta["t_ops"])
    m.enable_spatial_tiling(meta["spatial"])
    m.enable_temporal_tiling(meta["temporal"])
    x = W.make_video(tuple(meta["shape"]), meta["video_seed"]).to(_dev())
    post = m.encode(x).latent_dist
    assert post.parameters.shape == a["moments"].shape
    assert O.rel_err(a["moments"], post.parameters.cpu()) < FP32_TOL
    mean, _ = O.posterior_mean_logvar(a["moments"])
    dec = m.decode(mean.to(_dev())).sample
    assert dec.shape == a["dec"].shape
    assert O.rel_err(a["dec"], dec.cpu()) < FP32_TOL
    assert O.psnr(a["dec"], dec.cpu()) > 80


def _exact_eval_of_16bit_model(meta, dtype):
    """fp32 oracle evaluated on the SAME parameters and input the 16-bit model holds (weights and video
    rounded to `dtype`, then exact arithmetic): the value a 16-bit implementation is approximating."""
    cfg = getattr(W, meta["cfg"])
    sd = {k: v.to(dtype).float() for k, v in W.make_state_dict(cfg, meta["weight_seed"]).items()}
    tl = O.Tiling.from_cfg(cfg, meta["spatial"], meta["temporal"])
    x = W.make_video(tuple(meta["shape"]), meta["video_seed"]).to(dtype).float()
    mom = O.encode_moments(sd, cfg, x, tl, meta["t_ops"])
    mean, _ = O.posterior_mean_logvar(mom)
    zin = mean.to(dtype).float()
    return mean, zin, O.decode(sd, cfg, zin.clone(), tl, meta["t_ops"])


@pytest.mark.parametrize("name", ["small_untiled", "small_tiled", "hy_untiled"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_model_16bit_vs_reference(name, dtype):
    """BASELINE.json tolerances for 16-bit models: latents within 2e-2 relative, decode PSNR >= 45 dB.
    Checked against (a) the exact evaluation of the same 16-bit parameters and (b) the fp32 golden output of
    the unmodified reference (which additionally contains the 16-bit weight rounding itself)."""
    meta, a = load_golden(name)
    m = _build(getattr(W, meta["cfg"]), dtype)
    m.enable_spatial_tiling(meta["spatial"])
    m.enable_temporal_tiling(meta["temporal"])
    x = W.make_video(tuple(meta["shape"]), meta["video_seed"]).to(_dev(), dtype)
    mean_exact, zin, dec_exact = _exact_eval_of_16bit_model(meta, dtype)
    mean = m.encode(x).latent_dist.mode().float().cpu()
    dec = m.decode(zin.to(_dev(), dtype)).sample.float().cpu()
    assert O.rel_err(mean_exact, mean) < BF16_TOL
    assert O.rel_err(dec_exact, dec) < BF16_TOL and O.psnr(dec_exact, dec) > PSNR_MIN
    # (b) vs the fp32-weight golden: this also contains the error of rounding the WEIGHTS to 16 bits, which belongs
    # to the model the caller built (vae.to(dtype)), not to the implementation; fp16 still meets the tolerance,
    # bf16 weights alone cost ~1e-2 / ~40 dB on random-init weights (tests/dev/precision_study.py).
    mean_gold, _ = O.posterior_mean_logvar(a["moments"])
    dec_g = m.decode(mean_gold.to(_dev(), dtype)).sample.float().cpu()
    if dtype == torch.float16:
        assert O.rel_err(mean_gold, mean) < BF16_TOL
        assert O.rel_err(a["dec"], dec_g) < BF16_TOL and O.psnr(a["dec"], dec_g) > PSNR_MIN
    else:
        assert O.rel_err(mean_gold, mean) < 3 * BF16_TOL and O.psnr(a["dec"], dec_g) > 35.0


def test_pure_bf16_mode_is_no_worse_than_the_reference_in_bf16():
    """`bf16_compute = "bf16"` keeps every operand in bf16 like the reference.  The reference's own all-bf16 CPU
    run of this case deviates from its fp32 run by 2.7e-2 (latent) / 7.2e-2 (decode), PSNR 38.0 dB (measured with
    oracle/_refshim, DESIGN.md "Precision"); the CUDA path must not be worse than that by more than 25 %."""
    meta, a = load_golden("small_untiled")
    m = _build(getattr(W, meta["cfg"]), torch.bfloat16)
    m.bf16_compute = "bf16"
    x = W.make_video(tuple(meta["shape"]), meta["video_seed"]).to(_dev(), torch.bfloat16)
    mean_gold, _ = O.posterior_mean_logvar(a["moments"])
    mean = m.encode(x).latent_dist.mode().float().cpu()
    dec = m.decode(mean_gold.to(_dev(), torch.bfloat16)).sample.float().cpu()
    assert O.rel_err(mean_gold, mean) < 1.25 * 2.7e-2
    assert O.rel_err(a["dec"], dec) < 1.25 * 7.2e-2


def test_groupnorm_and_decoder_are_bit_reproducible():
    N = _N()
    x = N.Vol(1, 9, 64, 64, 128, torch.bfloat16, _dev())
    x.t.normal_()
    g, b = torch.ones(128, device=_dev()), torch.zeros(128, device=_dev())
    ys = [N.groupnorm(x, g, b, 32, 1e-6, True, pad=(2, 1, 1)).t.clone() for _ in range(4)]
    assert all(torch.equal(ys[0], y) for y in ys[1:])
    cfg = dict(W.SMALL_CONFIG, block_out_channels=[64, 128, 256, 256])
    m = _build(cfg, torch.bfloat16)
    z = W.make_latent((1, 16, 5, 12, 10)).to(_dev(), torch.bfloat16)
    outs = [m.decode(z).sample.clone() for _ in range(4)]
    assert all(torch.equal(outs[0], o) for o in outs[1:])


def test_tc_model_path_matches_direct_path_bf16():
    """Same bf16 model through the tcgen05 convs and through the CUDA-core convs (HYVAE_FORCE_DIRECT)."""
    N = _N()
    if not N.device_supports_tc():
        pytest.skip("needs sm_100")
    cfg = dict(W.SMALL_CONFIG, block_out_channels=[64, 128, 256, 256])
    m = _build(cfg, torch.bfloat16)
    m.enable_tiling()
    x = W.make_video((1, 3, 21, 40, 48)).to(_dev(), torch.bfloat16)
    sd = {k: v.bfloat16().float() for k, v in W.make_state_dict(cfg).items()}
    ref_dec, ref_mean, _ = O.forward(sd, cfg, x.float().cpu(), O.Tiling.from_cfg(cfg, True, True))
    n0 = N.launch_count()
    dec_tc, post_tc = m(x, return_dict=False, return_posterior=True)
    assert N.launch_count() > n0
    os.environ["HYVAE_FORCE_DIRECT"] = "1"
    try:
        dec_d, post_d = m(x, return_dict=False, return_posterior=True)
    finally:
        os.environ.pop("HYVAE_FORCE_DIRECT")
    for dec, post in ((dec_tc, post_tc), (dec_d, post_d)):
        assert O.rel_err(ref_mean, post.mode().float().cpu()) < BF16_TOL
        assert O.psnr(ref_dec, dec.float().cpu()) > PSNR_MIN
    assert O.rel_err(dec_d.float().cpu(), dec_tc.float().cpu()) < BF16_TOL


def test_tile_streams_do_not_change_results():
    """Tiles dealt over 1, 2 and 3 CUDA streams (run_tiles): bit-identical latents and reconstruction, repeatedly."""
    cfg = dict(W.SMALL_CONFIG, block_out_channels=[64, 128, 256, 256])
    m = _build(cfg, torch.bfloat16)
    m.enable_tiling()
    x = W.make_video((1, 3, 21, 72, 88)).to(_dev(), torch.bfloat16)
    outs = []
    for n in (1, 2, 3, 2):
        m.tile_streams = n
        dec, post = m(x, return_dict=False, return_posterior=True)
        outs.append((dec.clone(), post.mode().clone()))
    torch.cuda.synchronize()
    for dec, lat in outs[1:]:
        assert torch.equal(dec, outs[0][0]) and torch.equal(lat, outs[0][1])


def test_full_size_tile_properties_bf16():
    """HY config at one canonical decoder tile shape (BASELINE config 2's unit): linearity-free properties
    that do not need the oracle at full size: determinism, and tiled == untiled when one tile covers the input."""
    N = _N()
    m = _build(W.HY_VAE_CONFIG, torch.bfloat16)
    z = W.make_latent((1, 16, 5, 32, 32)).to(_dev(), torch.bfloat16)
    d1 = m.decode(z).sample
    m.enable_tiling()
    d2 = m.decode(z).sample
    assert d1.shape == (1, 3, 17, 256, 256) and torch.equal(d1, d2)
    assert torch.isfinite(d1.float()).all()


def test_pipeline_tail_matches_oracle():
    """pipeline_hunyuan_video.py:1060-1092: latents / scaling_factor -> tiled decode -> (x / 2 + 0.5).clamp(0, 1) -> float,
    against the ORACLE evaluating that chain (fp16 division of the latents as the pipeline does it, fp32 decode of the
    fp16-rounded parameters, the image ops in fp32).  decode_to_image folds the scaling into post_quant_conv's weights and
    the image ops into the last assembly kernel's epilogue; both rewrites are also checked against the unfused kernels."""
    from hunyuanvideo_efficiency_b200.pipeline_tail import decode_latents
    N = _N()
    cfg = dict(W.SMALL_CONFIG, block_out_channels=[64, 128, 256, 256], scaling_factor=0.476986)
    m = _build(cfg, torch.float16)
    z = W.make_latent((1, 16, 5, 12, 10)).half()
    out = decode_latents(m, z.to(_dev()))
    assert out.dtype == torch.float32 and out.device.type == "cpu" and out.shape == (1, 3, 17, 96, 80)
    sd = {k: v.half().float() for k, v in W.make_state_dict(cfg).items()}
    zin = (z / cfg["scaling_factor"]).float()                         # the pipeline divides in the latents' dtype (:1069)
    ref = O.decode(sd, cfg, zin, O.Tiling.from_cfg(cfg, True, True))
    ref = (ref / 2 + 0.5).clamp(0, 1)
    assert O.rel_err(ref, out) < BF16_TOL and O.psnr(ref, out, data_range=1.0) > PSNR_MIN, (O.rel_err(ref, out), O.psnr(ref, out, 1.0))
    # the fused epilogue is bit-identical to decode() followed by the stand-alone post-process kernel ...
    m.enable_tiling()
    plain = m.decode(z.to(_dev()), return_dict=False)[0]
    assert torch.equal(m.decode_to_image(z.to(_dev())), N.image_postprocess(plain))
    assert torch.equal(N.image_postprocess(plain).cpu(), (plain / 2 + 0.5).clamp(0, 1).float().cpu())
    for t in (plain.float(), plain[..., :5, :3].contiguous(), plain.bfloat16()):   # fp32 input, odd sizes, bf16: same kernel
        assert torch.equal(N.image_postprocess(t).cpu(), (t / 2 + 0.5).clamp(0, 1).float().cpu())
    # ... and folding the scale into the weights moves one rounding point only
    unfolded = N.image_postprocess(m.decode(z.to(_dev()) / cfg["scaling_factor"], return_dict=False)[0]).cpu()
    assert O.psnr(unfolded, out, data_range=1.0) > 55.0
    m.disable_tiling()
    assert torch.equal(m.decode_to_image(z.to(_dev())), N.image_postprocess(m.decode(z.to(_dev()), return_dict=False)[0]))   # untiled path
    img = decode_latents(m, z[:, :, 0].to(_dev()), to_cpu=False)      # 4-D latents: one frame, temporal dim squeezed
    assert img.shape == (1, 3, 96, 80) and img.is_cuda


def test_fp16_range_guard_reruns_overflowing_tiles_in_bf16():
    """A bf16 model computes with fp16 operands by default (range 65 504).  With conv_in scaled so that its output exceeds
    that range the reference's bf16 evaluation is still fine; the guard must notice the non-finite tiles and re-run them with
    bf16 operands, landing where `bf16_compute = "bf16"` lands (and within the pure-bf16 tolerance of the exact evaluation)."""
    cfg = dict(W.SMALL_CONFIG, block_out_channels=[64, 128, 256, 256])
    sd = W.make_state_dict(cfg)
    for k in ("encoder.conv_in.conv.weight", "encoder.conv_in.conv.bias", "decoder.conv_in.conv.weight", "decoder.conv_in.conv.bias"):
        sd[k] = sd[k] * 1.0e6
    from hunyuanvideo_efficiency_b200.vae import AutoencoderKLCausal3D
    m = AutoencoderKLCausal3D.from_config(cfg)
    m.load_state_dict(sd)
    m = m.to(torch.bfloat16).to(_dev()).eval().requires_grad_(False)
    m.enable_spatial_tiling()
    x = W.make_video((1, 3, 9, 48, 40)).bfloat16()
    sd_b = {k: v.bfloat16().float() for k, v in sd.items()}
    tl = O.Tiling.from_cfg(cfg, True, False)
    ref_mom = O.encode_moments(sd_b, cfg, x.float(), tl)
    ref_mean, _ = O.posterior_mean_logvar(ref_mom)
    zin = ref_mean.bfloat16()
    ref_dec = O.decode(sd_b, cfg, zin.float(), tl)
    m.fp16_range_guard = False
    assert not torch.isfinite(m.encode(x.to(_dev())).latent_dist.mode().float()).all()     # fp16 operands alone overflow
    m.fp16_range_guard = True
    n0 = m.range_guard_reruns
    lat = m.encode(x.to(_dev())).latent_dist.mode()
    dec = m.decode(zin.to(_dev())).sample
    assert m.range_guard_reruns - n0 == 2 * 4                         # every tile of the 2 x 2 grid, both directions
    assert torch.isfinite(lat.float()).all() and torch.isfinite(dec.float()).all()
    m.bf16_compute = "bf16"
    assert torch.equal(lat, m.encode(x.to(_dev())).latent_dist.mode()) and torch.equal(dec, m.decode(zin.to(_dev())).sample)
    assert O.rel_err(ref_mean, lat.float().cpu()) < 1.25 * 2.7e-2 and O.rel_err(ref_dec, dec.float().cpu()) < 1.25 * 7.2e-2
    # a model that stays in range pays one flag read and no re-run
    m2 = _build(cfg, torch.bfloat16)
    m2.enable_spatial_tiling()
    m2.encode(x.to(_dev()))
    assert m2.range_guard_reruns == 0


def test_clip_driver_on_gpu(tmp_path):
    """hunyuanvideo_efficiency_b200/infer.py with the real model: pinned H2D on a copy stream, round trip, async D2H + save."""
    from hunyuanvideo_efficiency_b200 import infer as I
    cfg = dict(W.SMALL_CONFIG, block_out_channels=[64, 128, 256, 256])
    m = _build(cfg, torch.float16)
    d, out = os.path.join(str(tmp_path), "in"), os.path.join(str(tmp_path), "out")
    os.makedirs(d)
    clips = [W.make_video((3, 9, 32, 40), seed) for seed in (1, 2, 3)]
    for i, c in enumerate(clips):
        torch.save(c, os.path.join(d, f"c{i}.pt"))
    done = I.run_clips(I.roundtrip(m), d, out, device=_dev(), in_dtype=torch.float16)
    assert len(done) == 3
    for i, c in enumerate(clips):
        y = torch.load(os.path.join(out, f"c{i}.pt"))
        ref = m(c[None].to(_dev(), torch.float16), return_dict=False, return_posterior=True)[0].cpu().float()
        assert y.dtype == torch.float32 and torch.equal(y, ref)


# ----------------------------------------------------------------------------------------- metrics of the t-ops sweeps
def test_video_quantisation_matches_save_videos_grid_rule():
    from oracle import metrics_oracle as MO
    N = _N()
    g = torch.Generator().manual_seed(11)
    v = (torch.rand(3, 4, 21, 35, generator=g) * 2.6 - 1.3)      # values outside [-1, 1] exercise the clamp
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        vd = v.to(dt)
        ref = MO.video_to_frames_u8(vd.float().numpy(), True)
        out = N.video_to_frames_u8(vd.to(_dev()), True).cpu().numpy()
        assert (out == ref).all()
    sl = v[:, 1:, ::2, 3:30]                                       # strided view, no rescale
    ref = MO.video_to_frames_u8(sl.numpy(), False)
    assert (N.video_to_frames_u8(sl.to(_dev())[:, :, :, :], False).cpu().numpy() == ref).all()
    big = v.to(_dev())[:, :, ::2, 3:30]
    assert (N.video_to_frames_u8(big, True).cpu().numpy() == MO.video_to_frames_u8(big.cpu().numpy(), True)).all()


@pytest.mark.parametrize("T,H,W,C", [(3, 37, 53, 3), (2, 240, 432, 3), (2, 7, 7, 3), (2, 16, 40, 1)])
def test_frame_psnr_ssim_match_oracle(T, H, W, C):
    from oracle import metrics_oracle as MO
    from hunyuanvideo_efficiency_b200 import metrics as M
    rng = __import__("numpy").random.default_rng(T * 1000 + H)
    a = rng.integers(0, 256, (T, H, W, C), dtype="uint8")
    b = (a.astype("int64") + rng.integers(-25, 26, a.shape)).clip(0, 255).astype("uint8")
    b[1] = a[1]                                                    # identical frame: PSNR 100, SSIM 1
    ps, ss = M.frame_metrics(torch.from_numpy(a).to(_dev()), torch.from_numpy(b).to(_dev()))
    for i in range(T):
        assert ps[i] == pytest.approx(MO.psnr_frame(a[i], b[i]), rel=1e-12)
        assert ss[i] == pytest.approx(MO.ssim_frame(a[i], b[i]), abs=1e-10)
    flat = a.copy(); flat[0] = 9                                   # constant original frame -> SSIM 1 (compute_metrics.py:39-40)
    _, ss = M.frame_metrics(torch.from_numpy(flat).to(_dev()), torch.from_numpy(b).to(_dev()))
    assert ss[0] == 1.0
    # zip semantics: the shorter stack decides
    ps2, _ = M.frame_metrics(torch.from_numpy(a).to(_dev()), torch.from_numpy(b[:1]).to(_dev()))
    assert len(ps2) == 1 and ps2[0] == ps[0]


def test_sweep_config_roundtrip_and_scores(tmp_path):
    """One pool experiment and one stride experiment through sweep.run_config on a small fp16 model: the scores equal the
    oracle's metrics of (input, our reconstruction), T changes as the hooks dictate, and the module is restored."""
    from oracle import metrics_oracle as MO
    from hunyuanvideo_efficiency_b200 import sweep as S
    m = _build(W.SMALL_CONFIG, torch.float16)
    x = W.make_video((3, 9, 40, 48))
    torch.save(x, tmp_path / "clip0.pt")
    base = S.default_t_ops_config()
    plain = m(x[None].to(_dev(), torch.float16), return_dict=False)[0].float().cpu()
    for cfg in (None, S.enumerate_pool_configs(base)[0][1], S.enumerate_stride_configs(base)[0][1]):
        res = S.run_config(m, cfg, str(tmp_path), ["clip0.pt"], _dev(), torch.float16, save_dir=str(tmp_path / "rec"))
        rec = torch.load(tmp_path / "rec" / "clip0.pt")
        ref = MO.compare_videos([(MO.video_to_frames_u8(x.numpy()), MO.video_to_frames_u8(rec[0].numpy()))])
        assert res["PSNR"] == pytest.approx(ref["PSNR"], rel=1e-9) and res["SSIM"] == pytest.approx(ref["SSIM"], abs=1e-9)
        if cfg is None:
            assert torch.equal(rec, plain)
    after = m(x[None].to(_dev(), torch.float16), return_dict=False)[0].float().cpu()
    assert torch.equal(after, plain)                                # hooks and strides restored
