/* hyvae.h — C ABI of the B200-native HunyuanVideo 3D causal VAE kernels (libhyvae.so).
 *
 * The reference (c976237222/HunyuanVideo_efficiency) is pure Python: its VAE has NO native / FFI
 * boundary — every device op is reached through torch.nn.functional.  This header
 * therefore defines the boundary a maintainer would bind; each entry point names the reference call
 * site (file:line under /root/reference) whose ATen/cuDNN call chain it replaces.  The binding is
 * ctypes (hunyuanvideo_efficiency_b200/_native.py); INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types; every function returns 0 on success or a
 *    negative hyvae_status, and hyvae_last_error() returns a thread-local message.
 *  - no allocation inside: the caller owns every buffer, including workspaces.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *  - activations are CHANNELS-LAST volumes described by hyvae_vol (below).
 */
#ifndef HYVAE_H_
#define HYVAE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HYVAE_VERSION 121 /* 0.1.2: Winograd-T convs, GroupNorm finalize folded into the convs, blend post-process epilogue */

typedef enum { HYVAE_OK = 0, HYVAE_EINVAL = -1, HYVAE_ECUDA = -2, HYVAE_EUNSUPPORTED = -3 } hyvae_status;
typedef enum { HYVAE_BF16 = 0, HYVAE_F32 = 1, HYVAE_F16 = 2 } hyvae_dtype;

/* A channels-last activation volume in HBM: element (b,t,h,w,c) of the LOGICAL tensor lives at
 *   data[(((b*(T+pt) + t+pt)*(H+2*ph) + h+ph)*(W+2*pw) + w+pw)*C + c].
 * pt/ph/pw is a physical halo holding the reference's replicate padding already materialised
 * (F.pad(x,(pw,pw,ph,ph,pt,0),'replicate'), unet_causal_3d_blocks.py:68,74): pt frames IN FRONT of T
 * only (causal), ph rows / pw columns on both sides.  Producers that are given a padded destination
 * fill the halo; the tensor-core conv reads it through TMA. */
typedef struct {
  void* data;
  int32_t dtype; /* hyvae_dtype */
  int32_t B, T, H, W, C;
  int32_t pt, ph, pw;
} hyvae_vol;

int hyvae_version(void);
const char* hyvae_last_error(void);
/* 1 if the current device can run the tcgen05 path (compute capability 10.x). */
int hyvae_device_supports_tc(void);

/* ---- layout: torch NCTHW <-> channels-last volume ---------------------------------------------
 * Replaces the implicit NCDHW layout of every reference tensor; `dst` halo is filled by replication. */
/* src_strides: element strides of the (possibly sliced) source view in (B,C,T,H,W) order.  src_C <= dst->C:
 * extra destination channels are zero (3 -> 8 channel padding for the tensor-core conv_in).  dst_C <= src->C:
 * only the first dst_C channels are written out (8 -> 3 after the tensor-core conv_out). */
int hyvae_ncthw_to_vol(const void* src, int32_t src_dtype, int32_t src_C, const int64_t* src_strides, const hyvae_vol* dst,
                       void* stream);
int hyvae_vol_to_ncthw(const hyvae_vol* src, void* dst, int32_t dst_dtype, int32_t dst_C, void* stream);
/* conv_in operand with the kw taps packed along the channel axis: dst channel kw*src_C + c of voxel (t,h,w) = source
 * channel c of voxel (t,h,clamp(w+kw-1)), other channels zero (3*src_C <= dst->C, dst->C % 8 == 0).  With it the 3x3x3
 * conv_in (vae.py:118-121) is a 9-tap (kt,kh) conv: hyvae_conv3d_causal_tc with variant bit 9 and w = [9][Cout][16]. */
int hyvae_ncthw_to_vol_kw3(const void* src, int32_t src_dtype, int32_t src_C, const int64_t* src_strides, const hyvae_vol* dst,
                           void* stream);

/* ---- CausalConv3d ------------------------------------------------------------------------------
 * Replaces F.pad(replicate)+nn.Conv3d, unet_causal_3d_blocks.py:73-75 (k=3 or k=1; stride from
 * DownsampleCausal3D :223-225, mutable at run time :741), with the residual add of
 * ResnetBlockCausal3D.forward :415 and the nearest upsample of UpsampleCausal3D.forward :152-171
 * optionally folded in.
 *   w:        [k*k*k][Cout][Cin] in x's dtype (tap-major, Cin contiguous); tap = (kt*k+kh)*k+kw.
 *   bias:     Cout fp32 or NULL.   residual: volume shaped like y, or NULL (y = conv + bias + residual).
 *   up_*:     1 or 2 — the conv input is the nearest-upsampled x (first frame not upsampled in T).
 *   round_like_ref: 1 = round conv+bias to the activation dtype before adding the residual (what
 *             the reference's two separate kernels do in bf16/fp16).
 *   y:        same dtype as x, or HYVAE_F32 (used for the attention scores S = Q K^T).
 * _direct: CUDA-core implicit GEMM, any shape/dtype; replicate padding by index clamping.
 * _tc:     tcgen05/TMEM implicit GEMM fed by TMA; needs bf16/f16, up_*==1 and
 *          x carrying the halo (pt,ph,pw) = (k-1,k/2,k/2).  Cin%8==0 and Cout%8==0 suffice (TMA zero-fills the
 *          rest); gn_partials != NULL additionally emits GroupNorm partial statistics of y from the epilogue.
 *          variant: 0 = automatic kernel choice (low byte 1..7 force a kernel, tests only).  Bit 8 (0x100), stride-1
 *          3x3x3 only: `w` holds 45 tap slices — the 27 above, then W[kt=0]+W[1]+W[2] (9 slices, (kh,kw) order) and
 *          W[0]+W[1] (9 slices).  Output frames 0 and 1 read frame 0 under three / two of their frame taps (causal
 *          replicate padding :68,74), so the kernels that support it run them with one / two folded taps: 1/T fewer MACs.
 *          Bit 9 (0x200): x is a kw-packed thin volume (hyvae_ncthw_to_vol_kw3) and `w` is [9 = kt*3+kh][Cout][16]. */
int hyvae_conv3d_causal_direct(const hyvae_vol* x, const void* w, const float* bias, const hyvae_vol* residual,
                               const hyvae_vol* y, int32_t k, int32_t st, int32_t sh, int32_t sw,
                               int32_t up_t, int32_t up_h, int32_t up_w, int32_t round_like_ref, void* stream);
int hyvae_conv3d_causal_tc(const hyvae_vol* x, const void* w, const float* bias, const hyvae_vol* residual,
                           const hyvae_vol* y, int32_t k, int32_t st, int32_t sh, int32_t sw,
                           int32_t round_like_ref, int32_t variant, double* gn_partials, int32_t gn_groups, void* stream);
/* conv2 of a ResnetBlockCausal3D whose skip path is a 1x1x1 conv_shortcut (unet_causal_3d_blocks.py:338-348,407-415):
 *   y = conv3x3x3(x; w) + conv1x1x1(sc_x; sc_w) + bias,   bias = conv2.bias + conv_shortcut.bias (summed by the caller).
 * x carries the halo (2,1,1); sc_x is the block input (any halo) with y's extent; sc_w: [1][Cout][Csc] in x's dtype.
 * Stride 1, Cout > 64, 16-bit output (halo kernel for Cout <= 128, kh-trick pair kernel above); shapes those kernels do
 * not take return HYVAE_EUNSUPPORTED and the caller runs the shortcut as its own k=1 conv feeding `residual`. */
int hyvae_conv3d_causal_tc_shortcut(const hyvae_vol* x, const void* w, const float* bias, const hyvae_vol* sc_x,
                                    const void* sc_w, const hyvae_vol* y, double* gn_partials, int32_t gn_groups,
                                    int32_t w_has_fold /* 1: w holds the 45 slices of variant bit 8 above */, void* stream);
/* One output-parity phase of UpsampleCausal3D.forward (nearest x2 + 3x3x3 CausalConv3d, unet_causal_3d_blocks.py:
 * 152-175) computed directly from the LOW-resolution volume: the 27 high-res taps collapse onto nkt x 2 x 2 low-res taps
 * (nkt = 2 if up_t == 2, else 3) whose weights are sums of the original ones.
 *   x:  low-res volume carrying the halo (nkt-1, 1, 1);  y: high-res volume (T' = 2T-1 if up_t == 2 else T, 2H, 2W).
 *   w:  [nkt*2*2][Cout][Cin] combined taps of THIS phase, tap = (kt*2 + kh)*2 + kw, in x's dtype.
 *   pt/ph/pw: output parity (t' = 2j - pt, h' = 2i + ph, w' = 2i' + pw); pt must be 0 when up_t == 1.
 *   gn_partials: as for hyvae_conv3d_causal_tc; the phases of one conv accumulate into the same buffer.
 * The caller issues 4 (up_t == 1) or 8 calls per conv. */
int hyvae_conv3d_upphase_tc(const hyvae_vol* x, const void* w, const float* bias, const hyvae_vol* y, int32_t up_t,
                            int32_t pt, int32_t ph, int32_t pw, double* gn_partials, int32_t gn_groups, void* stream);
/* ---- stride-1 3x3x3 CausalConv3d as Winograd F(2,3) along T ------------------------------------------
 * Same op as hyvae_conv3d_causal_tc (k = 3, stride 1; unet_causal_3d_blocks.py:68-75, residual :415) with 1.5x fewer MACs:
 * two output frames come from four transformed PLANES and four 9-tap GEMMs instead of six (oracle/winograd.py restates the
 * algebra).  The operand is written by hyvae_groupnorm_apply_wino (every stride-1 3x3x3 conv of a ResnetBlockCausal3D is
 * fed by GroupNorm + SiLU, :359-363,401-405):
 *   planes: [B][NP][H+2][W+2][Cin], halo (0,1,1), NP = hyvae_wino_planes(T) = 1 + 4*((T-1)/2) (+3 for an even T):
 *           plane 0 = f(x[0]); pair p (output frames 2p+1, 2p+2): d0 - d2, d1 + d2, d2 - d1, d1 - d3 with
 *           d = f(x[max(2p-1,0)]), f(x[2p]), f(x[2p+1]), f(x[2p+2]); an even T ends with the first three of
 *           (x[T-3], x[T-2], x[T-1]).
 *   uw:     [5][9][Cout][Cin] tap groups in the planes' dtype: g0, (g0+g1+g2)/2, (g0-g1+g2)/2, g2, g0+g1+g2 (g_kt = W[:,:,kt]),
 *           each [kh*3+kw][Cout][Cin].
 *   y, residual, bias, gn_partials: as for hyvae_conv3d_causal_tc (y may carry a halo: only its interior is written).
 *   sc_x, sc_w (or NULL, NULL): fused 1x1x1 conv_shortcut of the resnet block (:338-348,407-415) as in
 *           hyvae_conv3d_causal_tc_shortcut: y = conv(x) + conv1x1x1(sc_x) + bias with bias = conv2.bias + conv_shortcut.bias;
 *           sc_x is the block input (y's extent, any halo, C % 8 == 0) and sc_w is [2][Cout][Csc] = (+Ws, -Ws): the second
 *           output frame of a pair is M1 - M2 - M3, so its shortcut term rides, negated, in M3's accumulator.
 * Cin % 64 == 0 and Cout % 128 == 0, else HYVAE_EUNSUPPORTED (the caller then runs the plain path). */
int32_t hyvae_wino_planes(int32_t T);
int hyvae_groupnorm_apply_wino(const hyvae_vol* x, const double* sums, const float* gamma, const float* beta, int32_t groups,
                               float eps, int32_t silu, const hyvae_vol* planes, void* stream);
int hyvae_conv3d_causal_wino(const hyvae_vol* planes, int32_t T, const void* uw, const float* bias, const hyvae_vol* residual,
                             const hyvae_vol* sc_x, const void* sc_w, const hyvae_vol* y, double* gn_partials, int32_t gn_groups,
                             double* gn_sums, void* stream);
/* rows-per-batch of the gn_partials buffer: gn_partials is [B][rows][gn_groups][2] fp64, ZEROED by the caller; every
 * (CTA, warp) of the conv accumulates into its own row, so several launches may add into one buffer (the phases of
 * an upsampling conv) before hyvae_groupnorm_finalize reduces the rows in a fixed order. */
int64_t hyvae_conv3d_tc_gn_rows(void);
/* Size (in doubles) of a gn_partials buffer for B batch items and `groups` groups: the [B][rows][groups][2] warp rows,
 * followed by per-CTA rows and an arrival ticket.  hyvae_conv3d_causal_wino uses the tail when `gn_sums` ([B][groups][2])
 * is given: it then finishes the statistics itself (every CTA folds its warp rows, the last CTA to arrive adds the CTA
 * rows in index order: bit-reproducible) and leaves the whole buffer zeroed, so no hyvae_groupnorm_finalize launch follows. */
int64_t hyvae_gn_partials_doubles(int32_t B, int32_t groups);

/* ---- GroupNorm (+SiLU) -------------------------------------------------------------------------
 * Replaces nn.GroupNorm(32,C,eps=1e-6) + SiLU, unet_causal_3d_blocks.py:359-363,401-405 and
 * vae.py:131-133,287-291.  Two launches: statistics (fp32/fp64 accumulation) then apply.
 *   sums: [B][groups][2] float64 (sum, sum of squares), written by _stats.
 *   workspace: hyvae_groupnorm_workspace_bytes() bytes (per-block partials, combined in a fixed order by
 *   the last block to finish, so the statistics are bit-reproducible; no floating-point atomics).
 *   y may carry a halo: it is filled with the replicated normalised values. */
int64_t hyvae_groupnorm_workspace_bytes(const hyvae_vol* x, int32_t groups);
int hyvae_groupnorm_stats(const hyvae_vol* x, int32_t groups, double* sums, void* workspace, int64_t workspace_bytes,
                          void* stream);
int hyvae_groupnorm_apply(const hyvae_vol* x, const double* sums, const float* gamma, const float* beta,
                          int32_t groups, float eps, int32_t silu, int32_t round_like_ref, const hyvae_vol* y,
                          void* stream);
/* Statistics from the PRODUCER: hyvae_conv3d_causal_tc can emit partial sums of its own output (gn_partials,
 * [B][rows][groups][2] fp64 with rows = hyvae_conv3d_tc_gn_rows()); _finalize adds the rows in a fixed order into
 * `sums`, replacing the _stats pass over the tensor, and ZEROES `partials` again, so one buffer zeroed once can
 * serve every conv enqueued on the same stream. */
int hyvae_groupnorm_finalize(double* partials, int32_t B, int64_t rows, int32_t groups, double* sums, void* stream);

/* ---- tile exchange of the multi-GPU decode ------------------------------------------------------
 * Asynchronous device-to-device copy of `bytes` bytes on `stream` (a stream of the SOURCE device) into memory that may
 * live on a peer GPU (rank 0's tile arena mapped through CUDA IPC): copy engines over NVLink, nothing is enqueued on the
 * destination device.  Replaces the per-rank torch.cat / replicated decode of pipeline_hunyuan_video.py:1074-1082. */
int hyvae_peer_copy(void* dst, const void* src, int64_t bytes, void* stream);

/* ---- pad / nearest upsample --------------------------------------------------------------------
 * Replaces F.pad(replicate) :74 and F.interpolate(nearest)+cat of UpsampleCausal3D.forward :152-171:
 * y (T' = 1+up_t*(T-1) if up_t==2, H*up_h, W*up_w, with any halo) <- x.  y->C may exceed x->C (zero channels). */
int hyvae_pad_upsample(const hyvae_vol* x, const hyvae_vol* y, int32_t up_t, int32_t up_h, int32_t up_w, void* stream);
/* Replicate halo of a volume whose interior was written in place (a conv may write into the interior of a padded `y`:
 * every entry point honours y->pt/ph/pw): only the halo voxels are written, from the clamped interior voxel.  Spares the
 * full-tensor pad pass (F.pad of the NEXT CausalConv3d, :74) between a resnet block and a down/upsampler. */
int hyvae_halo_fill(const hyvae_vol* y, void* stream);

/* ---- mid-block attention softmax ---------------------------------------------------------------
 * Replaces prepare_causal_attention_mask :38-46 + the softmax inside F.scaled_dot_product_attention
 * (diffusers Attention, call site :661): P[i][j] = softmax_j(scale*S[i][j]) over j < (i/n_hw+1)*n_hw,
 * 0 elsewhere.  S: [B][L][L] fp32, P: [B][L][L] in `p_dtype`.  The mask is never materialised. */
int hyvae_softmax_frame_causal(const float* S, void* P, int32_t p_dtype, int32_t B, int32_t L, int32_t n_hw,
                               float scale, void* stream);

/* ---- fused mid-block attention core ------------------------------------------------------------
 * Replaces, in one tcgen05 kernel, F.scaled_dot_product_attention with the additive frame-causal mask of
 * prepare_causal_attention_mask (unet_causal_3d_blocks.py:38-46; diffusers Attention call site :661, one head):
 *   O[i][:] = sum_j softmax_j(scale * Q[i].K[j], j < (i/n_hw+1)*n_hw) * V[j][:] + bv
 * q, k: [L][D]; vt: V transposed, [D][L]; o: [L][D]; all `dtype` (bf16/f16), fp32 bv[D] (may be NULL; rows of the
 * softmax sum to 1, so the value projection's bias is added once at the end).  S and P stay in TMEM / shared memory.
 * Returns HYVAE_EUNSUPPORTED unless D is 128, 256 or 512 and L % 8 == 0: the caller then runs the unfused
 * GEMM -> hyvae_softmax_frame_causal -> GEMM schedule. */
int hyvae_attn_block_causal(const void* q, const void* k, const void* vt, const float* bv, void* o, int32_t dtype,
                            int64_t L, int32_t n_hw, int32_t D, float scale, void* stream);

/* ---- reconstruction metrics of the stride/pool/bucket experiments --------------------------------
 * video_to_frames_u8 replaces save_videos_grid's quantisation (hyvideo/utils/file_utils.py:58-66) for one video:
 *   src (C, T, H, W) with element strides strides_cthw[4] -> dst frames [T][H][W][C] uint8,
 *   x = (x+1)/2 if rescale; clamp(0,1); (uint8)(x*255) (truncation), all in fp32.
 * frame_metrics_u8 replaces compute_psnr / compute_ssim of evaluation/compute_metrics.py:31-41 per frame pair
 * (a = original, b = reconstruction, both [N][H][W][C] uint8, C <= 4): the integer sum of squared differences and the
 * value ranges (PSNR, data_range and the constant-frame rule are finished by the caller), and scikit-image's
 * structural_similarity(win_size=7, uniform window, sample covariance, K1=.01, K2=.03, data_range=max(a)-min(a),
 * channel_axis=-1) in fp64.  workspace: hyvae_frame_metrics_workspace_bytes(N,H,W) + N*sizeof(hyvae_frame_stats) bytes. */
typedef struct { uint64_t ssd; int32_t min_a, max_a, min_b, max_b; int32_t pad[2]; } hyvae_frame_stats;
int hyvae_video_to_frames_u8(const void* src, int32_t dtype, const int64_t* strides_cthw, int32_t C, int32_t T, int32_t H,
                             int32_t W, int32_t rescale, void* dst, void* stream);
int64_t hyvae_frame_metrics_workspace_bytes(int32_t N, int32_t H, int32_t W);
int hyvae_frame_metrics_u8(const void* a, const void* b, int32_t N, int32_t H, int32_t W, int32_t C, void* stats,
                           double* ssim, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- temporal ops of the stride/pool/bucket experiments ---------------------------------------
 * avgpool_t replaces F.pad((0,0,0,0,k-1,0),'replicate')+F.avg_pool3d((k,1,1),(s,1,1)) :665-668,767-772;
 * interp_t replaces F.interpolate(scale_factor=(sc,1,1), mode='nearest') :893-897,906-910
 * (y.T = floor(x.T*scale); src = min(floor(dst*inv_scale), x.T-1) with inv_scale = (float)(1/scale), which is
 * ATen's nearest source-index rule). */
int hyvae_avgpool_t(const hyvae_vol* x, const hyvae_vol* y, int32_t k, int32_t s, void* stream);
int hyvae_interp_t_nearest(const hyvae_vol* x, const hyvae_vol* y, float inv_scale, void* stream);
/* The other modes F.interpolate accepts for a 5-D tensor (the t-ops JSON passes `interp_mode` through, :889-897,904-910):
 * 'trilinear' (align_corners=False; with unit H / W scale a linear interpolation of two frames), 'area' (adaptive average)
 * and 'nearest-exact', with ATen's source-index rules for a given scale_factor.  y.T = floor(x.T * scale). */
enum { HYVAE_INTERP_NEAREST = 0, HYVAE_INTERP_LINEAR = 1, HYVAE_INTERP_AREA = 2, HYVAE_INTERP_NEAREST_EXACT = 3 };
int hyvae_interp_t(const hyvae_vol* x, const hyvae_vol* y, int32_t mode, float inv_scale, void* stream);

/* ---- tile blend + crop + scatter ---------------------------------------------------------------
 * Replaces blend_v / blend_h / blend_t (autoencoder_kl_causal_3d.py:344-360), the [:limit] crops and
 * torch.cat (:410-412,463-465,501,537) with one pass.  All tensors are dense [N][Y][X]:
 *   cur   [N][Yc][Xc]  blended IN PLACE (first ev rows from `above`, then first eh columns from `left`)
 *   above [N][Ya][Xc]  or NULL: cur[y] = above[Ya-ev+y]*(1-y/ev) + cur[y]*(y/ev)
 *   left  [N][Yc][Xl]  or NULL: cur[x] = left[Xl-eh+x]*(1-x/eh) + cur[x]*(x/eh)
 *   out   rows [y0, y0+crop_y) x cols [x0, x0+crop_x) of a dense [N][Yo][Xo] tensor <- cur[:crop_y,:crop_x]
 * Spatial tiles use N=(b,c,t), Y=h, X=w; temporal tiles use N=(b,c), Y=t, X=(h,w).
 * n_strides: NULL for dense tensors, else the element stride between consecutive n of {cur, above, left, out}
 * (lets a caller blend a view that dropped leading rows, e.g. tile[:, :, 1:] at :491,527).
 * post = 1: `out` is FP32 and receives the pipeline tail's image float(clamp(T(v / 2 + 0.5), 0, 1))
 * (pipeline_hunyuan_video.py:1090-1092) instead of v: the last assembly pass of a decode emits the final image. */
int hyvae_blend_crop_scatter(void* cur, const void* above, const void* left, int32_t dtype, int64_t N,
                             int32_t Yc, int32_t Xc, int32_t Ya, int32_t Xl, int32_t ev, int32_t eh,
                             void* out, int32_t Yo, int32_t Xo, int32_t y0, int32_t x0, int32_t crop_y,
                             int32_t crop_x, const int64_t* n_strides, int32_t post, void* stream);

/* ---- pipeline tail ------------------------------------------------------------------------------
 * Replaces `image = (image / 2 + 0.5).clamp(0, 1); image = image.cpu().float()` of
 * pipeline_hunyuan_video.py:1090-1092 (device part): dst[i] = float(clamp(T(src[i] / 2 + 0.5), 0, 1)), one pass.
 * Any n and any of the three dtypes; 16-byte (fp32: 32-byte) aligned pointers take the vector path.  src in `src_dtype`, dst fp32. */
int hyvae_image_postprocess(const void* src, int32_t src_dtype, float* dst, int64_t n, void* stream);

/* ---- measurement hooks (bench.py) ---------------------------------------------------------------
 * profile_begin/end bracket a region; while on, every C-ABI call is timed with two CUDA events on its
 * stream.  profile_end synchronises and returns, per kernel class (0 conv_tc, 1 conv_direct, 2 gn_stats,
 * 3 gn_apply, 4 pad_upsample, 5 softmax, 6 layout, 7 blend, 8 temporal, 9 attn, 10 attn_proj), the summed device
 * milliseconds, the summed ALGORITHMIC work (flops for convs, bytes for the HBM-bound classes) and the launch count.
 * profile_class_override(c >= 0) books the conv launches that follow under class c until reset with -1: the attention's
 * q/k/v/out projections are k=1 launches of the conv kernels but are not nn.Conv3d FLOPs of the reference. */
#define HYVAE_PROFILE_CLASSES 11
int hyvae_profile_class_override(int32_t cls);
int hyvae_profile_begin(void);
int hyvae_profile_end(double* ms, double* work, int64_t* launches, int32_t n_classes);
/* Tensor-core flops actually issued by the conv launches since hyvae_profile_begin: equals the algorithmic work except
 * for hyvae_conv3d_upphase_tc, which needs 8/27 (12/27) of the reference's MACs. */
double hyvae_profile_executed_flops(void);

/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
int64_t hyvae_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HYVAE_H_ */
