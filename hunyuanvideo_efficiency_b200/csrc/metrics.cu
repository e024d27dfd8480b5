// Reconstruction metrics of the stride / pool / bucket experiments on the GPU: the frame quantisation of
// save_videos_grid (hyvideo/utils/file_utils.py:58-66) and the per-frame PSNR / SSIM of evaluation/compute_metrics.py
// (:31-41).  SSIM there is scikit-image's structural_similarity(win_size=7 uniform window, sample covariance,
// K1=0.01, K2=0.03, data_range = max - min of the first frame, channel_axis=-1) — third-party code restated from its
// published algorithm: S = ((2 ux uy + C1)(2 vxy + C2)) / ((ux^2 + uy^2 + C1)(vx + vy + C2)) per pixel and channel,
// averaged over the pixels whose 7x7 window lies inside the frame (the crop of (win-1)/2 = 3 border pixels), then over
// channels.  All integer arithmetic (window sums of uint8 values and products, squared differences) is exact; the
// per-pixel S is evaluated in fp64 like skimage does for uint8 input.  HBM-bound: each frame is read once per kernel.
#include "common.cuh"

namespace hyvae {

// ---- video (C, T, H, W), any float dtype, arbitrary strides  ->  frames [T][H][W][C] uint8
template <typename T>
__global__ void video_to_frames_u8_kernel(const T* __restrict__ src, int64_t sC, int64_t sT, int64_t sH, int64_t sW, int C, int Tn,
                                          int H, int W, int rescale, uint8_t* __restrict__ dst) {
  const int64_t n = (int64_t)Tn * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const int t = (int)(i / ((int64_t)W * H));
    const T* p = src + (int64_t)t * sT + (int64_t)h * sH + (int64_t)w * sW;
    for (int c = 0; c < C; ++c) {
      float v = to_f<T>(p[(int64_t)c * sC]);
      if (rescale) v = (v + 1.0f) / 2.0f;          // file_utils.py:64
      v = fminf(fmaxf(v, 0.f), 1.f);               // :65
      dst[i * C + c] = (uint8_t)(int)(v * 255.f);  // :66  astype(uint8) truncates
    }
  }
}

// ---- per frame: sum of squared differences, min / max of a and of b (integer atomics: order independent)
struct FrameStats { unsigned long long ssd; int min_a, max_a, min_b, max_b; int pad[2]; };
static_assert(sizeof(FrameStats) == sizeof(hyvae_frame_stats), "ABI record");

__global__ void frame_stats_init_kernel(FrameStats* st, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { st[i].ssd = 0ull; st[i].min_a = 255; st[i].max_a = 0; st[i].min_b = 255; st[i].max_b = 0; st[i].pad[0] = st[i].pad[1] = 0; }
}

__global__ void frame_stats_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int64_t per_frame, int vec, FrameStats* st) {
  const int f = blockIdx.y;
  const uint8_t* pa = a + (int64_t)f * per_frame;
  const uint8_t* pb = b + (int64_t)f * per_frame;
  unsigned long long ssd = 0ull;
  int mna = 255, mxa = 0, mnb = 255, mxb = 0;
  const int64_t nvec = vec ? per_frame / 16 : 0;  // vec: every frame starts 16-byte aligned (checked by the host)
  const uint4* va = reinterpret_cast<const uint4*>(pa);
  const uint4* vb = reinterpret_cast<const uint4*>(pb);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 x = va[i], y = vb[i];
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int s = 0; s < 32; s += 8) {
        const int p = (xs[k] >> s) & 255, q = (ys[k] >> s) & 255, d = p - q;
        acc += (uint32_t)(d * d);
        mna = min(mna, p); mxa = max(mxa, p); mnb = min(mnb, q); mxb = max(mxb, q);
      }
    }
    ssd += acc;
  }
  if (blockIdx.x == 0) {  // tail bytes
    for (int64_t i = nvec * 16 + threadIdx.x; i < per_frame; i += blockDim.x) {
      const int p = pa[i], q = pb[i], d = p - q;
      ssd += (unsigned long long)(d * d);
      mna = min(mna, p); mxa = max(mxa, p); mnb = min(mnb, q); mxb = max(mxb, q);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ssd += __shfl_xor_sync(0xffffffffu, ssd, o);
    mna = min(mna, __shfl_xor_sync(0xffffffffu, mna, o)); mxa = max(mxa, __shfl_xor_sync(0xffffffffu, mxa, o));
    mnb = min(mnb, __shfl_xor_sync(0xffffffffu, mnb, o)); mxb = max(mxb, __shfl_xor_sync(0xffffffffu, mxb, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&st[f].ssd, ssd);
    atomicMin(&st[f].min_a, mna); atomicMax(&st[f].max_a, mxa);
    atomicMin(&st[f].min_b, mnb); atomicMax(&st[f].max_b, mxb);
  }
}

// ---- SSIM: one thread per window centre, all channels; block = 32 x 8 centres, halo tile in shared memory
constexpr int SSIM_BX = 32, SSIM_BY = 8, SSIM_WIN = 7, SSIM_MAXC = 4;

__global__ void __launch_bounds__(SSIM_BX * SSIM_BY)
frame_ssim_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int H, int W, int C, const FrameStats* __restrict__ st,
                  double* __restrict__ partial, int blocks_per_frame) {
  __shared__ uint8_t sa[(SSIM_BY + 6) * (SSIM_BX + 6) * SSIM_MAXC];
  __shared__ uint8_t sb[(SSIM_BY + 6) * (SSIM_BX + 6) * SSIM_MAXC];
  __shared__ double red[SSIM_BX * SSIM_BY / 32];
  const int f = blockIdx.z;
  const int x0 = blockIdx.x * SSIM_BX, y0 = blockIdx.y * SSIM_BY;  // top-left of the halo tile = first centre - 3
  const int vw = W - 6, vh = H - 6;                               // number of valid centres per axis
  const uint8_t* pa = a + (int64_t)f * H * W * C;
  const uint8_t* pb = b + (int64_t)f * H * W * C;
  const int tw = SSIM_BX + 6, th = SSIM_BY + 6;
  for (int i = threadIdx.x; i < th * tw * C; i += blockDim.x) {
    const int c = i % C, xx = (i / C) % tw, yy = i / (C * tw);
    const int gx = x0 + xx, gy = y0 + yy;
    uint8_t u = 0, v = 0;
    if (gx < W && gy < H) { const int64_t o = ((int64_t)gy * W + gx) * C + c; u = pa[o]; v = pb[o]; }
    sa[i] = u; sb[i] = v;
  }
  __syncthreads();
  const int lx = threadIdx.x % SSIM_BX, ly = threadIdx.x / SSIM_BX;
  double s_sum = 0.0;
  if (x0 + lx < vw && y0 + ly < vh) {
    const double R = (double)(st[f].max_a - st[f].min_a);  // data_range = img1.max() - img1.min(), compute_metrics.py:41
    const double C1 = (0.01 * R) * (0.01 * R), C2 = (0.03 * R) * (0.03 * R);
    const double NP = 49.0, cov_norm = NP / (NP - 1.0);
    for (int c = 0; c < C; ++c) {
      int sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
      for (int dy = 0; dy < SSIM_WIN; ++dy) {
#pragma unroll
        for (int dx = 0; dx < SSIM_WIN; ++dx) {
          const int o = ((ly + dy) * tw + lx + dx) * C + c;
          const int p = sa[o], q = sb[o];
          sx += p; sy += q; sxx += p * p; syy += q * q; sxy += p * q;
        }
      }
      const double ux = sx / NP, uy = sy / NP, uxx = sxx / NP, uyy = syy / NP, uxy = sxy / NP;
      const double vx = cov_norm * (uxx - ux * ux), vy = cov_norm * (uyy - uy * uy), vxy = cov_norm * (uxy - ux * uy);
      const double A1 = 2.0 * ux * uy + C1, A2 = 2.0 * vxy + C2, B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
      s_sum += (A1 * A2) / (B1 * B2);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s_sum += __shfl_xor_sync(0xffffffffu, s_sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < SSIM_BX * SSIM_BY / 32; ++i) t += red[i];
    partial[(int64_t)f * blocks_per_frame + blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

// fixed-order sum of a frame's block partials -> mean SSIM over centres and channels
__global__ void frame_ssim_finalize_kernel(const double* __restrict__ partial, int blocks_per_frame, double count, double* __restrict__ out, int N) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= N) return;
  double t = 0.0;
  for (int i = 0; i < blocks_per_frame; ++i) t += partial[(int64_t)f * blocks_per_frame + i];
  out[f] = t / count;
}

}  // namespace hyvae

using namespace hyvae;

extern "C" int hyvae_video_to_frames_u8(const void* src, int32_t dtype, const int64_t* strides_cthw, int32_t C, int32_t T, int32_t H,
                                        int32_t W, int32_t rescale, void* dst, void* stream) {
  HYVAE_CHECK_ARG(src && dst && strides_cthw, "video_to_frames_u8: null pointer");
  HYVAE_CHECK_ARG(C > 0 && T > 0 && H > 0 && W > 0, "video_to_frames_u8: empty video");
  HYVAE_CHECK_ARG(dtype == HYVAE_BF16 || dtype == HYVAE_F16 || dtype == HYVAE_F32, "video_to_frames_u8: bad dtype %d", dtype);
  const int64_t n = (int64_t)T * H * W;
  ProfScope prof(PC_LAYOUT, (double)n * C * (dtype_size(dtype) + 1), stream, "video_to_frames_u8");
  const int threads = 256;
  const int64_t want = (n + threads - 1) / threads;
  const int blocks = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
  HYVAE_DISPATCH_DTYPE(dtype, TT, (video_to_frames_u8_kernel<TT><<<blocks, threads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const TT*>(src), strides_cthw[0], strides_cthw[1], strides_cthw[2], strides_cthw[3], C, T, H, W, rescale,
      reinterpret_cast<uint8_t*>(dst))));
  return check_launch("video_to_frames_u8");
}

extern "C" int64_t hyvae_frame_metrics_workspace_bytes(int32_t N, int32_t H, int32_t W) {
  if (N <= 0 || H < SSIM_WIN || W < SSIM_WIN) return -1;
  const int64_t bx = (W - 6 + SSIM_BX - 1) / SSIM_BX, by = (H - 6 + SSIM_BY - 1) / SSIM_BY;
  return (int64_t)N * bx * by * 8;
}

// stats: hyvae_frame_stats[N]; ssim: double [N]
extern "C" int hyvae_frame_metrics_u8(const void* a, const void* b, int32_t N, int32_t H, int32_t W, int32_t C, void* stats,
                                      double* ssim, void* workspace, int64_t workspace_bytes, void* stream) {
  HYVAE_CHECK_ARG(a && b && stats && ssim && workspace, "frame_metrics_u8: null pointer");
  HYVAE_CHECK_ARG(N > 0 && C > 0 && C <= SSIM_MAXC, "frame_metrics_u8: needs 1..%d channels and N > 0 (C=%d N=%d)", SSIM_MAXC, C, N);
  HYVAE_CHECK_ARG(H >= SSIM_WIN && W >= SSIM_WIN, "frame_metrics_u8: win_size 7 exceeds the frame (%d x %d)", H, W);  // skimage raises too
  HYVAE_CHECK_ARG(N <= 65535, "frame_metrics_u8: at most 65535 frames per call");
  const int64_t need = hyvae_frame_metrics_workspace_bytes(N, H, W);
  HYVAE_CHECK_ARG(workspace_bytes >= need + (int64_t)N * (int64_t)sizeof(FrameStats), "frame_metrics_u8: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t per_frame = (int64_t)H * W * C;
  ProfScope prof(PC_LAYOUT, 4.0 * (double)N * per_frame, stream, "frame_metrics_u8");
  double* partial = reinterpret_cast<double*>(workspace);
  FrameStats* fs = reinterpret_cast<FrameStats*>(reinterpret_cast<uint8_t*>(workspace) + need);
  frame_stats_init_kernel<<<(N + 127) / 128, 128, 0, st>>>(fs, N);
  if (int e = check_launch("frame_stats_init")) return e;
  {
    // vector loads need every frame to start 16-byte aligned; otherwise the whole frame goes through the scalar tail
    const bool vec_ok = (per_frame % 16 == 0) && (((uintptr_t)a | (uintptr_t)b) % 16 == 0);
    int bx = (int)((per_frame / 16 + 255) / 256);
    bx = bx < 1 ? 1 : (bx > num_sms() * 4 ? num_sms() * 4 : bx);
    if (!vec_ok) bx = 1;
    frame_stats_kernel<<<dim3((unsigned)bx, (unsigned)N), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(a), reinterpret_cast<const uint8_t*>(b),
                                                                       per_frame, vec_ok ? 1 : 0, fs);
    if (int e = check_launch("frame_stats")) return e;
  }
  const int bx = (W - 6 + SSIM_BX - 1) / SSIM_BX, by = (H - 6 + SSIM_BY - 1) / SSIM_BY;
  frame_ssim_kernel<<<dim3((unsigned)bx, (unsigned)by, (unsigned)N), SSIM_BX * SSIM_BY, 0, st>>>(
      reinterpret_cast<const uint8_t*>(a), reinterpret_cast<const uint8_t*>(b), H, W, C, fs, partial, bx * by);
  if (int e = check_launch("frame_ssim")) return e;
  frame_ssim_finalize_kernel<<<(N + 127) / 128, 128, 0, st>>>(partial, bx * by, (double)(W - 6) * (double)(H - 6) * C, ssim, N);
  if (int e = check_launch("frame_ssim_finalize")) return e;
  // stats out: the raw 32-byte records (hyvae_frame_stats in include/hyvae.h)
  if (cudaMemcpyAsync(stats, fs, (size_t)N * sizeof(FrameStats), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return fail(HYVAE_ECUDA, "frame_metrics_u8: stats copy failed");
  return HYVAE_OK;
}
