// CausalConv3d as a CUDA-core implicit GEMM (NT form: y[m][n] = sum_k A[m][k] * W[n][k]).
//
// This is the GENERAL path: any Cin/Cout, k in {1,3}, any stride, fp32 / bf16 / fp16, replicate
// padding by index clamping, optional nearest-upsample folded into the gather.  It carries the
// thin layers (conv_in 3->128, conv_out 128->3, quant convs), the fp32 model, and every shape the
// tcgen05 kernel (conv_tc.cu) does not accept.  Reference: unet_causal_3d_blocks.py:49-75.
#include "common.cuh"

namespace hyvae {

struct DirectArgs {
  Vol x, y, res;
  const void* w;
  const float* bias;
  int k, st, sh, sw, up_t, up_h, up_w;
  int Tl, Hl, Wl;  // logical conv-input dims (after the optional upsample)
  int Cin, Cout;
  int64_t M;
  int round_like_ref;
};

constexpr int DM = 64, DN = 64, DK = 16;

template <typename T, typename OT, bool VEC4>
__global__ void __launch_bounds__(256) conv_direct_kernel(DirectArgs a) {
  __shared__ float As[DK][DM + 4];
  __shared__ float Bs[DK][DN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * DM;
  const int n0 = blockIdx.y * DN;
  const T* xs = reinterpret_cast<const T*>(a.x.p);
  const T* ws = reinterpret_cast<const T*>(a.w);

  // loader mapping: one row (voxel / cout) and 4 consecutive k per thread
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int64_t lm = m0 + lrow;
  const bool lvalid = lm < a.M;
  int ob = 0, ot = 0, oh = 0, ow = 0;
  if (lvalid) {
    int64_t r = lm;
    ow = (int)(r % a.y.W); r /= a.y.W;
    oh = (int)(r % a.y.H); r /= a.y.H;
    ot = (int)(r % a.y.T); ob = (int)(r / a.y.T);
  }
  const int ncout = n0 + lrow;
  const bool nvalid = ncout < a.Cout;

  // compute mapping: 4 voxels x 4 couts per thread
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int k = a.k, pad_s = k / 2, pad_t = k - 1;
  const int taps = k * k * k;
  for (int tap = 0; tap < taps; ++tap) {
    const int kt = tap / (k * k), kh = (tap / k) % k, kw = tap % k;
    int64_t xoff = 0;
    if (lvalid) {
      int ti = min(max(ot * a.st + kt - pad_t, 0), a.Tl - 1);
      int hi = min(max(oh * a.sh + kh - pad_s, 0), a.Hl - 1);
      int wi = min(max(ow * a.sw + kw - pad_s, 0), a.Wl - 1);
      if (a.up_t == 2) ti = (ti == 0) ? 0 : 1 + ((ti - 1) >> 1);
      if (a.up_h == 2) hi >>= 1;
      if (a.up_w == 2) wi >>= 1;
      xoff = a.x.at(ob, ti, hi, wi);
    }
    const T* wrow = ws + ((int64_t)tap * a.Cout + ncout) * a.Cin;
    for (int c0 = 0; c0 < a.Cin; c0 += DK) {
      float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
      const int c = c0 + lk;
      if (VEC4) {
        if (lvalid && c < a.Cin) {
          if (sizeof(T) == 4) {
            float4 q = *reinterpret_cast<const float4*>(xs + xoff + c);
            av[0] = q.x; av[1] = q.y; av[2] = q.z; av[3] = q.w;
          } else {
            uint2 q = *reinterpret_cast<const uint2*>(xs + xoff + c);
            const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) av[j] = to_f<T>(e[j]);
          }
        }
        if (nvalid && c < a.Cin) {
          if (sizeof(T) == 4) {
            float4 q = *reinterpret_cast<const float4*>(wrow + c);
            bv[0] = q.x; bv[1] = q.y; bv[2] = q.z; bv[3] = q.w;
          } else {
            uint2 q = *reinterpret_cast<const uint2*>(wrow + c);
            const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = to_f<T>(e[j]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (lvalid && c + j < a.Cin) av[j] = to_f<T>(xs[xoff + c + j]);
          if (nvalid && c + j < a.Cin) bv[j] = to_f<T>(wrow[c + j]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) { As[lk + j][lrow] = av[j]; Bs[lk + j][lrow] = bv[j]; }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < DK; ++kk) {
        float4 fa = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 fb = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float ar[4] = {fa.x, fa.y, fa.z, fa.w}, br[4] = {fb.x, fb.y, fb.z, fb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
    }
  }

  OT* yd = reinterpret_cast<OT*>(a.y.p);
  const T* rs = reinterpret_cast<const T*>(a.res.p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
    int64_t r = m;
    int w = (int)(r % a.y.W); r /= a.y.W;
    int h = (int)(r % a.y.H); r /= a.y.H;
    int t = (int)(r % a.y.T); int b = (int)(r / a.y.T);
    const int64_t yo = a.y.at(b, t, h, w);
    const int64_t ro = rs ? a.res.at(b, t, h, w) : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.Cout) continue;
      float v = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
      if (rs) {
        if (a.round_like_ref) v = rnd<T>(v);
        v += to_f<T>(rs[ro + n]);
      }
      yd[yo + n] = from_f<OT>(v);
    }
  }
}

}  // namespace hyvae

using namespace hyvae;

extern "C" int hyvae_conv3d_causal_direct(const hyvae_vol* x, const void* w, const float* bias, const hyvae_vol* residual,
                                          const hyvae_vol* y, int32_t k, int32_t st, int32_t sh, int32_t sw,
                                          int32_t up_t, int32_t up_h, int32_t up_w, int32_t round_like_ref, void* stream) {
  if (int e = check_vol(x, "x")) return e;
  if (int e = check_vol(y, "y")) return e;
  HYVAE_CHECK_ARG(w != nullptr, "w is null");
  HYVAE_CHECK_ARG(k == 1 || k == 3, "kernel size %d not supported (1 or 3)", k);
  HYVAE_CHECK_ARG(st >= 1 && sh >= 1 && sw >= 1, "bad stride");
  HYVAE_CHECK_ARG((up_t == 1 || up_t == 2) && (up_h == 1 || up_h == 2) && (up_w == 1 || up_w == 2), "up factors must be 1 or 2");
  HYVAE_CHECK_ARG((x->dtype == y->dtype || y->dtype == HYVAE_F32) && x->B == y->B, "x / y dtype or batch mismatch");
  HYVAE_CHECK_ARG(x->dtype == y->dtype || residual == nullptr, "residual needs y in x's dtype");
  DirectArgs a;
  a.x = make_vol(x); a.y = make_vol(y);
  a.Tl = up_t == 2 ? 1 + 2 * (x->T - 1) : x->T;
  a.Hl = x->H * up_h; a.Wl = x->W * up_w;
  // nn.Conv3d on the padded input: out = floor((L + (k-1) - k) / s) + 1 = floor((L - 1) / s) + 1
  HYVAE_CHECK_ARG(y->T == (a.Tl - 1) / st + 1 && y->H == (a.Hl - 1) / sh + 1 && y->W == (a.Wl - 1) / sw + 1,
                  "y dims %dx%dx%d do not match conv of %dx%dx%d stride %d,%d,%d", y->T, y->H, y->W, a.Tl, a.Hl, a.Wl, st, sh, sw);
  if (residual) {
    if (int e = check_vol(residual, "residual")) return e;
    HYVAE_CHECK_ARG(residual->dtype == y->dtype && residual->B == y->B && residual->T == y->T && residual->H == y->H &&
                    residual->W == y->W && residual->C == y->C, "residual shape mismatch");
    a.res = make_vol(residual);
  } else {
    a.res = a.y; a.res.p = nullptr;
  }
  a.w = w; a.bias = bias; a.k = k; a.st = st; a.sh = sh; a.sw = sw; a.up_t = up_t; a.up_h = up_h; a.up_w = up_w;
  a.Cin = x->C; a.Cout = y->C; a.M = (int64_t)y->B * y->T * y->H * y->W; a.round_like_ref = round_like_ref;
  dim3 grid((unsigned)((a.M + DM - 1) / DM), (unsigned)((a.Cout + DN - 1) / DN));
  ProfScope prof(PC_CONV_DIRECT, 2.0 * a.M * a.Cout * a.Cin * k * k * k, stream);
  const bool vec = (a.Cin % 4 == 0);
  const bool f32out = (y->dtype == HYVAE_F32);
  HYVAE_DISPATCH_DTYPE(x->dtype, T, {
    if (f32out) {
      if (vec) conv_direct_kernel<T, float, true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
      else conv_direct_kernel<T, float, false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    } else {
      if (vec) conv_direct_kernel<T, T, true><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
      else conv_direct_kernel<T, T, false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    }
  });
  return check_launch("conv3d_causal_direct");
}
