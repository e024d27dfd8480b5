// HBM-bound passes of the causal 3D VAE: layout conversion, GroupNorm(+SiLU), replicate-pad /
// nearest-upsample, frame-causal softmax, temporal pool / interp, tile blend+crop+scatter.
// All are coalesced, 16-byte vectorised along the channel (or W) axis, and written in the "gather"
// form: one thread per DESTINATION element, source index by clamping — which is what turns the
// reference's F.pad(replicate) copies (unet_causal_3d_blocks.py:74) into index arithmetic.
#include <cstdlib>

#include <type_traits>

#include "common.cuh"

namespace hyvae {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
bool g_prof_on = false;
int g_prof_class_override = -1;
double g_prof_exec_flops = 0.0;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_prof_pool;

static inline int grid_for(int64_t n, int block, int cap_mult = 32) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)num_sms() * cap_mult;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------
// NCTHW <-> channels-last volume
// ------------------------------------------------------------------------------------------------
template <typename S, typename D>
__global__ void ncthw_to_vol_kernel(const S* __restrict__ src, Vol d, int src_C, int64_t sb, int64_t sc, int64_t st, int64_t sh, int64_t sw) {
  const int64_t nvox = (int64_t)d.B * d.Tp() * d.Hp() * d.Wp();
  D* dst = reinterpret_cast<D*>(d.p);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvox; i += (int64_t)gridDim.x * blockDim.x) {
    int wp = (int)(i % d.Wp());
    int64_t r = i / d.Wp();
    int hp = (int)(r % d.Hp()); r /= d.Hp();
    int tp = (int)(r % d.Tp());
    int b = (int)(r / d.Tp());
    int t = max(tp - d.pt, 0);
    int h = min(max(hp - d.ph, 0), d.H - 1);
    int w = min(max(wp - d.pw, 0), d.W - 1);
    const S* s = src + b * sb + t * st + h * sh + w * sw;
    D* o = dst + i * d.C;
    if ((d.C & 7) == 0) {  // 16-byte stores (conv_in: 3 channels stored as 16)
      for (int c0 = 0; c0 < d.C; c0 += 8) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (c0 + j) < src_C ? to_f<S>(s[(c0 + j) * sc]) : 0.f;
        Vec8<D> q; q.set(f); q.store(o + c0);
      }
    } else {
      for (int c = 0; c < d.C; ++c) o[c] = c < src_C ? from_f<D>(to_f<S>(s[c * sc])) : from_f<D>(0.f);  // zero channel padding
    }
  }
}

// conv_in operand with the three kw taps packed along the channel axis: destination channel kw * src_C + c of voxel
// (t, h, w) = source channel c of voxel (t, h, clamp(w + kw - 1)); the remaining channels are zero.  A 3x3x3 conv over the
// 3-channel clip then is a 3x3 (kt, kh) conv over 9 (of 16 stored) channels: one K = 16 MMA slice per (kt, kh) instead of
// one per (kt, kh, kw) — a third of the MMAs for the same bytes (hyvae_conv3d_causal_tc, variant bit 9).
template <typename S, typename D>
__global__ void ncthw_to_vol_kw3_kernel(const S* __restrict__ src, Vol d, int src_C, int64_t sb, int64_t sc, int64_t st, int64_t sh, int64_t sw) {
  const int64_t nvox = (int64_t)d.B * d.Tp() * d.Hp() * d.Wp();
  D* dst = reinterpret_cast<D*>(d.p);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvox; i += (int64_t)gridDim.x * blockDim.x) {
    int wp = (int)(i % d.Wp());
    int64_t r = i / d.Wp();
    int hp = (int)(r % d.Hp()); r /= d.Hp();
    int tp = (int)(r % d.Tp());
    int b = (int)(r / d.Tp());
    int t = max(tp - d.pt, 0);
    int h = min(max(hp - d.ph, 0), d.H - 1);
    int w = min(max(wp - d.pw, 0), d.W - 1);
    const S* s = src + b * sb + t * st + h * sh;
    D* o = dst + i * d.C;
    for (int c0 = 0; c0 < d.C; c0 += 8) {   // d.C % 8 == 0 (checked by the host)
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = c0 + j, kw = ch / src_C, c = ch - kw * src_C;
        f[j] = kw < 3 ? to_f<S>(s[min(max(w + kw - 1, 0), d.W - 1) * sw + c * sc]) : 0.f;
      }
      Vec8<D> q; q.set(f); q.store(o + c0);
    }
  }
}

template <typename S, typename D>
__global__ void vol_to_ncthw_kernel(Vol s, D* __restrict__ dst, int dst_C) {
  const int64_t thw = (int64_t)s.T * s.H * s.W;
  const int64_t nvox = (int64_t)s.B * thw;
  const S* src = reinterpret_cast<const S*>(s.p);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvox; i += (int64_t)gridDim.x * blockDim.x) {
    int w = (int)(i % s.W);
    int64_t r = i / s.W;
    int h = (int)(r % s.H); r /= s.H;
    int t = (int)(r % s.T);
    int b = (int)(r / s.T);
    const S* p = src + s.at(b, t, h, w);
    D* o = dst + (int64_t)b * dst_C * thw + ((int64_t)t * s.H + h) * s.W + w;
    for (int c = 0; c < dst_C; ++c) o[c * thw] = from_f<D>(to_f<S>(p[c]));
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics: per (b, group) sum and sum of squares, bit-reproducible run to run.
// grid = (chunks, B); a thread owns one 8-channel vector position and strides over voxels, so its 16
// partial sums stay in registers; the block combines them through shared memory in a FIXED order
// (no floating-point atomics), writes one fp64 partial per (group, moment), and the last block to
// finish (atomic ticket) adds the partials in block order into `sums`.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void gn_stats_kernel(Vol x, int groups, int vox_per_block, double* __restrict__ partials,
                                unsigned int* __restrict__ tickets, double* __restrict__ sums) {
  extern __shared__ float sh[];  // [lanes][2*C]
  __shared__ bool is_last;
  const int C = x.C, CV = C / 8;
  const int b = blockIdx.y, nblk = gridDim.x;
  const int64_t nvox = (int64_t)x.T * x.H * x.W;
  const int64_t v0 = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v1 = min(v0 + vox_per_block, nvox);
  const int lanes = blockDim.x / CV;
  const int cv = threadIdx.x % CV, vl = threadIdx.x / CV;
  if (vl < lanes) {
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
    const T* base = reinterpret_cast<const T*>(x.p);
    const bool dense = (x.pt | x.ph | x.pw) == 0;
    for (int64_t v = v0 + vl; v < v1; v += lanes) {
      int64_t off;
      if (dense) off = ((int64_t)b * nvox + v) * C;
      else {
        int w = (int)(v % x.W); int64_t r = v / x.W; int h = (int)(r % x.H); int t = (int)(r / x.H);
        off = x.at(b, t, h, w);
      }
      Vec8<T> q; q.load(base + off + cv * 8);
      float f[8]; q.get(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] = fmaf(f[j], f[j], ss[j]); }
    }
    float* row = sh + (size_t)vl * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) { row[cv * 8 + j] = s[j]; row[C + cv * 8 + j] = ss[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {  // fixed-order sum over voxel lanes
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += sh[(size_t)l * 2 * C + c];
    sh[c] = a;  // lane-0 slot of this column: only this thread touches column c
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double a = 0.0, q = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a += (double)sh[c]; q += (double)sh[C + c]; }
    double* p = partials + (((int64_t)b * nblk + blockIdx.x) * groups + g) * 2;
    p[0] = a; p[1] = q;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&tickets[b], 1u) == (unsigned)(nblk - 1));
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
      double a = 0.0, q = 0.0;
      for (int k = 0; k < nblk; ++k) {
        const volatile double* p = partials + (((int64_t)b * nblk + k) * groups + g) * 2;
        a += p[0]; q += p[1];
      }
      sums[((int64_t)b * groups + g) * 2 + 0] = a;
      sums[((int64_t)b * groups + g) * 2 + 1] = q;
    }
  }
}

// GroupNorm apply (+SiLU) into a possibly padded destination.  grid = (row chunks, B).
// A block walks destination rows (tp, hp): Wp*C contiguous elements each, read from the clamped source row,
// so the inner loop is a coalesced 16-byte stream with 32-bit index math only.
template <typename T> struct FastMath { static constexpr bool value = true; };
template <> struct FastMath<float> { static constexpr bool value = false; };

// y = silu(u).  16-bit paths: u * rcp(1 + ex2(-u * log2 e)) = FMUL + MUFU.EX2 + FADD + MUFU.RCP + FMUL.  Two MUFU per
// element is 8 elements/clk/SM, ~2.2 T elements/s chip-wide, above what HBM can feed (1.6 T/s); the previous Newton
// reciprocal made the kernel issue bound (ncu: 80 % issue slots busy at 48 % DRAM).  fp32 keeps expf and a division.
// SILU / RLR are template flags: as run-time flags the compiler predicated the rounding instructions, which still
// occupied issue slots (ncu: 21 warp instructions per element instead of ~10).
// 16-bit paths: silu(u) = h + h * tanh(h) with h = u / 2 and ONE MUFU (tanh.approx.f32, max relative error 2^-11) per
// element.  With ex2 + rcp the kernel was MUFU bound: 2 x 1.6e8 ops / (148 SMs x 16 per clk) = 72 us of a 128 us launch.
// The absolute error, <= 2.4e-4 * |u|, is that of rounding a value of size |u| / 2 to fp16, i.e. it is of the size of the
// storage rounding the result receives anyway.  fp32 keeps expf and a true division.
template <typename T, bool SILU, bool RLR>
__device__ __forceinline__ float gn_act(float u) {
  if (RLR) u = rnd<T>(u);
  if (!SILU) return u;
  if (FastMath<T>::value) {
    const float h = 0.5f * u;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  }
  return u / (1.f + expf(-u));
}

template <typename T, bool SILU, bool RLR>
__global__ void __launch_bounds__(256, 3) gn_apply_kernel(Vol x, Vol y, const double* __restrict__ sums, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int groups, float eps, int rows_per_block, double inv_n) {
  extern __shared__ float sh[];  // scale[C], shift[C]
  const int C = x.C, CV = C / 8, b = blockIdx.y;
  const int cpg = C / groups;
  // Every block derives scale / shift itself.  A large tensor has ~17 k blocks of ~130 KB each, so this prologue must
  // stay short: with two fp64 divisions, an fp64 sqrt and an fp64 reciprocal per channel it took a few microseconds in
  // which the block had no loads in flight — a quarter of a block's lifetime at the HBM rate.  Now: inv_n from the
  // host, the variance (the one cancellation-prone step) in fp64, and an IEEE fp32 rsqrt for the 16-bit types.
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    int g = c / cpg;
    const double mean = sums[((int64_t)b * groups + g) * 2] * inv_n;
    double var = fma(-mean, mean, sums[((int64_t)b * groups + g) * 2 + 1] * inv_n);
    if (var < 0) var = 0;
    float rstd;
    if (FastMath<T>::value) rstd = __frsqrt_rn((float)(var + (double)eps));
    else rstd = (float)(1.0 / sqrt(var + (double)eps));
    float sc = gamma[c] * rstd;
    sh[c] = sc;
    sh[C + c] = beta[c] - (float)mean * sc;
  }
  __syncthreads();
  const int Wp = y.Wp(), Hp = y.Hp();
  const int nrows = y.Tp() * Hp;
  const int row_elems = Wp * CV;  // 8-channel vectors per destination row
  const T* xs = reinterpret_cast<const T*>(x.p);
  T* yd = reinterpret_cast<T*>(y.p) + (int64_t)b * y.sB;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, nrows);
  const bool pow2 = (CV & (CV - 1)) == 0;
  const int shift = 31 - __clz(CV);
  const bool fixed_cv = (blockDim.x % CV) == 0;
  const int mycv = threadIdx.x % CV;
  float rsc[8], rsf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { rsc[j] = sh[mycv * 8 + j]; rsf[j] = sh[C + mycv * 8 + j]; }
  for (int r = r0; r < r1; ++r) {
    const int tp = r / Hp, hp = r - tp * Hp;
    const int t = max(tp - y.pt, 0), h = min(max(hp - y.ph, 0), y.H - 1);
    const T* srow = xs + x.at(b, t, h, 0);
    T* drow = yd + (int64_t)r * Wp * C;
    if (fixed_cv) {
      // blockDim % CV == 0: this thread always lands on the same 8 channels -> scale/shift live in registers.
      // Four 16-byte loads are issued before the first use (the kernel was latency bound with two: ncu showed 42 % of
      // the stall samples on the first consumer of the load at 45 % occupancy).
      constexpr int U = 8;
      for (int i = threadIdx.x; i < row_elems; i += U * blockDim.x) {
        Vec8<T> q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int iu = i + u * blockDim.x;
          if (iu < row_elems) {
            const int wp = pow2 ? (iu >> shift) : (iu / CV);
            const int w = min(max(wp - y.pw, 0), y.W - 1);
            q[u].load(srow + (int64_t)w * x.sW + mycv * 8);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int iu = i + u * blockDim.x;
          if (iu < row_elems) {
            float f[8];
            q[u].get(f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = gn_act<T, SILU, RLR>(fmaf(f[j], rsc[j], rsf[j]));
            q[u].set(f);
            q[u].store(drow + (int64_t)iu * 8);
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < row_elems; i += blockDim.x) {
        const int wp = i / CV;
        const int cv = i - wp * CV;
        const int w = min(max(wp - y.pw, 0), y.W - 1);
        Vec8<T> q; q.load(srow + (int64_t)w * x.sW + cv * 8);
        float f[8]; q.get(f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = gn_act<T, SILU, RLR>(fmaf(f[j], sh[cv * 8 + j], sh[C + cv * 8 + j]));
        q.set(f);
        q.store(drow + (int64_t)i * 8);
      }
    }
  }
}

// GroupNorm apply (+SiLU) that writes the Winograd-T PLANE volume of the following stride-1 3x3x3 conv (conv_wino.cu):
//   plane 0 = f(x[0]);  pair p: V0 = d0 - d2, V1 = d1 + d2, V2 = d2 - d1, V3 = d1 - d3 with d = f(x[max(2p-1,0)]), f(x[2p]),
//   f(x[2p+1]), f(x[2p+2]);  even T: V0, V1, V2 of the last frame;  f = SiLU(GroupNorm(.)) evaluated in fp32, each plane
//   rounded once.  A thread owns one (padded row, padded column, 8-channel vector) and walks T, so every input voxel is
//   read once (the two previous frames stay in registers) and 2 planes are written per input frame; the 1-voxel replicate
//   halo of the planes comes from clamping the source coordinates (replicate padding commutes with the transform).
// 4-element (8-byte) vectors of a 16-bit type <-> fp32
template <typename T> struct Vec4h;
template <> struct Vec4h<__half> {
  static __device__ __forceinline__ void get(const uint2& v, float* f) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
  static __device__ __forceinline__ uint2 set(const float* f) {
    uint2 v; const __half2 a = __floats2half2_rn(f[0], f[1]), b = __floats2half2_rn(f[2], f[3]);
    v.x = *reinterpret_cast<const uint32_t*>(&a); v.y = *reinterpret_cast<const uint32_t*>(&b); return v;
  }
};
template <> struct Vec4h<__nv_bfloat16> {
  static __device__ __forceinline__ void get(const uint2& v, float* f) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
  static __device__ __forceinline__ uint2 set(const float* f) {
    uint2 v; const __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
    v.x = *reinterpret_cast<const uint32_t*>(&a); v.y = *reinterpret_cast<const uint32_t*>(&b); return v;
  }
};

// A thread owns FOUR channels of one padded (row, column) and walks T.  The first version (eight channels per thread, one
// frame pair in flight, 95 registers, 2 blocks per SM) was latency bound: 3.7 warps per scheduler, 75 % of the stall
// samples waiting on the loads, 4.8 TB/s (profiles/r02_ncu_gn_apply_wino_128_v1.txt).  Now: half the per-thread state
// (4 blocks per SM) and the loads of the NEXT TWO pairs in flight while two pairs are transformed.
template <typename T, bool SILU>
__global__ void __launch_bounds__(256, 4) gn_apply_wino_kernel(Vol x, Vol y, int Tn, const double* __restrict__ sums, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int groups, float eps, double inv_n) {
  extern __shared__ float sh[];  // scale[C], shift[C]
  const int C = x.C, CQ = C / 4, b = blockIdx.y;
  const int cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double mean = sums[((int64_t)b * groups + g) * 2] * inv_n;
    double var = fma(-mean, mean, sums[((int64_t)b * groups + g) * 2 + 1] * inv_n);
    if (var < 0) var = 0;
    const float rstd = __frsqrt_rn((float)(var + (double)eps));
    const float sc = gamma[c] * rstd;
    sh[c] = sc;
    sh[C + c] = beta[c] - (float)mean * sc;
  }
  __syncthreads();
  const int Wp = y.Wp(), Hp = y.Hp();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Hp * Wp * CQ) return;
  const int cq = (int)(idx % CQ);
  const int wp = (int)((idx / CQ) % Wp), hp = (int)(idx / ((int64_t)CQ * Wp));
  const int h = min(max(hp - 1, 0), x.H - 1), w = min(max(wp - 1, 0), x.W - 1);
  float rsc[4], rsf[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { rsc[j] = sh[cq * 4 + j]; rsf[j] = sh[C + cq * 4 + j]; }
  const T* src = reinterpret_cast<const T*>(x.p) + x.at(b, 0, h, w) + cq * 4;
  T* dst = reinterpret_cast<T*>(y.p) + (int64_t)b * y.sB + (int64_t)hp * y.sH + (int64_t)wp * y.sW + cq * 4;
  const int64_t xsT = x.sT, ysT = y.sT;
  auto ld = [&](int t) -> uint2 { return __ldg(reinterpret_cast<const uint2*>(src + (int64_t)t * xsT)); };
  auto act = [&](const uint2& q, float* f) {
    Vec4h<T>::get(q, f);
#pragma unroll
    for (int j = 0; j < 4; ++j) f[j] = gn_act<T, SILU, false>(fmaf(f[j], rsc[j], rsf[j]));
  };
  auto put = [&](int plane, const float* f) { *reinterpret_cast<uint2*>(dst + (int64_t)plane * ysT) = Vec4h<T>::set(f); };
  float da[4], db[4];  // f(x[max(2p-1, 0)]), f(x[2p])
  {
    const uint2 q = ld(0);
    act(q, db);
    put(0, db);
#pragma unroll
    for (int j = 0; j < 4; ++j) da[j] = db[j];
  }
  const int P = (Tn - 1) / 2;
  // pair p reads frames 2p + 1 and 2p + 2; a batch is two pairs = frames 2p + 1 .. 2p + 4
  auto load_batch = [&](int p0, uint2 (&q)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = 2 * p0 + 1 + i;
      q[i] = t <= 2 * P ? ld(t) : make_uint2(0u, 0u);
    }
  };
  auto pair = [&](int p, const uint2& q2, const uint2& q3) {
    float d2[4], d3[4], v[4];
    act(q2, d2); act(q3, d3);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = da[j] - d2[j];
    put(1 + 4 * p, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = db[j] + d2[j];
    put(2 + 4 * p, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = d2[j] - db[j];
    put(3 + 4 * p, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = db[j] - d3[j];
    put(4 + 4 * p, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { da[j] = d2[j]; db[j] = d3[j]; }
  };
  uint2 cur[4], nxt[4];
  if (P > 0) load_batch(0, cur);
  for (int p = 0; p < P; p += 2) {
    if (p + 2 < P) load_batch(p + 2, nxt);
    pair(p, cur[0], cur[1]);
    if (p + 1 < P) pair(p + 1, cur[2], cur[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
  }
  if ((Tn & 1) == 0) {  // even T: the last frame alone (three planes)
    const uint2 q = ld(Tn - 1);
    float d2[4], v[4];
    act(q, d2);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = da[j] - d2[j];
    put(1 + 4 * P, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = db[j] + d2[j];
    put(2 + 4 * P, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = d2[j] - db[j];
    put(3 + 4 * P, v);
  }
}

// Sum per-tile GroupNorm partials (written by the conv epilogue) in a fixed order: grid = (groups, B).
// part: [B][rows][groups][2] fp64 -> sums: [B][groups][2] fp64.  One block per (group, batch); threads stride
// over rows, then a fixed-shape tree in shared memory: bit-reproducible.
__global__ void __launch_bounds__(256) gn_finalize_kernel(double* __restrict__ part, int64_t rows, int groups,
                                                          double* __restrict__ sums) {
  __shared__ double sa[256], sq[256];
  const int g = blockIdx.x, b = blockIdx.y;
  double* p = part + ((int64_t)b * rows * groups + g) * 2;
  double a = 0.0, q = 0.0;
  // eight independent 16-byte loads in flight per thread (the rows of one group are 512 B apart: a chain of dependent
  // loads made this tiny kernel ~9 us); the summation order per thread is still row order: bit-reproducible
  constexpr int U = 8;
  for (int64_t r0 = threadIdx.x; r0 < rows; r0 += (int64_t)U * blockDim.x) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + (int64_t)u * blockDim.x;
      v[u] = r < rows ? *reinterpret_cast<const double2*>(p + r * groups * 2) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + (int64_t)u * blockDim.x;
      a += v[u].x; q += v[u].y;
      if (r < rows && (v[u].x != 0.0 || v[u].y != 0.0))   // leave the buffer zeroed for the next conv that accumulates into it
        *reinterpret_cast<double2*>(p + r * groups * 2) = make_double2(0.0, 0.0);
    }
  }
  sa[threadIdx.x] = a; sq[threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sq[threadIdx.x] += sq[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sums[((int64_t)b * groups + g) * 2] = sa[0];
    sums[((int64_t)b * groups + g) * 2 + 1] = sq[0];
  }
}

// ------------------------------------------------------------------------------------------------
// replicate halo of a volume whose INTERIOR was written in place (by a conv epilogue): instead of a pad pass that copies
// the whole tensor (1 read + 1 write), only the halo voxels are written — 4-6 % of the tensor at the big shapes.
// grid = (Tp * Hp destination rows, B); a halo row copies its clamped interior row (Wp voxels), an interior row only its
// 2 * pw edge voxels.  Sources are interior voxels only, so the order of the writes does not matter.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) halo_fill_kernel(Vol y) {
  const int CV = y.C / 8, b = blockIdx.y;
  const int Hp = y.Hp(), Wp = y.Wp();
  const int tp = blockIdx.x / Hp, hp = blockIdx.x - tp * Hp;
  const bool halo_row = tp < y.pt || hp < y.ph || hp >= y.ph + y.H;
  const int t = max(tp - y.pt, 0), h = min(max(hp - y.ph, 0), y.H - 1);
  T* base = reinterpret_cast<T*>(y.p);
  const T* srow = base + y.at(b, t, h, 0);
  T* drow = base + (int64_t)b * y.sB + (int64_t)tp * y.sT + (int64_t)hp * y.sH;
  const int nvox = halo_row ? Wp : 2 * y.pw;
  for (int i = threadIdx.x; i < nvox * CV; i += blockDim.x) {
    const int v = i / CV, cv = i - v * CV;
    const int wp = halo_row ? v : (v < y.pw ? v : y.W + v);  // interior rows: left pw voxels, then the right pw voxels
    const int w = min(max(wp - y.pw, 0), y.W - 1);
    Vec8<T> q; q.load(srow + (int64_t)w * y.sW + cv * 8);
    q.store(drow + (int64_t)wp * y.sW + cv * 8);
  }
}

// ------------------------------------------------------------------------------------------------
// replicate pad / nearest upsample (first frame not upsampled in T)
// ------------------------------------------------------------------------------------------------
// grid = (row chunks, B): a block walks destination rows (tp, hp) of Wp voxels, so the per-vector index math is 32-bit
// shifts / one small division instead of 64-bit div/mod chains, and consecutive threads copy consecutive 16-byte vectors.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) pad_upsample_kernel(Vol x, Vol y, int up_t, int up_h, int up_w, int rows_per_block) {
  const int CV = y.C / VEC;  // y.C == x.C on the vector path; y.C >= x.C (zero channel padding) on the scalar path
  const int b = blockIdx.y;
  const int Wp = y.Wp(), Hp = y.Hp();
  const int nrows = y.Tp() * Hp;
  const int row_elems = Wp * CV;
  const T* xs = reinterpret_cast<const T*>(x.p);
  T* yd = reinterpret_cast<T*>(y.p) + (int64_t)b * y.sB;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, nrows);
  for (int r = r0; r < r1; ++r) {
    const int tp = r / Hp, hp = r - tp * Hp;
    const int t = max(tp - y.pt, 0), h = min(max(hp - y.ph, 0), y.H - 1);
    const int ts = (up_t == 2) ? (t == 0 ? 0 : 1 + ((t - 1) >> 1)) : t;
    const int hs = (up_h == 2) ? (h >> 1) : h;
    const T* srow = xs + x.at(b, ts, hs, 0);
    T* drow = yd + (int64_t)r * Wp * y.C;
    for (int i = threadIdx.x; i < row_elems; i += blockDim.x) {
      const int wp = i / CV, cv = i - wp * CV;
      const int w = min(max(wp - y.pw, 0), y.W - 1);
      const int ws = (up_w == 2) ? (w >> 1) : w;
      const T* sp = srow + (int64_t)ws * x.sW + cv * VEC;
      T* o = drow + (int64_t)i * VEC;
      if (VEC == 8) { Vec8<T> q; q.load(sp); q.store(o); }
      else { o[0] = (cv < x.C) ? sp[0] : from_f<T>(0.f); }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// frame-causal softmax: one block per row.  Row i attends to keys j < (i / n_hw + 1) * n_hw.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  v = (lane < nw) ? sh[lane] : (is_max ? -INFINITY : 0.f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  return v;
}

// One block per row; the visible part of the row (<= L floats) is read from HBM ONCE into shared memory, max / exp /
// sum run on the staged copy, and the probabilities are written with 8-byte stores (the masked tail as zeros).
template <typename T>
__global__ void __launch_bounds__(256) softmax_frame_causal_kernel(const float* __restrict__ S, T* __restrict__ P, int L, int n_hw, float scale) {
  extern __shared__ float row[];  // lim floats
  __shared__ float sh[32];
  const int64_t r = blockIdx.x;  // b*L + i
  const int i = (int)(r % L);
  const int lim = min((i / n_hw + 1) * n_hw, L);
  const float* s = S + r * L;
  T* p = P + r * L;
  const bool vec = (L & 3) == 0 && (lim & 3) == 0;
  float m = -INFINITY;
  if (vec) {
    for (int j = threadIdx.x * 4; j < lim; j += blockDim.x * 4) {
      const float4 v = *reinterpret_cast<const float4*>(s + j);
      *reinterpret_cast<float4*>(row + j) = v;
      m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
  } else {
    for (int j = threadIdx.x; j < lim; j += blockDim.x) { const float v = s[j]; row[j] = v; m = fmaxf(m, v); }
  }
  m = block_reduce(m, sh, true);
  float sum = 0.f;
  for (int j = threadIdx.x; j < lim; j += blockDim.x) { const float e = expf((row[j] - m) * scale); row[j] = e; sum += e; }
  sum = block_reduce(sum, sh, false);
  const float inv = 1.f / sum;
  if (vec && sizeof(T) == 2) {
    for (int j = threadIdx.x * 4; j < L; j += blockDim.x * 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < lim) v = *reinterpret_cast<const float4*>(row + j);
      T o[4] = {from_f<T>(v.x * inv), from_f<T>(v.y * inv), from_f<T>(v.z * inv), from_f<T>(v.w * inv)};
      *reinterpret_cast<uint2*>(p + j) = *reinterpret_cast<const uint2*>(o);
    }
  } else {
    for (int j = threadIdx.x; j < L; j += blockDim.x) p[j] = from_f<T>(j < lim ? row[j] * inv : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// temporal avg-pool (front-replicated) and nearest temporal interpolation
// ------------------------------------------------------------------------------------------------
// MODE 0: avgpool (k, s);  1: nearest interp (inv_scale);  2: linear interp along T = F.interpolate(mode='trilinear',
// align_corners=False) with unit H / W scale: src = max((t + 0.5) * inv_scale - 0.5, 0), lerp of frames floor(src), floor(src)+1
// in fp32;  3: 'area' = adaptive average over [floor(t * Tin / Tout), ceil((t + 1) * Tin / Tout));  4: 'nearest-exact' =
// frame min(floor((t + 0.5) * inv_scale), Tin - 1)   (ATen's UpSample.h index rules for a given scale_factor)
template <typename T, int MODE>
__global__ void temporal_kernel(Vol x, Vol y, int k, int s, float inv_scale) {
  const int C = x.C;
  const int64_t total = (int64_t)y.B * y.T * y.H * y.W * C;
  const T* xs = reinterpret_cast<const T*>(x.p);
  T* yd = reinterpret_cast<T*>(y.p);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); int64_t v = i / C;
    int w = (int)(v % y.W); int64_t r = v / y.W;
    int h = (int)(r % y.H); r /= y.H;
    int t = (int)(r % y.T); int b = (int)(r / y.T);
    float o;
    if (MODE == 0) {
      float acc = 0.f;
      for (int j = 0; j < k; ++j) acc += to_f<T>(xs[x.at(b, max(t * s + j - (k - 1), 0), h, w) + c]);
      o = acc / (float)k;
    } else if (MODE == 1) {
      int ts = min((int)floorf((float)t * inv_scale), x.T - 1);
      o = to_f<T>(xs[x.at(b, ts, h, w) + c]);
    } else if (MODE == 2) {
      const float src = fmaxf(((float)t + 0.5f) * inv_scale - 0.5f, 0.f);
      const int t0 = min((int)src, x.T - 1), t1 = min(t0 + 1, x.T - 1);
      const float l1 = src - (float)t0, l0 = 1.f - l1;
      o = l0 * to_f<T>(xs[x.at(b, t0, h, w) + c]) + l1 * to_f<T>(xs[x.at(b, t1, h, w) + c]);
    } else if (MODE == 3) {
      const int a0 = (int)(((int64_t)t * x.T) / y.T), a1 = (int)((((int64_t)t + 1) * x.T + y.T - 1) / y.T);
      float acc = 0.f;
      for (int j = a0; j < a1; ++j) acc += to_f<T>(xs[x.at(b, j, h, w) + c]);
      o = acc / (float)(a1 - a0);
    } else {
      int ts = min((int)floorf(((float)t + 0.5f) * inv_scale), x.T - 1);
      o = to_f<T>(xs[x.at(b, ts, h, w) + c]);
    }
    yd[y.at(b, t, h, w) + c] = from_f<T>(o);
  }
}

// ------------------------------------------------------------------------------------------------
// blend (in place, raster order is the caller's) + crop + scatter
// Products and the sum are rounded separately, in the tensor dtype, exactly like the reference's
// `a * (1 - y / e) + b * (y / e)` on tensors (autoencoder_kl_causal_3d.py:347,353,359).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float blend2(float a, float b, int y, int e) {
  const float wa = (float)(1.0 - (double)y / (double)e), wb = (float)((double)y / (double)e);
  return rnd<T>(__fadd_rn(rnd<T>(__fmul_rn(a, wa)), rnd<T>(__fmul_rn(b, wb))));
}

// POST: the scattered output is the pipeline tail's image, float((v / 2 + 0.5).clamp(0, 1)) with the sum rounded in the tile
// dtype like the tensor op (pipeline_hunyuan_video.py:1090-1092), written as fp32: the last assembly kernel of a decode
// then emits the final image and no separate post-process pass reads the video back.
template <typename T> __device__ __forceinline__ float post_image(float v) { return fminf(fmaxf(rnd<T>(fmaf(v, 0.5f, 0.5f)), 0.f), 1.f); }

template <typename T, bool POST>
__global__ void blend_crop_scatter_kernel(T* __restrict__ cur, const T* __restrict__ above, const T* __restrict__ left,
                                          int64_t N, int Yc, int Xc, int Ya, int Xl, int ev, int eh,
                                          typename std::conditional<POST, float, T>::type* __restrict__ out, int Yo, int Xo, int y0, int x0,
                                          int crop_y, int crop_x, int64_t cur_ns, int64_t above_ns, int64_t left_ns, int64_t out_ns) {
  const int64_t total = N * Yc * Xc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % Xc); int64_t r = i / Xc;
    int y = (int)(r % Yc); int64_t n = r / Yc;
    const int64_t ci = n * cur_ns + (int64_t)y * Xc + x;
    float v = to_f<T>(cur[ci]);
    bool mod = false;
    if (above != nullptr && y < ev) {
      v = blend2<T>(to_f<T>(above[n * above_ns + (int64_t)(Ya - ev + y) * Xc + x]), v, y, ev);
      mod = true;
    }
    if (left != nullptr && x < eh) {
      v = blend2<T>(to_f<T>(left[n * left_ns + (int64_t)y * Xl + (Xl - eh + x)]), v, x, eh);
      mod = true;
    }
    if (mod) cur[ci] = from_f<T>(v);
    if (out != nullptr && y < crop_y && x < crop_x) {
      if constexpr (POST) out[n * out_ns + (int64_t)(y0 + y) * Xo + (x0 + x)] = post_image<T>(v);
      else out[n * out_ns + (int64_t)(y0 + y) * Xo + (x0 + x)] = from_f<T>(v);
    }
  }
}

// 16-byte form of the kernel above for 16-bit tiles whose rows, windows and strides are multiples of 8 elements (every tile
// of the 720p / 544x960 splits): one thread owns 8 consecutive x of one row.  The blend weights are those of blend2 —
// (float)(1 - y / e) and (float)(y / e) evaluated in double like the reference's Python scalars — tabulated once per block
// in shared memory instead of two fp64 divisions per element; products and sums round exactly as in blend2.  The scalar
// kernel took ~60 us per 65 x 256 x 256 tile (2-byte accesses, 64-bit div / mod per element): 10 ms per 720p step, all of it
// on rank 0 after the gather of a multi-GPU decode.
constexpr int kBlendMaxExtent = 256;
template <typename T, bool POST>
__global__ void __launch_bounds__(256) blend_crop_scatter_vec8_kernel(
    T* __restrict__ cur, const T* __restrict__ above, const T* __restrict__ left, int64_t N, int Yc, int Xc, int Ya, int Xl, int ev, int eh,
    typename std::conditional<POST, float, T>::type* __restrict__ out, int Yo, int Xo, int y0, int x0, int crop_y, int crop_x,
    int64_t cur_ns, int64_t above_ns, int64_t left_ns, int64_t out_ns) {
  __shared__ float wv[2][kBlendMaxExtent], wh[2][kBlendMaxExtent];
  for (int i = threadIdx.x; i < ev; i += blockDim.x) { wv[0][i] = (float)(1.0 - (double)i / (double)ev); wv[1][i] = (float)((double)i / (double)ev); }
  for (int i = threadIdx.x; i < eh; i += blockDim.x) { wh[0][i] = (float)(1.0 - (double)i / (double)eh); wh[1][i] = (float)((double)i / (double)eh); }
  __syncthreads();
  const int Xv = Xc >> 3;
  const int64_t total = N * Yc * Xv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % Xv) * 8; const int64_t r = i / Xv;
    const int y = (int)(r % Yc); const int64_t n = r / Yc;
    const int64_t ci = n * cur_ns + (int64_t)y * Xc + x;
    Vec8<T> q; q.load(cur + ci);
    float v[8]; q.get(v);
    bool mod = false;
    if (above != nullptr && y < ev) {
      Vec8<T> a; a.load(above + n * above_ns + (int64_t)(Ya - ev + y) * Xc + x);
      float af[8]; a.get(af);
      const float wa = wv[0][y], wb = wv[1][y];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = rnd<T>(__fadd_rn(rnd<T>(__fmul_rn(af[j], wa)), rnd<T>(__fmul_rn(v[j], wb))));
      mod = true;
    }
    if (left != nullptr && x < eh) {   // eh % 8 == 0: the whole vector lies inside the strip
      Vec8<T> l; l.load(left + n * left_ns + (int64_t)y * Xl + (Xl - eh + x));
      float lf[8]; l.get(lf);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = rnd<T>(__fadd_rn(rnd<T>(__fmul_rn(lf[j], wh[0][x + j])), rnd<T>(__fmul_rn(v[j], wh[1][x + j]))));
      mod = true;
    }
    if (mod) { q.set(v); q.store(cur + ci); }
    if (out != nullptr && y < crop_y && x < crop_x) {   // crop_x % 8 == 0
      const int64_t oi = n * out_ns + (int64_t)(y0 + y) * Xo + (x0 + x);
      if constexpr (POST) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = post_image<T>(v[j]);
        Vec8<float> ov; ov.set(o); ov.store(out + oi);
      } else {
        if (!mod) q.set(v);
        q.store(out + oi);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pipeline tail: image = (image / 2 + 0.5).clamp(0, 1) in the image dtype, then .float()
// (pipeline_hunyuan_video.py:1090-1092) as ONE pass: 16-bit in, fp32 out.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) image_postprocess_kernel(const T* __restrict__ src, float* __restrict__ dst, int64_t n8, int64_t n) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n8; i += nthr) {  // 8-element vectors (n8 = 0 when a pointer is not 16-byte aligned)
    Vec8<T> q; q.load(src + i * 8);
    float f[8]; q.get(f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = post_image<T>(f[j]);  // x/2 is exact; the sum rounds like the tensor op
    Vec8<float> o; o.set(f); o.store(dst + i * 8);
  }
  for (int64_t i = n8 * 8 + tid; i < n; i += nthr) dst[i] = post_image<T>(to_f<T>(src[i]));  // tail / unaligned: scalar
}

}  // namespace hyvae

using namespace hyvae;

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int hyvae_version(void) { return HYVAE_VERSION; }
const char* hyvae_last_error(void) { return g_err; }
int64_t hyvae_launch_count(void) { return g_launches.load(); }

int hyvae_profile_begin(void) {
  for (auto& r : g_prof) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
  g_prof.clear();
  g_prof_exec_flops = 0.0;
  g_prof_on = true;
  return HYVAE_OK;
}

double hyvae_profile_executed_flops(void) { return g_prof_exec_flops; }

int hyvae_profile_class_override(int32_t cls) {
  HYVAE_CHECK_ARG(cls >= -1 && cls < PC_COUNT, "profile class %d out of range", cls);
  g_prof_class_override = cls;
  return HYVAE_OK;
}

int hyvae_profile_end(double* ms, double* work, int64_t* launches, int32_t n_classes) {
  g_prof_on = false;
  HYVAE_CHECK_ARG(ms && work && launches && n_classes >= PC_COUNT, "profile_end needs %d classes", (int)PC_COUNT);
  for (int i = 0; i < n_classes; ++i) { ms[i] = 0; work[i] = 0; launches[i] = 0; }
  // records may sit on several streams (tile_streams > 1): wait for every event, not only the last one recorded
  for (auto& r : g_prof)
    if (cudaEventSynchronize(r.b) != cudaSuccess) return fail(HYVAE_ECUDA, "profile: event sync failed");
  FILE* dump = nullptr;
  if (const char* path = getenv("HYVAE_PROFILE_DUMP")) dump = fopen(path, "w");
  if (dump) fprintf(dump, "class,tag,work,ms\n");
  for (auto& r : g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return fail(HYVAE_ECUDA, "profile: elapsed time failed");
    ms[r.cls] += t; work[r.cls] += r.work; launches[r.cls] += 1;
    if (dump) fprintf(dump, "%d,%s,%.6g,%.6f\n", r.cls, r.tag, r.work, t);
    g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b);
  }
  g_prof.clear();
  if (dump) fclose(dump);
  return HYVAE_OK;
}

int hyvae_device_supports_tc(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10;
}

int hyvae_ncthw_to_vol(const void* src, int32_t src_dtype, int32_t src_C, const int64_t* ss, const hyvae_vol* dst, void* stream) {
  if (int e = check_vol(dst, "dst")) return e;
  HYVAE_CHECK_ARG(src != nullptr && ss != nullptr, "src is null");
  HYVAE_CHECK_ARG(src_C > 0 && src_C <= dst->C, "src_C=%d must be in [1, dst C=%d]", src_C, dst->C);
  Vol d = make_vol(dst);
  int64_t n = (int64_t)d.B * d.Tp() * d.Hp() * d.Wp();
  ProfScope prof(PC_LAYOUT, (double)n * d.C * (dtype_size(src_dtype) + dtype_size(dst->dtype)), stream);
  HYVAE_DISPATCH_DTYPE(src_dtype, S, HYVAE_DISPATCH_DTYPE(dst->dtype, D,
      (ncthw_to_vol_kernel<S, D><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const S*)src, d, src_C, ss[0], ss[1], ss[2], ss[3], ss[4]))));
  return check_launch("ncthw_to_vol");
}

int hyvae_ncthw_to_vol_kw3(const void* src, int32_t src_dtype, int32_t src_C, const int64_t* ss, const hyvae_vol* dst, void* stream) {
  if (int e = check_vol(dst, "dst")) return e;
  HYVAE_CHECK_ARG(src != nullptr && ss != nullptr, "src is null");
  HYVAE_CHECK_ARG(src_C > 0 && 3 * src_C <= dst->C && dst->C % 8 == 0, "kw-packed layout needs 3*src_C=%d <= dst C=%d, a multiple of 8", 3 * src_C, dst->C);
  Vol d = make_vol(dst);
  int64_t n = (int64_t)d.B * d.Tp() * d.Hp() * d.Wp();
  ProfScope prof(PC_LAYOUT, (double)n * (3.0 * src_C * dtype_size(src_dtype) + d.C * dtype_size(dst->dtype)), stream);
  HYVAE_DISPATCH_DTYPE(src_dtype, S, HYVAE_DISPATCH_DTYPE(dst->dtype, D,
      (ncthw_to_vol_kw3_kernel<S, D><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const S*)src, d, src_C, ss[0], ss[1], ss[2], ss[3], ss[4]))));
  return check_launch("ncthw_to_vol_kw3");
}

int hyvae_vol_to_ncthw(const hyvae_vol* src, void* dst, int32_t dst_dtype, int32_t dst_C, void* stream) {
  if (int e = check_vol(src, "src")) return e;
  HYVAE_CHECK_ARG(dst != nullptr, "dst is null");
  HYVAE_CHECK_ARG(dst_C > 0 && dst_C <= src->C, "dst_C=%d must be in [1, src C=%d]", dst_C, src->C);
  Vol s = make_vol(src);
  int64_t n = (int64_t)s.B * s.T * s.H * s.W;
  ProfScope prof(PC_LAYOUT, (double)n * dst_C * (dtype_size(src->dtype) + dtype_size(dst_dtype)), stream);
  HYVAE_DISPATCH_DTYPE(src->dtype, S, HYVAE_DISPATCH_DTYPE(dst_dtype, D,
      (vol_to_ncthw_kernel<S, D><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(s, (D*)dst, dst_C))));
  return check_launch("vol_to_ncthw");
}

static void gn_stats_plan(const hyvae_vol* x, int* vpb_out, int* nblk_out) {
  const int block = 256, lanes = block / (x->C / 8);
  const int64_t nvox = (int64_t)x->T * x->H * x->W;
  int vpb = lanes * 64;
  // keep at least ~4 blocks per SM when the tensor is large enough
  while (vpb > lanes * 8 && (nvox + vpb - 1) / vpb * x->B < 4 * num_sms()) vpb >>= 1;
  *vpb_out = vpb;
  *nblk_out = (int)((nvox + vpb - 1) / vpb);
}

int64_t hyvae_groupnorm_workspace_bytes(const hyvae_vol* x, int32_t groups) {
  if (x == nullptr || x->C % 8 != 0 || x->C / 8 > 256 || groups <= 0) return -1;
  int vpb, nblk;
  gn_stats_plan(x, &vpb, &nblk);
  return 256 + (int64_t)sizeof(double) * 2 * groups * x->B * nblk;
}

int hyvae_groupnorm_stats(const hyvae_vol* x, int32_t groups, double* sums, void* workspace, int64_t workspace_bytes,
                          void* stream) {
  if (int e = check_vol(x, "x")) return e;
  HYVAE_CHECK_ARG(sums != nullptr && workspace != nullptr, "sums / workspace is null");
  HYVAE_CHECK_ARG(groups > 0 && x->C % groups == 0, "C=%d not divisible by groups=%d", x->C, groups);
  HYVAE_CHECK_ARG(x->C % 8 == 0 && x->C / 8 <= 256, "GroupNorm needs C%%8==0 and C<=2048 (C=%d)", x->C);
  HYVAE_CHECK_ARG(x->B <= 64, "GroupNorm batch %d > 64", x->B);
  HYVAE_CHECK_ARG(workspace_bytes >= hyvae_groupnorm_workspace_bytes(x, groups), "workspace too small");
  Vol v = make_vol(x);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(PC_GN_STATS, (double)x->B * x->T * x->H * x->W * x->C * dtype_size(x->dtype), stream);
  if (cudaMemsetAsync(workspace, 0, 256, st) != cudaSuccess) return fail(HYVAE_ECUDA, "memset tickets failed");
  int vpb, nblk;
  gn_stats_plan(x, &vpb, &nblk);
  const int block = 256, lanes = block / (x->C / 8);
  dim3 grid((unsigned)nblk, (unsigned)x->B);
  size_t smem = sizeof(float) * 2 * x->C * lanes;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(workspace);
  double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256);
  HYVAE_DISPATCH_DTYPE(x->dtype, T, (gn_stats_kernel<T><<<grid, block, smem, st>>>(v, groups, vpb, partials, tickets, sums)));
  return check_launch("groupnorm_stats");
}

int hyvae_groupnorm_apply(const hyvae_vol* x, const double* sums, const float* gamma, const float* beta,
                          int32_t groups, float eps, int32_t silu, int32_t round_like_ref, const hyvae_vol* y,
                          void* stream) {
  if (int e = check_vol(x, "x")) return e;
  if (int e = check_vol(y, "y")) return e;
  HYVAE_CHECK_ARG(sums && gamma && beta, "null parameter");
  HYVAE_CHECK_ARG(x->dtype == y->dtype, "dtype mismatch");
  HYVAE_CHECK_ARG(x->B == y->B && x->T == y->T && x->H == y->H && x->W == y->W && x->C == y->C, "shape mismatch");
  HYVAE_CHECK_ARG(groups > 0 && x->C % groups == 0 && x->C % 8 == 0, "bad C=%d / groups=%d", x->C, groups);
  Vol vx = make_vol(x), vy = make_vol(y);
  char tag[56];
  snprintf(tag, sizeof(tag), "gn C%d %dx%dx%dx%d pad%d%d%d", x->C, x->B, x->T, x->H, x->W, y->pt, y->ph, y->pw);
  ProfScope prof(PC_GN_APPLY, 2.0 * x->B * x->T * x->H * x->W * x->C * dtype_size(x->dtype), stream, tag);
  // rows of Wp*C contiguous elements; give each block >= ~64 KB of output, and keep >= ~8 blocks per SM when possible
  const int nrows = vy.Tp() * vy.Hp();
  const int64_t row_bytes = (int64_t)vy.Wp() * x->C * dtype_size(x->dtype);
  int rpb = (int)((65536 + row_bytes - 1) / row_bytes);
  while (rpb > 1 && (int64_t)((nrows + rpb - 1) / rpb) * x->B < 8 * num_sms()) rpb >>= 1;
  dim3 grid((unsigned)((nrows + rpb - 1) / rpb), (unsigned)x->B);
  size_t smem = sizeof(float) * 2 * x->C;
  const double inv_n = 1.0 / ((double)x->T * x->H * x->W * (x->C / groups));
#define HYVAE_GN_LAUNCH(S, R) \
  HYVAE_DISPATCH_DTYPE(x->dtype, T, (gn_apply_kernel<T, S, R><<<grid, 256, smem, (cudaStream_t)stream>>>(vx, vy, sums, gamma, beta, groups, eps, rpb, inv_n)))
  if (silu) { if (round_like_ref) { HYVAE_GN_LAUNCH(true, true); } else { HYVAE_GN_LAUNCH(true, false); } }
  else { if (round_like_ref) { HYVAE_GN_LAUNCH(false, true); } else { HYVAE_GN_LAUNCH(false, false); } }
#undef HYVAE_GN_LAUNCH
  return check_launch("groupnorm_apply");
}

int hyvae_groupnorm_apply_wino(const hyvae_vol* x, const double* sums, const float* gamma, const float* beta, int32_t groups,
                               float eps, int32_t silu, const hyvae_vol* planes, void* stream) {
  if (int e = check_vol(x, "x")) return e;
  if (int e = check_vol(planes, "planes")) return e;
  HYVAE_CHECK_ARG(sums && gamma && beta, "null parameter");
  HYVAE_CHECK_ARG(x->dtype == planes->dtype && x->dtype != HYVAE_F32, "Winograd planes are 16-bit, in x's dtype");
  const int np = x->T <= 0 ? 0 : 1 + 4 * ((x->T - 1) / 2) + ((x->T % 2 == 0) ? 3 : 0);
  HYVAE_CHECK_ARG(planes->B == x->B && planes->T == np && planes->H == x->H && planes->W == x->W && planes->C == x->C &&
                  planes->pt == 0 && planes->ph == 1 && planes->pw == 1,
                  "planes must be [B][%d][H+2][W+2][C] with halo (0,1,1)", np);
  HYVAE_CHECK_ARG(groups > 0 && x->C % groups == 0 && x->C % 8 == 0 && x->C <= 4096, "bad C=%d / groups=%d", x->C, groups);
  Vol vx = make_vol(x), vy = make_vol(planes);
  char tag[56];
  snprintf(tag, sizeof(tag), "gn C%d %dx%dx%dx%d wino", x->C, x->B, x->T, x->H, x->W);
  // algorithmic bytes as for hyvae_groupnorm_apply (1 read + 1 write of the tensor); the kernel WRITES ~2x that
  ProfScope prof(PC_GN_APPLY, 2.0 * x->B * x->T * x->H * x->W * x->C * dtype_size(x->dtype), stream, tag);
  const int64_t nthreads = (int64_t)vy.Hp() * vy.Wp() * (x->C / 4);
  dim3 grid((unsigned)((nthreads + 255) / 256), (unsigned)x->B);
  const size_t smem = sizeof(float) * 2 * x->C;
  const double inv_n = 1.0 / ((double)x->T * x->H * x->W * (x->C / groups));
  if (x->dtype == HYVAE_BF16) {
    if (silu) gn_apply_wino_kernel<__nv_bfloat16, true><<<grid, 256, smem, (cudaStream_t)stream>>>(vx, vy, x->T, sums, gamma, beta, groups, eps, inv_n);
    else gn_apply_wino_kernel<__nv_bfloat16, false><<<grid, 256, smem, (cudaStream_t)stream>>>(vx, vy, x->T, sums, gamma, beta, groups, eps, inv_n);
  } else {
    if (silu) gn_apply_wino_kernel<__half, true><<<grid, 256, smem, (cudaStream_t)stream>>>(vx, vy, x->T, sums, gamma, beta, groups, eps, inv_n);
    else gn_apply_wino_kernel<__half, false><<<grid, 256, smem, (cudaStream_t)stream>>>(vx, vy, x->T, sums, gamma, beta, groups, eps, inv_n);
  }
  return check_launch("groupnorm_apply_wino");
}

int hyvae_groupnorm_finalize(double* partials, int32_t B, int64_t rows, int32_t groups, double* sums, void* stream) {
  HYVAE_CHECK_ARG(partials && sums && B > 0 && rows > 0 && groups > 0, "bad finalize arguments");
  ProfScope prof(PC_GN_STATS, (double)B * rows * groups * 8, stream);
  gn_finalize_kernel<<<dim3((unsigned)groups, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(partials, rows, groups, sums);
  return check_launch("groupnorm_finalize");
}

int hyvae_halo_fill(const hyvae_vol* y, void* stream) {
  if (int e = check_vol(y, "y")) return e;
  HYVAE_CHECK_ARG(y->dtype != HYVAE_F32 && y->C % 8 == 0, "halo_fill: 16-bit volumes with C %% 8 == 0 (C=%d)", y->C);
  if (y->pt == 0 && y->ph == 0 && y->pw == 0) return HYVAE_OK;
  Vol vy = make_vol(y);
  const int64_t rows = (int64_t)vy.Tp() * vy.Hp();
  HYVAE_CHECK_ARG(rows < (1ll << 31) && y->B <= 65535, "halo_fill: volume too large");
  const double halo_vox = (double)y->B * ((double)vy.Tp() * vy.Hp() * vy.Wp() - (double)y->T * y->H * y->W);
  ProfScope prof(PC_PAD_UPSAMPLE, 2.0 * halo_vox * y->C * dtype_size(y->dtype), stream, "halo_fill");
  dim3 grid((unsigned)rows, (unsigned)y->B);
  if (y->dtype == HYVAE_BF16) halo_fill_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>(vy);
  else halo_fill_kernel<__half><<<grid, 128, 0, (cudaStream_t)stream>>>(vy);
  return check_launch("halo_fill");
}

int hyvae_pad_upsample(const hyvae_vol* x, const hyvae_vol* y, int32_t up_t, int32_t up_h, int32_t up_w, void* stream) {
  if (int e = check_vol(x, "x")) return e;
  if (int e = check_vol(y, "y")) return e;
  HYVAE_CHECK_ARG(x->dtype == y->dtype && x->C <= y->C && x->B == y->B, "dtype/C/B mismatch");
  HYVAE_CHECK_ARG((up_t == 1 || up_t == 2) && (up_h == 1 || up_h == 2) && (up_w == 1 || up_w == 2), "up factors must be 1 or 2");
  int Te = up_t == 2 ? 1 + 2 * (x->T - 1) : x->T;
  HYVAE_CHECK_ARG(y->T == Te && y->H == x->H * up_h && y->W == x->W * up_w, "y dims %dx%dx%d do not match upsampled x", y->T, y->H, y->W);
  Vol vx = make_vol(x), vy = make_vol(y);
  int64_t nvp = (int64_t)vy.B * vy.Tp() * vy.Hp() * vy.Wp();
  ProfScope prof(PC_PAD_UPSAMPLE, (double)nvp * x->C * dtype_size(x->dtype) + (double)x->B * x->T * x->H * x->W * x->C * dtype_size(x->dtype), stream);
  // rows of Wp*C contiguous elements; >= ~32 KB of output per block and >= ~8 blocks per SM when possible
  const int nrows = vy.Tp() * vy.Hp();
  const int64_t row_bytes = (int64_t)vy.Wp() * y->C * dtype_size(x->dtype);
  int rpb = (int)((32768 + row_bytes - 1) / row_bytes);
  while (rpb > 1 && (int64_t)((nrows + rpb - 1) / rpb) * x->B < 8 * num_sms()) rpb >>= 1;
  dim3 grid((unsigned)((nrows + rpb - 1) / rpb), (unsigned)x->B);
  if (x->C % 8 == 0 && x->C == y->C) {
    HYVAE_DISPATCH_DTYPE(x->dtype, T, (pad_upsample_kernel<T, 8><<<grid, 256, 0, (cudaStream_t)stream>>>(vx, vy, up_t, up_h, up_w, rpb)));
  } else {
    HYVAE_DISPATCH_DTYPE(x->dtype, T, (pad_upsample_kernel<T, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(vx, vy, up_t, up_h, up_w, rpb)));
  }
  return check_launch("pad_upsample");
}

int hyvae_softmax_frame_causal(const float* S, void* P, int32_t p_dtype, int32_t B, int32_t L, int32_t n_hw,
                               float scale, void* stream) {
  HYVAE_CHECK_ARG(S && P && B > 0 && L > 0 && n_hw > 0 && L % n_hw == 0, "bad softmax arguments (L=%d n_hw=%d)", L, n_hw);
  ProfScope prof(PC_SOFTMAX, (double)B * L * L * (4 + dtype_size(p_dtype)), stream);
  HYVAE_CHECK_ARG(L <= 48 * 1024, "softmax row of %d floats does not fit the shared-memory staging (max 49152)", L);
  const size_t smem = (size_t)L * sizeof(float);
  HYVAE_DISPATCH_DTYPE(p_dtype, T, {
    static DeviceOnce attr_once;
    if (attr_once.first()) {
      if (cudaFuncSetAttribute(softmax_frame_causal_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024 * 4) != cudaSuccess)
        return fail(HYVAE_ECUDA, "softmax: cannot opt in to 192 KB of shared memory");
      attr_once.done();
    }
    softmax_frame_causal_kernel<T><<<(unsigned)((int64_t)B * L), 256, smem, (cudaStream_t)stream>>>(S, (T*)P, L, n_hw, scale);
  });
  return check_launch("softmax_frame_causal");
}

int hyvae_avgpool_t(const hyvae_vol* x, const hyvae_vol* y, int32_t k, int32_t s, void* stream) {
  if (int e = check_vol(x, "x")) return e;
  if (int e = check_vol(y, "y")) return e;
  HYVAE_CHECK_ARG(k >= 1 && s >= 1, "bad pool k=%d s=%d", k, s);
  HYVAE_CHECK_ARG(y->T == (x->T - 1) / s + 1 && y->H == x->H && y->W == x->W && y->C == x->C && y->B == x->B && x->dtype == y->dtype,
                  "avgpool_t: y shape mismatch");
  Vol vx = make_vol(x), vy = make_vol(y);
  int64_t total = (int64_t)y->B * y->T * y->H * y->W * y->C;
  ProfScope prof(PC_TEMPORAL, (double)total * dtype_size(x->dtype) * 2, stream);
  HYVAE_DISPATCH_DTYPE(x->dtype, T, (temporal_kernel<T, 0><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(vx, vy, k, s, 0.f)));
  return check_launch("avgpool_t");
}

int hyvae_interp_t_nearest(const hyvae_vol* x, const hyvae_vol* y, float inv_scale, void* stream) {
  if (int e = check_vol(x, "x")) return e;
  if (int e = check_vol(y, "y")) return e;
  HYVAE_CHECK_ARG(y->H == x->H && y->W == x->W && y->C == x->C && y->B == x->B && x->dtype == y->dtype, "interp_t: y shape mismatch");
  Vol vx = make_vol(x), vy = make_vol(y);
  int64_t total = (int64_t)y->B * y->T * y->H * y->W * y->C;
  ProfScope prof(PC_TEMPORAL, (double)total * dtype_size(x->dtype) * 2, stream);
  HYVAE_DISPATCH_DTYPE(x->dtype, T, (temporal_kernel<T, 1><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(vx, vy, 0, 0, inv_scale)));
  return check_launch("interp_t_nearest");
}

int hyvae_interp_t(const hyvae_vol* x, const hyvae_vol* y, int32_t mode, float inv_scale, void* stream) {
  if (mode == HYVAE_INTERP_NEAREST) return hyvae_interp_t_nearest(x, y, inv_scale, stream);
  if (int e = check_vol(x, "x")) return e;
  if (int e = check_vol(y, "y")) return e;
  HYVAE_CHECK_ARG(y->H == x->H && y->W == x->W && y->C == x->C && y->B == x->B && x->dtype == y->dtype, "interp_t: y shape mismatch");
  HYVAE_CHECK_ARG(mode == HYVAE_INTERP_LINEAR || mode == HYVAE_INTERP_AREA || mode == HYVAE_INTERP_NEAREST_EXACT, "interp_t: unknown mode %d", mode);
  Vol vx = make_vol(x), vy = make_vol(y);
  int64_t total = (int64_t)y->B * y->T * y->H * y->W * y->C;
  ProfScope prof(PC_TEMPORAL, (double)total * dtype_size(x->dtype) * 2, stream);
  const int g = grid_for(total, 256);
  if (mode == HYVAE_INTERP_LINEAR) { HYVAE_DISPATCH_DTYPE(x->dtype, T, (temporal_kernel<T, 2><<<g, 256, 0, (cudaStream_t)stream>>>(vx, vy, 0, 0, inv_scale))); }
  else if (mode == HYVAE_INTERP_AREA) { HYVAE_DISPATCH_DTYPE(x->dtype, T, (temporal_kernel<T, 3><<<g, 256, 0, (cudaStream_t)stream>>>(vx, vy, 0, 0, inv_scale))); }
  else { HYVAE_DISPATCH_DTYPE(x->dtype, T, (temporal_kernel<T, 4><<<g, 256, 0, (cudaStream_t)stream>>>(vx, vy, 0, 0, inv_scale))); }
  return check_launch("interp_t");
}

int hyvae_image_postprocess(const void* src, int32_t src_dtype, float* dst, int64_t n, void* stream) {
  HYVAE_CHECK_ARG(src && dst && n > 0, "image_postprocess: null pointer or n=%lld", (long long)n);
  HYVAE_CHECK_ARG(src_dtype == HYVAE_BF16 || src_dtype == HYVAE_F16 || src_dtype == HYVAE_F32, "image_postprocess: bad dtype %d", src_dtype);
  const bool aligned = ((uintptr_t)src & (src_dtype == HYVAE_F32 ? 31 : 15)) == 0 && ((uintptr_t)dst & 31) == 0;
  const int64_t n8 = aligned ? n / 8 : 0;
  ProfScope prof(PC_LAYOUT, (double)n * (dtype_size(src_dtype) + 4), stream);
  HYVAE_DISPATCH_DTYPE(src_dtype, T, (image_postprocess_kernel<T><<<grid_for(aligned ? (n + 7) / 8 : n, 256), 256, 0, (cudaStream_t)stream>>>((const T*)src, dst, n8, n)));
  return check_launch("image_postprocess");
}

// Device-to-device copy into PEER memory (a tile pushed into rank 0's arena, vae/tile_parallel.py): one cudaMemcpyAsync on the
// SOURCE device's stream.  Nothing is enqueued on the destination device, whose other processes' contexts would otherwise be
// time-sliced against its owner's kernels (measured at 8 GPUs: rank 0's convs ran 3 % slower with torch's cross-device
// copy_, which records and waits events on the destination device).
int hyvae_peer_copy(void* dst, const void* src, int64_t bytes, void* stream) {
  HYVAE_CHECK_ARG(dst != nullptr && src != nullptr && bytes >= 0, "bad peer copy arguments");
  if (bytes == 0) return HYVAE_OK;
  const cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(HYVAE_ECUDA, "peer copy failed: %s", cudaGetErrorString(e));
  return HYVAE_OK;
}

int hyvae_blend_crop_scatter(void* cur, const void* above, const void* left, int32_t dtype, int64_t N,
                             int32_t Yc, int32_t Xc, int32_t Ya, int32_t Xl, int32_t ev, int32_t eh,
                             void* out, int32_t Yo, int32_t Xo, int32_t y0, int32_t x0, int32_t crop_y,
                             int32_t crop_x, const int64_t* ns, int32_t post, void* stream) {
  const int64_t cur_ns = ns ? ns[0] : (int64_t)Yc * Xc, above_ns = ns ? ns[1] : (int64_t)Ya * Xc;
  const int64_t left_ns = ns ? ns[2] : (int64_t)Yc * Xl, out_ns = ns ? ns[3] : (int64_t)Yo * Xo;
  HYVAE_CHECK_ARG(cur != nullptr && N > 0 && Yc > 0 && Xc > 0, "bad blend arguments");
  HYVAE_CHECK_ARG(above == nullptr || (ev > 0 && ev <= Ya && ev <= Yc), "bad vertical extent %d", ev);
  HYVAE_CHECK_ARG(left == nullptr || (eh > 0 && eh <= Xl && eh <= Xc), "bad horizontal extent %d", eh);
  HYVAE_CHECK_ARG(out == nullptr || (crop_y <= Yc && crop_x <= Xc && y0 + crop_y <= Yo && x0 + crop_x <= Xo && y0 >= 0 && x0 >= 0),
                  "crop window does not fit");
  int64_t total = N * Yc * Xc;
  ProfScope prof(PC_BLEND, (double)total * dtype_size(dtype) + (out ? (double)N * crop_y * crop_x * dtype_size(dtype) : 0.0), stream);
  HYVAE_CHECK_ARG(!post || out != nullptr, "post-processed output requested without an output buffer");
  // 16-byte path: 16-bit tiles with every row length, window and plane stride a multiple of 8 elements and aligned bases
  auto mult8 = [](int64_t v) { return (v & 7) == 0; };
  const bool vec8 = dtype != HYVAE_F32 && mult8(Xc) && mult8(cur_ns) && ((uintptr_t)cur & 15) == 0 &&
                    (above == nullptr || (mult8(above_ns) && ((uintptr_t)above & 15) == 0 && ev <= kBlendMaxExtent)) &&
                    (left == nullptr || (mult8(Xl) && mult8(eh) && mult8(left_ns) && ((uintptr_t)left & 15) == 0 && eh <= kBlendMaxExtent)) &&
                    (out == nullptr || (mult8(Xo) && mult8(x0) && mult8(crop_x) && mult8(out_ns) && ((uintptr_t)out & (post ? 31 : 15)) == 0));
  if (vec8) {
    const int g = grid_for(total / 8, 256);
    if (post) {
      HYVAE_DISPATCH_DTYPE(dtype, T, (blend_crop_scatter_vec8_kernel<T, true><<<g, 256, 0, (cudaStream_t)stream>>>(
          (T*)cur, (const T*)above, (const T*)left, N, Yc, Xc, Ya, Xl, above ? ev : 0, left ? eh : 0, (float*)out, Yo, Xo, y0, x0, crop_y, crop_x,
          cur_ns, above_ns, left_ns, out_ns)));
    } else {
      HYVAE_DISPATCH_DTYPE(dtype, T, (blend_crop_scatter_vec8_kernel<T, false><<<g, 256, 0, (cudaStream_t)stream>>>(
          (T*)cur, (const T*)above, (const T*)left, N, Yc, Xc, Ya, Xl, above ? ev : 0, left ? eh : 0, (T*)out, Yo, Xo, y0, x0, crop_y, crop_x,
          cur_ns, above_ns, left_ns, out_ns)));
    }
    return check_launch("blend_crop_scatter");
  }
  if (post) {
    HYVAE_DISPATCH_DTYPE(dtype, T, (blend_crop_scatter_kernel<T, true><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (T*)cur, (const T*)above, (const T*)left, N, Yc, Xc, Ya, Xl, ev, eh, (float*)out, Yo, Xo, y0, x0, crop_y, crop_x,
        cur_ns, above_ns, left_ns, out_ns)));
  } else {
    HYVAE_DISPATCH_DTYPE(dtype, T, (blend_crop_scatter_kernel<T, false><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        (T*)cur, (const T*)above, (const T*)left, N, Yc, Xc, Ya, Xl, ev, eh, (T*)out, Yo, Xo, y0, x0, crop_y, crop_x,
        cur_ns, above_ns, left_ns, out_ns)));
  }
  return check_launch("blend_crop_scatter");
}

}  // extern "C"
