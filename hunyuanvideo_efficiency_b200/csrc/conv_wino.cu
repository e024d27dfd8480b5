// Stride-1 3x3x3 CausalConv3d as Winograd F(2,3) ALONG T on tcgen05/TMEM: 4 instead of 6 plane-GEMMs per pair of
// output frames, i.e. 1.5x fewer MACs on the layers that are ~75 % of the VAE's step.
//
// Reference semantics: F.pad(x, (1,1,1,1,2,0), 'replicate') + nn.Conv3d(k=3), unet_causal_3d_blocks.py:68-75 (+ the residual
// add of ResnetBlockCausal3D.forward :415).  CPU restatement of the algebra: oracle/winograd.py.
//
// With xp = x preceded by TWO copies of frame 0 (the causal padding) and g_kt = W[:, :, kt]:
//   frame 0:            y[0]     = conv2d(x[0], g0 + g1 + g2)                        (all three frame taps read frame 0)
//   pair p (p >= 0):    d_i = xp[2p + 1 + i] = x[max(2p-1,0)], x[2p], x[2p+1], x[2p+2]
//                       V0 = d0 - d2, V1 = d1 + d2, V2 = d2 - d1, V3 = d1 - d3         (written by gn_apply_wino_kernel)
//                       M_i = conv2d(V_i, U_i),  U = g0, (g0+g1+g2)/2, (g0-g1+g2)/2, g2
//                       y[2p+1] = M0 + M1 + M2,   y[2p+2] = M1 - M2 - M3
//   even T, last frame: y[T-1]   = M0 + M1 + M2 from V0, V1, V2 of (x[T-3], x[T-2], x[T-1])
// Starting the pairs at frame 1 makes an odd T (65, 33, 17, 9: every tile of the 720p split) come out without a tail and
// gives frame 0 its single folded tap group: 2T - 1 plane-GEMMs instead of 3T.
//
// The operand is a PLANE volume [B][NP][H+2][W+2][C] (16-bit, 1-voxel replicate halo in H and W; replicate padding commutes
// with the linear transform) and the weights are five tap groups [5][9][Cout][Cin] (U0..U3, g0+g1+g2).  One work item =
// (time unit, pair of 16x8-voxel m-tiles, 128-channel n-tile); a CTA pair (cta_group::2, M = 256) runs up to four 9-tap GEMMs
// into FOUR TMEM accumulators of 128 columns (all 512 columns) and the epilogue combines them.  TMEM cannot double-buffer
// four accumulators, so the overlap is by accumulator instead: the MMA order is M0, M1, M2, M3; eight epilogue warps (two
// per TMEM lane quarter, 64 columns each) store y_a = M0 + M1 + M2 while M3 runs and release M0 for the next item's first
// GEMM, then store y_b = M1 - M2 - M3 while that GEMM runs.  Operand staging is the halo scheme of conv_halo.cu: one
// {64 ch, 10, 18} halo patch per (plane, 64-channel chunk) feeds all nine (kh, kw) taps.
// Warps (352 threads): 0 = A (plane) TMA producer, 1 = B (weight) TMA producer, 2 = MMA issuer + TMEM allocator,
// 3..10 = epilogue (TMEM lane quarter = warp & 3, column half = (warp - 3) >> 2).
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "conv_internal.h"
#include "tcgen05.cuh"

namespace hyvae {

constexpr int WINO_THREADS = 352;
constexpr int WINO_BN = 128;

// Shared-memory plan.  A B stage holds the THREE kw taps of one kh row (one TMA box {64, 64, 3}): the MMA warp then pays one
// barrier wait, one elect and one commit per 12 MMAs.  With one tap per stage (the first version) its issue loop took ~360
// clocks per stage against 256 clocks of MMA work and the tensor pipe was active 54 % of the time (ncu,
// profiles/r02_ncu_conv_wino_128_v1.txt); the direct kernels hide the same loop behind 8 MMAs per stage.
template <int NA_, int NB_> struct WinoCfgT {
  static constexpr int TWH = 10, THH = 18, PITCH = TWH;
  static constexpr int A_TX = TWH * THH * 128;                      // 23040 bytes per plane stage
  static constexpr int A_BYTES = (A_TX + 1023) / 1024 * 1024;
  static constexpr int NA = NA_;
  static constexpr int TB = 3;                                      // taps per B stage
  static constexpr int B_TAP_BYTES = (WINO_BN / 2) * 128;           // this CTA's half of one tap's weight tile (8 KB)
  static constexpr int B_BYTES = TB * B_TAP_BYTES;
  static constexpr int NB = NB_;
  static constexpr int OUT_BYTES = 2 * 2 * 16384;                   // [y_a, y_b][64-channel half][128 rows x 128 B]
  static constexpr int BAR_BYTES = 1024;
  static constexpr int BIAS_BYTES = 8 * 64 * 4;
  static constexpr int SMEM_BYTES = NA * A_BYTES + NB * B_BYTES + OUT_BYTES + BAR_BYTES + BIAS_BYTES + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(8 * (2 * NA + 2 * NB + 8 + 8) + 8 <= BAR_BYTES, "barrier block");
};
using WinoCfg = WinoCfgT<3, 3>;      // (a <2, 4> split, deeper weight ring, measured the same: 33.59 vs 33.65 frames/s)

__device__ __forceinline__ uint64_t wino_a_desc(uint32_t addr, uint32_t sbo_bytes) {  // see make_halo_desc in conv_halo.cu
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// one work item, decoded identically by every role
struct WinoItem {
  int b, nt, kind;      // kind 0: frame 0 (1 GEMM); 1: pair (4 GEMMs, two output frames); 2: last frame of an even T (3 GEMMs)
  int ngemm;
  int plane0;           // first plane of the item (GEMM i reads plane0 + i)
  int wgroup0;          // weight tap group of GEMM 0 (GEMM i uses wgroup0 + i): 4 for kind 0, else 0
  int ta, tb;           // output frames (tb only for kind 1)
  int h0, w0;           // this CTA's m-tile origin
  bool valid;           // this CTA's m-tile exists (an odd group count leaves the last pair half empty)
};

__device__ __forceinline__ WinoItem wino_decode(const WinoArgs& a, uint32_t item, uint32_t rank) {
  WinoItem r;
  r.nt = (int)(item % (uint32_t)a.n_tiles); item /= (uint32_t)a.n_tiles;
  const uint32_t j = item % (uint32_t)a.upf; item /= (uint32_t)a.upf;
  const int tu = (int)(item % (uint32_t)a.ntu);
  r.b = (int)(item / (uint32_t)a.ntu);
  const uint32_t g = 2 * j + rank;
  r.valid = g < (uint32_t)a.gpf;
  r.h0 = (int)(g / (uint32_t)a.groups_w) * 16;
  r.w0 = (int)(g % (uint32_t)a.groups_w) * 8;
  if (tu == 0) { r.kind = 0; r.ngemm = 1; r.plane0 = 0; r.wgroup0 = 4; r.ta = 0; r.tb = 0; }
  else if (tu <= a.npairs) { r.kind = 1; r.ngemm = 4; r.plane0 = 1 + 4 * (tu - 1); r.wgroup0 = 0; r.ta = 2 * tu - 1; r.tb = 2 * tu; }
  else { r.kind = 2; r.ngemm = 3; r.plane0 = 1 + 4 * a.npairs; r.wgroup0 = 0; r.ta = a.T - 1; r.tb = 0; }
  if (!r.valid) r.b = a.B;  // every TMA box of a missing m-tile lies outside the tensor: zero fill, nothing stored
  return r;
}

// Epilogue of this warp's 32 rows x 64 columns of one output tile: TMEM -> registers -> combine -> + bias (+ residual, already in
// the staging rows) -> 16-bit -> swizzled staging rows; GroupNorm partial sums per lane.  MODE 0: a0; 1: a0 + a1 + a2; 2: a1 - a2 - a3.
template <typename T, int CPG, int MODE>
__device__ __forceinline__ void wino_epi_half(uint32_t tq /* tmem base + lane quarter */, int col0, int lane, uint32_t srow, uint32_t sbias,
                                              bool has_res, bool valid, float (&lacc)[32]) {
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint32_t c = tq + (uint32_t)(col0 + j * 32);
    float f[32];
    if constexpr (MODE == 0) {
      uint32_t v0[32];
      tmem_ld32(c, v0);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v0[e]);
    } else {
      uint32_t v0[32], v1[32], v2[32];
      tmem_ld32(c + (MODE == 1 ? 0 : WINO_BN), v0);
      tmem_ld32(c + (MODE == 1 ? WINO_BN : 2 * WINO_BN), v1);
      tmem_ld32(c + (MODE == 1 ? 2 * WINO_BN : 3 * WINO_BN), v2);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e)
        f[e] = MODE == 1 ? (__uint_as_float(v0[e]) + __uint_as_float(v1[e])) + __uint_as_float(v2[e])
                         : (__uint_as_float(v0[e]) - __uint_as_float(v1[e])) - __uint_as_float(v2[e]);
    }
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const float4 b0 = lds_f4(sbias + (uint32_t)(j * 32 + q4 * 8) * 4), b1 = lds_f4(sbias + (uint32_t)(j * 32 + q4 * 8 + 4) * 4);
      f[q4 * 8 + 0] += b0.x; f[q4 * 8 + 1] += b0.y; f[q4 * 8 + 2] += b0.z; f[q4 * 8 + 3] += b0.w;
      f[q4 * 8 + 4] += b1.x; f[q4 * 8 + 5] += b1.y; f[q4 * 8 + 6] += b1.z; f[q4 * 8 + 7] += b1.w;
      const uint32_t sa16 = srow + ((uint32_t)((j * 4 + q4) ^ (lane & 7)) << 4);
      if (has_res) {
        Vec8<T> r; r.v = lds128(sa16);
        float rf[8]; r.get(rf);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[q4 * 8 + e] += rf[e];
      }
      Vec8<T> o; o.set(&f[q4 * 8]);
      sts128(sa16, o.v);
    }
    if constexpr (CPG > 0) {
#pragma unroll
      for (int g = 0; g < 32 / CPG; ++g) {
        float sm = 0.f, sq = 0.f;
#pragma unroll
        for (int e = 0; e < CPG; ++e) { const float u = valid ? f[g * CPG + e] : 0.f; sm += u; sq = fmaf(u, u, sq); }
        lacc[(j * (32 / CPG) + g) * 2] += sm;
        lacc[(j * (32 / CPG) + g) * 2 + 1] += sq;
      }
    }
  }
}

#define HYVAE_WINO_EPI(T, cpg, MODE, ...)                                   \
  switch (cpg) {                                                            \
    case 0: wino_epi_half<T, 0, MODE>(__VA_ARGS__); break;                  \
    case 4: wino_epi_half<T, 4, MODE>(__VA_ARGS__); break;                  \
    case 8: wino_epi_half<T, 8, MODE>(__VA_ARGS__); break;                  \
    case 16: wino_epi_half<T, 16, MODE>(__VA_ARGS__); break;                \
    default: wino_epi_half<T, 32, MODE>(__VA_ARGS__); break;                \
  }

template <typename T, typename Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WINO_THREADS, 1)
conv_wino_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const WinoArgs a) {
  constexpr int NA = Cfg::NA, NB = Cfg::NB, PITCH = Cfg::PITCH;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + NA * Cfg::A_BYTES;
  const uint32_t sOut = sB + NB * Cfg::B_BYTES;
  const uint32_t bars = sOut + Cfg::OUT_BYTES;
  const uint32_t afull = bars, aempty = afull + 8 * NA;
  const uint32_t bfull = aempty + 8 * NA, bempty = bfull + 8 * NB;
  const uint32_t accfull = bempty + 8 * NB, accempty = accfull + 8 * 4;
  const uint32_t rfull = accempty + 8 * 4;                   // [8 epilogue warps]
  const uint32_t tmem_slot = rfull + 8 * 8;
  const uint32_t sbias_all = bars + Cfg::BAR_BYTES;          // [8 warps][64] fp32
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmY);
    if (a.has_res) tma_prefetch_desc(&tmR);
    if (a.sc_chunks) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW); }
    for (int s = 0; s < NA; ++s) { mbar_init(afull + 8 * s, 2); mbar_init(aempty + 8 * s, 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(bfull + 8 * s, 2); mbar_init(bempty + 8 * s, 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(accfull + 8 * s, 1); mbar_init(accempty + 8 * s, 256 * 2); }
    for (int s = 0; s < 8; ++s) mbar_init(rfull + 8 * s, 1);
    fence_barrier_init();
  }
  cluster_sync_all();  // both CTAs' barriers exist before anything remote touches them
  if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // TMEM of BOTH CTAs is allocated before the leader's first MMA can write it
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int kchunks = a.Cin / 64;
  // Fused 1x1x1 conv_shortcut of the resnet block (unet_causal_3d_blocks.py:338-348,407-415): extra K chunks at the centre
  // tap, reading the block input's halo patch of the output frame.  y_a = M0 + ... takes + Ws * x[t_a] in accumulator 0;
  // y_b = M1 - M2 - M3 takes - Ws * x[t_b] in accumulator 3 (tap 1 of tmW holds the negated weights).
  auto sc_steps = [&](const WinoItem& m, int gi) -> int { return (a.sc_chunks && (gi == 0 || (gi == 3 && m.kind == 1))) ? a.sc_chunks : 0; };
  const uint32_t item0 = blockIdx.x >> 1, istride = gridDim.x >> 1, nitems = (uint32_t)a.total;

  if (warp == 0) {
    // ================= A producer: one {64 ch, 10, 18} plane patch per (GEMM, 64-channel chunk) =================
    int sa = 0; uint32_t pa = 0;
    int afills = 0;
    for (uint32_t it = item0; it < nitems; it += istride) {
      const WinoItem m = wino_decode(a, it, rank);
      for (int gi = 0; gi < m.ngemm; ++gi) {
        const int nsc = sc_steps(m, gi), nst = kchunks + nsc;
        for (int st = 0; st < nst; ++st) {
          const bool sc = st >= kchunks;   // shortcut input: logical (unpadded) coordinates, the 1-voxel rim of the box is never read
          mbar_wait(aempty + 8 * sa, pa ^ 1);
          const bool skipa = (a.probe & 8) && afills >= NA;   // measurement only
          ++afills;
          if (elect_one()) {
            if (leader) { if (skipa) mbar_arrive(afull + 8 * sa); else mbar_expect_tx(afull + 8 * sa, 2 * Cfg::A_TX); }
            if (skipa) {}
            else if (sc) tma_load_5d_2sm(sA + sa * Cfg::A_BYTES, &tmX, afull + 8 * sa, (st - kchunks) * 64, m.w0 - 1, m.h0 - 1, gi == 0 ? m.ta : m.tb, m.b);
            else tma_load_5d_2sm(sA + sa * Cfg::A_BYTES, &tmA, afull + 8 * sa, st * 64, m.w0, m.h0, m.plane0 + gi, m.b);
            if (!leader) mbar_arrive_leader(afull + 8 * sa);
          }
          __syncwarp();
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= B producer: this CTA's 64 rows of the three kw taps of one kh row per stage =================
    int sb = 0; uint32_t pb = 0;
    int bfills = 0;
    for (uint32_t it = item0; it < nitems; it += istride) {
      const WinoItem m = wino_decode(a, it, rank);
      const int n0 = m.nt * WINO_BN + (int)rank * (WINO_BN / 2);
      for (int gi = 0; gi < m.ngemm; ++gi) {
        const int nsc = sc_steps(m, gi);
        for (int st = 0; st < kchunks; ++st) {
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh) {
            mbar_wait(bempty + 8 * sb, pb ^ 1);
            const bool skipb = (a.probe & 1) && bfills >= NB;   // measurement only
            ++bfills;
            if (elect_one()) {
              if (leader) { if (skipb) mbar_arrive(bfull + 8 * sb); else mbar_expect_tx(bfull + 8 * sb, 2 * Cfg::B_BYTES); }
              if (!skipb) tma_load_3d_2sm(sB + sb * Cfg::B_BYTES, &tmB, bfull + 8 * sb, st * 64, n0, (m.wgroup0 + gi) * 9 + kh * 3);
              if (!leader) mbar_arrive_leader(bfull + 8 * sb);
            }
            __syncwarp();
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
        }
        for (int c = 0; c < nsc; ++c) {   // the shortcut has a single tap: + Ws (tap 0) for accumulator 0, - Ws (tap 1) for accumulator 3
          mbar_wait(bempty + 8 * sb, pb ^ 1);
          if (elect_one()) {
            if (leader) mbar_expect_tx(bfull + 8 * sb, 2 * Cfg::B_TAP_BYTES);
            tma_load_3d_2sm(sB + sb * Cfg::B_BYTES, &tmW, bfull + 8 * sb, c * 64, n0, gi == 0 ? 0 : 1);
            if (!leader) mbar_arrive_leader(bfull + 8 * sb);
          }
          __syncwarp();
          if (++sb == NB) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    if (leader) {
      // ================= MMA issuer (leader CTA; warp-uniform loops, one elected lane issues) =================
      // Per B stage: one wait, then 12 MMAs whose descriptors are the stage bases plus COMPILE-TIME offsets (tap (kh, kw) of
      // the halo patch = + (kh * PITCH + kw) rows of 128 B; K = 16 slice k = + 32 B; tap kw of the B stage = + 8 KB).
      constexpr uint32_t idesc = make_idesc_m256(WINO_BN, TcFmt<T>::fmt);
      const uint64_t adesc0 = wino_a_desc(sA, PITCH * 128);
      const uint64_t bdesc0 = make_kmajor_sw128_desc(sB);
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      uint32_t use[4] = {0, 0, 0, 0};  // how often each accumulator has been handed to the epilogue
      for (uint32_t it = item0; it < nitems; it += istride) {
        const WinoItem m = wino_decode(a, it, rank);
#pragma unroll 1
        for (int gi = 0; gi < m.ngemm; ++gi) {
          mbar_wait(accempty + 8 * gi, (use[gi] & 1) ^ 1);  // the epilogue has drained this accumulator's previous contents
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(gi * WINO_BN);
          const int nsc = sc_steps(m, gi);
#pragma unroll 1
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(afull + 8 * sa, pa);
            const uint64_t ad = adesc0 + (uint64_t)((uint32_t)sa * (uint32_t)(Cfg::A_BYTES >> 4));
            const uint32_t first = kc != 0 ? 1u : 0u;   // the very first MMA of a GEMM overwrites the accumulator
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              mbar_wait(bfull + 8 * sb, pb);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)sb * (uint32_t)(Cfg::B_BYTES >> 4));
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16_2sm(d_tmem, ad + (uint64_t)((kh * PITCH + kw) * 8 + 2 * k), bd + (uint64_t)(kw * (Cfg::B_TAP_BYTES >> 4) + 2 * k),
                                 idesc, (kh | kw | k) != 0 ? 1u : first);
                }
                umma_commit_2sm(bempty + 8 * sb);
              }
              __syncwarp();
              if (++sb == NB) { sb = 0; pb ^= 1; }
            }
            if (elect_one()) umma_commit_2sm(aempty + 8 * sa);
            __syncwarp();
            if (++sa == NA) { sa = 0; pa ^= 1; }
          }
#pragma unroll 1
          for (int c = 0; c < nsc; ++c) {   // fused 1x1x1 shortcut: centre tap (kh, kw) = (1, 1) of the block input's halo patch
            mbar_wait(afull + 8 * sa, pa);
            mbar_wait(bfull + 8 * sb, pb);
            tc_fence_after();
            const int rem = a.sc_cin - c * 64;              // K = 16 slices that hold real channels
            const int nk = rem >= 64 ? 4 : (rem + 15) / 16;
            if (elect_one()) {
              const uint64_t ad = adesc0 + (uint64_t)((uint32_t)sa * (uint32_t)(Cfg::A_BYTES >> 4) + (uint32_t)((PITCH + 1) * 8));
              const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)sb * (uint32_t)(Cfg::B_BYTES >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < nk) umma_f16_2sm(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, 1u);
              umma_commit_2sm(bempty + 8 * sb);
              umma_commit_2sm(aempty + 8 * sa);
            }
            __syncwarp();
            if (++sb == NB) { sb = 0; pb ^= 1; }
            if (++sa == NA) { sa = 0; pa ^= 1; }
          }
          if (elect_one()) umma_commit_2sm(accfull + 8 * gi);
          __syncwarp();
          ++use[gi];
        }
      }
    }
  } else {
    // ================= epilogue warps (every CTA: its own m-tile, its own TMEM) =================
    const int q = warp & 3, hf = (warp - 3) >> 2, ew = warp - 3;
    const int hh = 4 * q + (lane >> 3), ww = lane & 7;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t rbar = rfull + 8 * ew;
    const uint32_t sbias = sbias_all + ew * 256;
    const uint32_t stage0 = sOut + hf * 16384 + q * 4096;  // slot s: + s * 32768
    uint32_t rph = 0;
    uint32_t use[4] = {0, 0, 0, 0};
    float lacc[32];  // per-lane GroupNorm partials of this warp's 64 columns: 2 * 64 / cpg <= 32 values (cpg >= 4)
#pragma unroll
    for (int e = 0; e < 32; ++e) lacc[e] = 0.f;
    int gb = -1, gnt = 0, bias_nt = -1;
    const int cpg = a.gn_part ? a.gn_cpg : 0;
    auto gn_flush = [&]() {
      if (a.gn_part == nullptr || gb < 0) return;
      // one halving tree leaves value v (= (group - first group) * 2 + moment) in lane v; the lane adds it to the warp's fp64 row
      float v[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) { v[e] = lacc[e]; lacc[e] = 0.f; }
      const float tot = halving_reduce<32>(v, lane);
      const int grp = (gnt * WINO_BN + hf * 64) / a.gn_cpg + (lane >> 1);
      double* row = a.gn_part + ((int64_t)gb * a.gn_rows + blockIdx.x * 8 + ew) * a.gn_groups * 2;
      if (lane < 2 * 64 / a.gn_cpg && grp < a.gn_groups) row[grp * 2 + (lane & 1)] += (double)tot;  // private slot: plain RMW
    };
    auto arrive_empty = [&](int i) {
      if (leader) mbar_arrive(accempty + 8 * i); else mbar_arrive_leader(accempty + 8 * i);
    };
    for (uint32_t it = item0; it < nitems; it += istride) {
      const WinoItem m = wino_decode(a, it, rank);
      const bool live = m.valid;
      if (live && (m.b != gb || m.nt != gnt)) { gn_flush(); gb = m.b; gnt = m.nt; }
      if (m.nt != bias_nt) {  // this warp's 64 bias values (zero padded)
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int n = m.nt * WINO_BN + hf * 64 + h2 * 32 + lane;
          const float bv = (a.bias != nullptr && n < a.Cout) ? a.bias[n] : 0.f;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + (uint32_t)(h2 * 32 + lane) * 4), "f"(bv) : "memory");
        }
        __syncwarp();
        bias_nt = m.nt;
      }
      const int n0 = m.nt * WINO_BN + hf * 64;
      const bool valid = live && (m.h0 + hh) < a.Ho && (m.w0 + ww) < a.Wo;
      // Each slot commits exactly one bulk group per item (an empty one if nothing is stored), so when a slot is acquired
      // the group that last read it is the second newest: one newer group may still be pending.
      auto stage_acquire = [&](int slot, int t) {
        if (lane == 0) {
          bulk_wait_read<1>();
          if (a.has_res && live) {
            mbar_expect_tx(rbar, 4096);
            tma_load_5d(stage0 + slot * 32768, &tmR, rbar, n0, m.w0, m.h0 + 4 * q, t, m.b);
          }
        }
        __syncwarp();
      };
      // ---- first output frame of the item: a0 (frame 0) or a0 + a1 + a2
      stage_acquire(0, m.ta);
      const int nfirst = m.kind == 0 ? 1 : 3;
      for (int i = 0; i < nfirst; ++i) mbar_wait(accfull + 8 * i, use[i] & 1);
      tc_fence_after();
      if (live) {
        if (a.has_res) { mbar_wait(rbar, rph); rph ^= 1u; }
        if (m.kind == 0) { HYVAE_WINO_EPI(T, cpg, 0, tq, hf * 64, lane, stage0 + lane * 128, sbias, a.has_res != 0, valid, lacc) }
        else { HYVAE_WINO_EPI(T, cpg, 1, tq, hf * 64, lane, stage0 + lane * 128, sbias, a.has_res != 0, valid, lacc) }
      }
      tc_fence_before();
      arrive_empty(0); ++use[0];                       // M0 is only read here: the next item's first GEMM may start
      if (m.kind != 1) {                               // no second frame: everything is drained
        for (int i = 1; i < nfirst; ++i) { arrive_empty(i); ++use[i]; }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (live && n0 < a.Cout) tma_store_5d(&tmY, stage0, n0, m.w0, m.h0 + 4 * q, m.ta, m.b);
        bulk_commit();
      }
      // ---- second output frame of a pair: a1 - a2 - a3
      if (m.kind == 1) {
        stage_acquire(1, m.tb);
        mbar_wait(accfull + 8 * 3, use[3] & 1);
        tc_fence_after();
        if (live) {
          if (a.has_res) { mbar_wait(rbar, rph); rph ^= 1u; }
          HYVAE_WINO_EPI(T, cpg, 2, tq, hf * 64, lane, stage0 + 32768 + lane * 128, sbias, a.has_res != 0, valid, lacc)
        }
        tc_fence_before();
        for (int i = 1; i < 4; ++i) { arrive_empty(i); ++use[i]; }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (live && n0 < a.Cout) tma_store_5d(&tmY, stage0 + 32768, n0, m.w0, m.h0 + 4 * q, m.tb, m.b);
          bulk_commit();
        }
      } else if (lane == 0) {
        bulk_commit();                                 // keep one group per slot and item
      }
    }
    gn_flush();
    if (lane == 0) bulk_wait0();
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  if (warp == 2) { tc_fence_after(); tmem_dealloc_2sm<512>(tmem_base); }

  // ---- GroupNorm statistics finished in the kernel (replaces a hyvae_groupnorm_finalize launch per conv) -----------------
  // Every CTA adds its eight warp rows (fixed order) into ITS row of the per-CTA buffer and re-zeroes them; the CTA whose
  // ticket is the last one adds the CTA rows in index order: same bits whatever the arrival order.
  if (a.gn_sums != nullptr) {
    volatile int* s_last = reinterpret_cast<volatile int*>(gen_base + (tmem_slot - smem_base) + 8);
    double* scratch = reinterpret_cast<double*>(gen_base + (sOut - smem_base));   // the staging tiles are idle by now
    const int G2 = a.gn_groups * 2, nval = a.B * G2;
    for (int idx = threadIdx.x; idx < nval; idx += blockDim.x) {
      const int b = idx / G2, v = idx - b * G2;
      double* p = a.gn_part + ((int64_t)b * a.gn_rows + (int64_t)blockIdx.x * 8) * G2 + v;
      double r[8];
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) r[w8] = __ldcg(p + (int64_t)w8 * G2);
      double acc = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) { acc += r[w8]; if (r[w8] != 0.0) p[(int64_t)w8 * G2] = 0.0; }
      a.gn_cta[((int64_t)b * a.gn_cta_rows + blockIdx.x) * G2 + v] = acc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *s_last = (atomicAdd(a.gn_ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (*s_last) {
      __threadfence();
      constexpr int PARTS = 5;                      // 5 x 64 = 320 of the 352 threads for the usual [1][32][2]
      const int nc = (int)gridDim.x, per = (nc + PARTS - 1) / PARTS;
      for (int base = 0; base < nval; base += 64) {
        const int v = base + (int)(threadIdx.x & 63), part = (int)(threadIdx.x >> 6);
        if (part < PARTS && v < nval) {
          const int b = v / G2, vv = v - b * G2;
          const double* q = a.gn_cta + ((int64_t)b * a.gn_cta_rows) * G2 + vv;
          double acc = 0.0;
          const int c1 = min(nc, (part + 1) * per);
          for (int c = part * per; c < c1; ++c) acc += __ldcg(q + (int64_t)c * G2);
          scratch[part * 64 + (threadIdx.x & 63)] = acc;
        }
        __syncthreads();
        if (threadIdx.x < 64 && v < nval) {
          double acc = 0.0;
#pragma unroll
          for (int pp = 0; pp < PARTS; ++pp) acc += scratch[pp * 64 + threadIdx.x];
          a.gn_sums[v] = acc;
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) *a.gn_ticket = 0u;     // ready for the next launch on this stream
    }
  }
}

template <typename T, typename Cfg>
static int launch_wino_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
                         const CUtensorMap& tmX, const CUtensorMap& tmW, const WinoArgs& a, cudaStream_t stream) {
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    if (cudaFuncSetAttribute(conv_wino_kernel<T, Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
      return fail(HYVAE_ECUDA, "conv_wino: cannot opt in to %d bytes of shared memory", Cfg::SMEM_BYTES);
    attr_once.done();
  }
  const int64_t max_pairs = conv_sms() / 2;
  const int64_t pairs = a.total < max_pairs ? a.total : max_pairs;
  conv_wino_kernel<T, Cfg><<<(unsigned)(2 * pairs), WINO_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmY, tmR, tmX, tmW, a);
  return check_launch("conv3d_causal_wino");
}

int launch_wino(int dtype, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
                const CUtensorMap& tmX, const CUtensorMap& tmW, const WinoArgs& a, cudaStream_t stream) {
  if (dtype == HYVAE_BF16) return launch_wino_t<__nv_bfloat16, WinoCfg>(tmA, tmB, tmY, tmR, tmX, tmW, a, stream);
  return launch_wino_t<__half, WinoCfg>(tmA, tmB, tmY, tmR, tmX, tmW, a, stream);
}

}  // namespace hyvae

using namespace hyvae;

typedef CUresult (*WinoEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WinoEncodeFn wino_encode_fn() {
  static WinoEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WinoEncodeFn>(p);
  }
  return fn;
}

extern "C" int32_t hyvae_wino_planes(int32_t T) { return T <= 0 ? 0 : 1 + 4 * ((T - 1) / 2) + ((T % 2 == 0) ? 3 : 0); }

extern "C" int hyvae_conv3d_causal_wino(const hyvae_vol* planes, int32_t T, const void* uw, const float* bias, const hyvae_vol* residual,
                                        const hyvae_vol* sc_x, const void* sc_w, const hyvae_vol* y, double* gn_partials,
                                        int32_t gn_groups, double* gn_sums, void* stream) {
  if (int e = check_vol(planes, "planes")) return e;
  if (int e = check_vol(y, "y")) return e;
  HYVAE_CHECK_ARG(uw != nullptr, "uw is null");
  HYVAE_CHECK_ARG((planes->dtype == HYVAE_BF16 || planes->dtype == HYVAE_F16) && planes->dtype == y->dtype, "Winograd conv needs 16-bit planes and y of one dtype");
  HYVAE_CHECK_ARG(T >= 1 && planes->T == hyvae_wino_planes(T) && planes->pt == 0 && planes->ph == 1 && planes->pw == 1,
                  "planes must be [B][%d][H+2][W+2][C] (halo (0,1,1)) for T=%d; got T=%d halo (%d,%d,%d)", hyvae_wino_planes(T), T, planes->T,
                  planes->pt, planes->ph, planes->pw);
  HYVAE_CHECK_ARG(y->B == planes->B && y->T == T && y->H == planes->H && y->W == planes->W, "y dims do not match the conv output");
  if (planes->C % 64 != 0 || y->C % WINO_BN != 0)
    return fail(HYVAE_EUNSUPPORTED, "Winograd conv takes Cin %% 64 == 0 and Cout %% 128 == 0 (Cin=%d Cout=%d)", planes->C, y->C);
  HYVAE_CHECK_ARG(((uintptr_t)planes->data & 15) == 0 && ((uintptr_t)uw & 15) == 0 && ((uintptr_t)y->data & 15) == 0, "pointers must be 16-byte aligned");
  WinoEncodeFn encode = wino_encode_fn();
  if (!encode) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");

  Vol vp = make_vol(planes), vy = make_vol(y);
  WinoArgs a;
  a.bias = bias; a.B = y->B; a.T = T; a.Ho = y->H; a.Wo = y->W; a.Cin = planes->C; a.Cout = y->C;
  a.tiles_h = (y->H + 15) / 16; a.groups_w = (y->W + 7) / 8;
  a.gpf = a.tiles_h * a.groups_w; a.upf = (a.gpf + 1) / 2;
  a.npairs = (T - 1) / 2; a.ntu = 1 + a.npairs + ((T % 2 == 0) ? 1 : 0);
  a.n_tiles = y->C / WINO_BN;
  const int64_t total = (int64_t)y->B * a.ntu * a.upf * a.n_tiles;
  HYVAE_CHECK_ARG(total < (1ll << 31), "too many work items");
  a.total = (int)total;
  a.has_res = residual != nullptr;
  a.sc_cin = sc_x ? sc_x->C : 0; a.sc_chunks = (a.sc_cin + 63) / 64;
  if (sc_x != nullptr) {
    if (int e = check_vol(sc_x, "sc_x")) return e;
    HYVAE_CHECK_ARG(sc_w != nullptr && residual == nullptr, "fused shortcut: sc_w is null, or a residual was passed as well");
    HYVAE_CHECK_ARG(sc_x->dtype == y->dtype && sc_x->B == y->B && sc_x->T == y->T && sc_x->H == y->H && sc_x->W == y->W && sc_x->C % 8 == 0,
                    "shortcut input must have y's extent and dtype and C %% 8 == 0");
    HYVAE_CHECK_ARG(((uintptr_t)sc_x->data & 15) == 0 && ((uintptr_t)sc_w & 15) == 0, "pointers must be 16-byte aligned");
  }
  a.gn_part = gn_partials; a.gn_groups = gn_groups; a.gn_cpg = 0; a.gn_rows = gn_partial_rows();
  a.gn_sums = nullptr; a.gn_cta = nullptr; a.gn_ticket = nullptr; a.gn_cta_rows = num_sms();
  { const char* pe = getenv("HYVAE_TC_PROBE"); a.probe = pe ? atoi(pe) : 0; }  // measurement only: results are garbage when set
  if (gn_partials && gn_sums) {   // the buffer has the hyvae_gn_partials_doubles() layout: warp rows | CTA rows | ticket
    HYVAE_CHECK_ARG(y->B * gn_groups * 2 <= 4096, "fused GroupNorm finalize: B * groups too large");
    a.gn_sums = gn_sums;
    a.gn_cta = gn_partials + gn_warp_rows_doubles(y->B, gn_groups);
    a.gn_ticket = reinterpret_cast<unsigned int*>(a.gn_cta + gn_cta_rows_doubles(y->B, gn_groups));
  }
  if (gn_partials) {
    HYVAE_CHECK_ARG(gn_groups > 0 && y->C % gn_groups == 0, "gn_groups=%d does not divide Cout=%d", gn_groups, y->C);
    a.gn_cpg = y->C / gn_groups;
    HYVAE_CHECK_ARG(a.gn_cpg == 4 || a.gn_cpg == 8 || a.gn_cpg == 16 || a.gn_cpg == 32,
                    "Winograd conv: fused GroupNorm statistics need Cout/groups in {4,8,16,32} (got %d)", a.gn_cpg);
  }
  const CUtensorMapDataType dt = planes->dtype == HYVAE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmB, tmY, tmR;
  {
    cuuint64_t dims[5] = {(cuuint64_t)planes->C, (cuuint64_t)vp.Wp(), (cuuint64_t)vp.Hp(), (cuuint64_t)vp.Tp(), (cuuint64_t)planes->B};
    cuuint64_t strides[4] = {(cuuint64_t)vp.sW * 2, (cuuint64_t)vp.sH * 2, (cuuint64_t)vp.sT * 2, (cuuint64_t)vp.sB * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)WinoCfg::TWH, (cuuint32_t)WinoCfg::THH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmA, dt, 5, planes->data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(A planes) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)planes->C, (cuuint64_t)y->C, 45};
    cuuint64_t strides[2] = {(cuuint64_t)planes->C * 2, (cuuint64_t)planes->C * y->C * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)(WINO_BN / 2), (cuuint32_t)WinoCfg::TB};   // the three kw taps of one kh row
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmB, dt, 3, const_cast<void*>(uw), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  auto out_map = [&](CUtensorMap* tm, const hyvae_vol* v, const Vol& vv) -> int {
    cuuint64_t dims[5] = {(cuuint64_t)v->C, (cuuint64_t)v->W, (cuuint64_t)v->H, (cuuint64_t)v->T, (cuuint64_t)v->B};
    cuuint64_t strides[4] = {(cuuint64_t)vv.sW * 2, (cuuint64_t)vv.sH * 2, (cuuint64_t)vv.sT * 2, (cuuint64_t)vv.sB * 2};
    cuuint32_t box[5] = {64, 8, 4, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    void* base = (char*)v->data + vv.at(0, 0, 0, 0) * 2;
    CUresult r = encode(tm, dt, 5, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(out) failed with %d", (int)r);
  };
  if (int e = out_map(&tmY, y, vy)) return e;
  if (residual) {
    if (int e = check_vol(residual, "residual")) return e;
    HYVAE_CHECK_ARG(residual->dtype == y->dtype && residual->B == y->B && residual->T == y->T && residual->H == y->H &&
                    residual->W == y->W && residual->C == y->C, "residual shape mismatch");
    Vol vr = make_vol(residual);
    if (int e = out_map(&tmR, residual, vr)) return e;
  } else {
    tmR = tmY;
  }
  CUtensorMap tmX = tmA, tmW = tmB;
  if (sc_x) {
    Vol vs = make_vol(sc_x);
    cuuint64_t dims[5] = {(cuuint64_t)sc_x->C, (cuuint64_t)sc_x->W, (cuuint64_t)sc_x->H, (cuuint64_t)sc_x->T, (cuuint64_t)sc_x->B};
    cuuint64_t strides[4] = {(cuuint64_t)vs.sW * 2, (cuuint64_t)vs.sH * 2, (cuuint64_t)vs.sT * 2, (cuuint64_t)vs.sB * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)WinoCfg::TWH, (cuuint32_t)WinoCfg::THH, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmX, dt, 5, (char*)sc_x->data + vs.at(0, 0, 0, 0) * 2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(shortcut x) failed with %d", (int)r);
    cuuint64_t wd[3] = {(cuuint64_t)sc_x->C, (cuuint64_t)y->C, 2};   // [+Ws, -Ws]
    cuuint64_t ws[2] = {(cuuint64_t)sc_x->C * 2, (cuuint64_t)sc_x->C * y->C * 2};
    cuuint32_t wb[3] = {64, (cuuint32_t)(WINO_BN / 2), 1};
    cuuint32_t we[3] = {1, 1, 1};
    r = encode(&tmW, dt, 3, const_cast<void*>(sc_w), wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(shortcut w) failed with %d", (int)r);
  }
  char tag[56];
  snprintf(tag, sizeof(tag), "k3 %d->%d %dx%dx%dx%d s111 wino%s", planes->C, y->C, y->B, y->T, y->H, y->W, sc_x ? "+sc" : "");
  const double vox1 = (double)y->B * y->H * y->W;
  const double sc_work = sc_x ? 2.0 * vox1 * T * y->C * sc_x->C : 0.0;
  const double work = 2.0 * vox1 * T * y->C * planes->C * 27.0 + sc_work;                      // what the reference executes
  const double executed = 2.0 * vox1 * hyvae_wino_planes(T) * y->C * planes->C * 9.0 + sc_work;  // one 9-tap GEMM per plane
  ProfScope prof(PC_CONV_TC, work, stream, tag, executed);
  return launch_wino(planes->dtype, tmA, tmB, tmY, tmR, tmX, tmW, a, (cudaStream_t)stream);
}
