// CausalConv3d as an implicit GEMM on the sm_100a tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Reference semantics: F.pad(replicate) + nn.Conv3d, unet_causal_3d_blocks.py:73-75 (+ residual :415).
//
// GEMM view (NT):  D[m][n] = sum_{tap, c} A_tap[m][c] * W[tap][n][c]
//   m  = one output voxel of a TH x TW spatial patch of one output frame (TH*TW = 128 = UMMA M)
//   n  = output channel (BN per CTA tile, UMMA N)
//   k  = (tap, 64-channel chunk): one TMA box each for A and B per pipeline stage
// A operand: the input volume already carries the causal replicate halo (hyvae_vol pt/ph/pw), so the
//   A tile of tap (kt,kh,kw) is the plain 5-D TMA box {64 ch, TW, TH, 1, 1} at
//   (c0, w0*sw+kw, h0*sh+kh, t*st+kt, b): 128 rows of 128 bytes, SWIZZLE_128B — exactly the K-major
//   UMMA shared-memory layout.  Strided convs use the tensor map's elementStrides; rows / columns
//   beyond the volume and channels beyond Cin are zero-filled by TMA.
// B operand: weights repacked to [tap][Cout][Cin]; 3-D box {64, BN, 1}.
// D: fp32 accumulators in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps
//   the main loop of tile i+1.  Epilogue: tcgen05.ld -> +bias (+residual) -> bf16/fp16 -> global.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
//   warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).  Persistent CTAs, static tile schedule
//   with the n-tile fastest so concurrent CTAs share the A halo in L2.
#include <cuda.h>

#include <type_traits>

#include <cstdlib>

#include "common.cuh"
#include "conv_internal.h"
#include "tcgen05.cuh"

namespace hyvae {

// ---------------------------------------------------------------------------------- kernel
struct TcArgs {
  void* y;
  const void* res;
  const float* bias;
  int64_t ysB, ysT, ysH, ysW, yoff;  // element strides / offset of logical (0,0,0,0) in y
  int64_t rsB, rsT, rsH, rsW, roff;
  int B, To, Ho, Wo, Cin, Cout;
  int k, st, sh, sw;
  int TH, TW, tiles_h, tiles_w, n_tiles;
  int64_t m_tiles;      // B * To * tiles_h * tiles_w
  int64_t total_tiles;  // ceil(m_tiles / MT) * n_tiles
  int round_like_ref;
  double* gn_part;      // optional [B][gn_rows][gn_groups][2] (sum, sum of squares) of the output; row = CTA * 4 + warp
  int gn_groups, gn_cpg, gn_rows;
  int64_t m_tiles_per_b;
  int probe;            // measurement only (HYVAE_TC_PROBE): bit 0 = stop issuing TMA once the ring is primed, bit 1 = all loads hit tile 0, bit 2 = skip the epilogue
  // kh-trick pair kernel: tap geometry.  The standard conv is nkt = nkw = nsub = 3 with origin (0,0,0); one phase of
  // the sub-pixel decomposition of nearest-upsample + conv (hyvae_conv3d_upphase_tc) has 2 (or 3) x 2 x 2 taps and a
  // box origin shifted by the phase.  Weight tap index = (kt * nsub + kh) * nkw + kw.
  int nkt, nkw, nsub, ot, oh, ow, a_tx;
  int halo2d;             // kh-trick pair kernel, sub-pixel phases: ONE {64 ch, TW + nkw - 1, TH + nsub - 1} halo stage per (kt, chunk) feeds all (kh, kw) taps
  int sc_chunks, sc_cin;  // kh-trick pair kernel: fused 1x1x1 shortcut conv (64-channel chunks / channels of its input)
  int tfold;              // kh-trick pair kernel, standard 3x3x3 taps: weights carry the folded first-frame taps (tfold_class)
};

constexpr int TC_THREADS = 192;
constexpr int TC2_KHT_THREADS = 224;   // kh-trick pair kernel: warp 6 = B (weight) TMA producer, warp 0 loads only the A halo stages
constexpr int A_STAGE_BYTES = 128 * 64 * 2;

// MT = number of 128-voxel m-tiles a CTA multiplies against one B (weight) tile per stage.  MT = 2 halves
// the shared-memory fill per MMA cycle for the narrow-N layers (Cout <= 128).
template <int BN, int MT> struct TcCfg {
  static constexpr int A_BYTES = MT * A_STAGE_BYTES;
  static constexpr int B_STAGE_BYTES = BN * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_COLS = MT * BN;  // fp32 accumulator columns per pipeline stage
  static constexpr int TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two <= 512");
};


// GroupNorm partial statistics of one 32-column chunk of the epilogue: per group of CPG channels, the sum and sum of
// squares over this warp's 32 rows (fixed shuffle tree); lane 0 adds them to the fp64 slot that this (CTA, warp)
// owns for the tile's batch item.  The tile order of a CTA is static, so the result is bit-reproducible.
template <int CPG>
__device__ __forceinline__ void gn_chunk_stats(const float* f, bool valid, double* dst, int lane, int ngroups_valid) {
#pragma unroll
  for (int g = 0; g < 32 / CPG; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) { const float u = valid ? f[g * CPG + c] : 0.f; s += u; q = fmaf(u, u, q); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (lane == 0 && g < ngroups_valid) { dst[2 * g] += (double)s; dst[2 * g + 1] += (double)q; }  // private slot: plain RMW
  }
}

struct MTile { int b, t, h0, w0; bool valid; };
__device__ __forceinline__ MTile decode_mtile(const TcArgs& a, int64_t mt) {
  MTile r;
  r.valid = mt < a.m_tiles;
  const int tw = (int)(mt % a.tiles_w); mt /= a.tiles_w;
  const int th = (int)(mt % a.tiles_h); mt /= a.tiles_h;
  r.t = (int)(mt % a.To); r.b = (int)(mt / a.To);  // an invalid tile gets b >= B: TMA zero-fills it
  r.h0 = th * a.TH; r.w0 = tw * a.TW;
  return r;
}

// first-frame fold class (tcgen05.cuh tfold_class) of the m-tile pair (2 mg, 2 mg + 1) of the CTA-pair kernel: 32-bit
// divisions only, because the MMA-issuing warp evaluates it per tile
__device__ __forceinline__ int tile_fold_class(const TcArgs& a, int64_t mg) {
  const uint32_t per_frame = (uint32_t)(a.tiles_h * a.tiles_w);
  const uint32_t i0 = (uint32_t)(2 * mg), i1 = i0 + 1;
  return tfold_class(1, (int)((i0 / per_frame) % (uint32_t)a.To), (int64_t)i0 < a.m_tiles, (int)((i1 / per_frame) % (uint32_t)a.To), (int64_t)i1 < a.m_tiles);
}

// Epilogue of one 128-voxel m-tile x BN channels held by this CTA: TMEM -> registers -> +bias (+residual) ->
// store, plus the optional GroupNorm partial statistics.  Called by the 4 epilogue warps (q = TMEM lane quarter).
template <typename T, typename OT, int BN>
__device__ __forceinline__ void epilogue_mtile(const TcArgs& a, const MTile& m, int q, int lane, uint32_t tmem_cols, int n0) {
  const int row = q * 32 + lane;
  OT* yd = reinterpret_cast<OT*>(a.y);
  const T* rs = reinterpret_cast<const T*>(a.res);
  const int h = m.h0 + row / a.TW, w = m.w0 + row % a.TW;
  const bool valid = (h < a.Ho) && (w < a.Wo);
  const int64_t yo = a.yoff + (int64_t)m.b * a.ysB + (int64_t)m.t * a.ysT + (int64_t)h * a.ysH + (int64_t)w * a.ysW;
  const int64_t ro = a.roff + (int64_t)m.b * a.rsB + (int64_t)m.t * a.rsT + (int64_t)h * a.rsH + (int64_t)w * a.rsW;
#pragma unroll 1
  for (int j = 0; j < BN / 32; ++j) {
    uint32_t v[32];
    tmem_ld32(tmem_cols + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 32), v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
    const int nc = n0 + j * 32;
    if (nc < a.Cout) {  // warp-uniform
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n = nc + g * 8;
        if (n < a.Cout) {
          if (a.bias) {
            const float4 b0 = *reinterpret_cast<const float4*>(a.bias + n);
            const float4 b1 = *reinterpret_cast<const float4*>(a.bias + n + 4);
            f[g * 8 + 0] += b0.x; f[g * 8 + 1] += b0.y; f[g * 8 + 2] += b0.z; f[g * 8 + 3] += b0.w;
            f[g * 8 + 4] += b1.x; f[g * 8 + 5] += b1.y; f[g * 8 + 6] += b1.z; f[g * 8 + 7] += b1.w;
          }
          if (rs && valid) {
            Vec8<T> r; r.load(rs + ro + n);
            float rf[8]; r.get(rf);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[g * 8 + e] = (a.round_like_ref ? rnd<T>(f[g * 8 + e]) : f[g * 8 + e]) + rf[e];
          }
          if (valid) {
            Vec8<OT> o; o.set(&f[g * 8]);
            o.store(yd + yo + n);
          }
        }
      }
      if (a.gn_part) {
        const int64_t prow = (int64_t)m.b * a.gn_rows + blockIdx.x * 4 + q;
        double* dst = a.gn_part + (prow * a.gn_groups + nc / a.gn_cpg) * 2;
        const int ng = (a.Cout - nc) / a.gn_cpg;
        switch (a.gn_cpg) {
          case 1: gn_chunk_stats<1>(f, valid, dst, lane, ng); break;
          case 2: gn_chunk_stats<2>(f, valid, dst, lane, ng); break;
          case 4: gn_chunk_stats<4>(f, valid, dst, lane, ng); break;
          case 8: gn_chunk_stats<8>(f, valid, dst, lane, ng); break;
          case 16: gn_chunk_stats<16>(f, valid, dst, lane, ng); break;
          default: gn_chunk_stats<32>(f, valid, dst, lane, ng); break;
        }
      }
    }
  }
}

template <typename T, typename OT, int BN, int MT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  using Cfg = TcCfg<BN, MT>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-B alignment
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES * Cfg::A_BYTES;
  const uint32_t bars = sB + STAGES * Cfg::B_STAGE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
  const uint32_t tfull_bar = bars + 16 * STAGES, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, 128); }
    fence_barrier_init();
  } else if (warp == 1) {
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int taps = a.k * a.k * a.k;
  const int kchunks = (a.Cin + 63) / 64;
  const int num_kb = taps * kchunks;

  if (warp == 0) {
    {
      // ================= TMA producer (warp-uniform loops, one elected lane issues; see elect_one) =================
      int stage = 0; uint32_t phase = 0;
      int64_t fills = 0;
      for (int64_t tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int nt = (int)(tile % a.n_tiles);
        const int64_t mg = tile / a.n_tiles;
        const int n0 = (a.probe & 2) ? 0 : nt * BN;
        MTile m[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) m[i] = decode_mtile(a, (a.probe & 2) ? (int64_t)i : mg * MT + i);
        for (int tap = 0; tap < taps; ++tap) {
          const int kt = tap / (a.k * a.k), kh = (tap / a.k) % a.k, kw = tap % a.k;
          for (int kc = 0; kc < kchunks; ++kc, ++fills) {
            mbar_wait(empty_bar + 8 * stage, phase ^ 1);
            if (elect_one()) {
              if ((a.probe & 1) && fills >= STAGES) {
                mbar_arrive(full_bar + 8 * stage);
              } else {
                mbar_expect_tx(full_bar + 8 * stage, Cfg::STAGE_BYTES);
#pragma unroll
                for (int i = 0; i < MT; ++i)
                  tma_load_5d(sA + stage * Cfg::A_BYTES + i * A_STAGE_BYTES, &tmA, full_bar + 8 * stage, kc * 64,
                              m[i].w0 * a.sw + kw, m[i].h0 * a.sh + kh, m[i].t * a.st + kt, m[i].b);
                tma_load_3d(sB + stage * Cfg::B_STAGE_BYTES, &tmB, full_bar + 8 * stage, kc * 64, n0, tap);
              }
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ================= MMA issuer (warp-uniform loops, one elected lane issues) =================
      constexpr uint32_t idesc = make_idesc(BN, TcFmt<T>::fmt);
      int stage = 0; uint32_t phase = 0;
      int iter = 0;
      for (int64_t tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t bdesc = make_kmajor_sw128_desc(sB + stage * Cfg::B_STAGE_BYTES);
#pragma unroll
            for (int i = 0; i < MT; ++i) {
              const uint64_t adesc = make_kmajor_sw128_desc(sA + stage * Cfg::A_BYTES + i * A_STAGE_BYTES);
#pragma unroll
              for (int k = 0; k < 4; ++k)  // 4 x (K = 16) per 64-channel chunk: +32 B along K inside the swizzle atom
                umma_f16(d_tmem + i * BN, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            }
            umma_commit(empty_bar + 8 * stage);  // frees the smem slot once these MMAs have read it
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(tfull_bar + 8 * acc);  // accumulator complete -> epilogue
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int iter = 0;
    for (int64_t tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const int nt = (int)(tile % a.n_tiles);
      const int64_t mg = tile / a.n_tiles;
      const int n0 = nt * BN;
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int i = 0; i < MT; ++i) {
        const MTile m = decode_mtile(a, mg * MT + i);
        if (!m.valid || (a.probe & 4)) continue;  // warp-uniform
        epilogue_mtile<T, OT, BN>(a, m, q, lane, tmem_base + (uint32_t)(acc * Cfg::ACC_COLS + i * BN), n0);
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + 8 * acc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------- 2-CTA (cta_group::2) kernel
// A CTA pair (cluster of 2, the two SMs of a TPC) computes a 256-voxel x BN tile with ONE tcgen05.mma.cta_group::2 per
// K=16 step, issued by the leader (cluster rank 0).  CTA r stages ITS m-tile (128 voxels, A operand) and rows
// [r*BN/2, (r+1)*BN/2) of the weight tile (B operand): the weight traffic from L2 and the B reads of the MMA are
// halved per SM compared to the 1-CTA kernel, and the smaller stage (16 KB + BN/2*128 B) allows a deeper TMA ring.
// Protocol: every TMA of both CTAs signals the LEADER's full barrier (address with the peer bit cleared); the leader's
// tcgen05.commit multicasts to the empty / tmem-full barriers of both CTAs; the epilogue warps of both CTAs arrive
// on the leader's tmem-empty barrier.  Each CTA's epilogue reads its own TMEM (its 128 voxels x BN channels).
// KHT ("kh trick", stride-1 3x3x3 convs): the tile is 16 rows x 8 columns of one frame, and ONE A stage holds the
// 18-row halo {64 ch, 8 w, 18 h} of a (kt, kw, channel-chunk) group.  Each h-row is 8 voxels = exactly one 1024-byte
// SWIZZLE_128B atom, so the A operand of tap kh is the same stage at byte offset kh*1024 — still atom aligned.  One
// A load then feeds the three kh taps (3 B stages): A traffic from L2 drops 2.7x, which is what the Cout=128 layers
// (half the MACs per A byte) need to leave the L2-bandwidth bound.  A and B run in separate TMA rings.
template <int BN, bool KHT, bool STAGED = KHT> struct Tc2Cfg {
  static constexpr int NSUB = KHT ? 3 : 1;                            // B stages (taps) consumed per A stage
  static constexpr int A_BYTES = KHT ? 20 * 1024 : A_STAGE_BYTES;     // KHT halo stage: 18 rows x 8 or 10 rows x 16 voxels
  static constexpr int B_STAGE_BYTES = (BN / 2) * 64 * 2;             // this CTA's half of the weight tile
  static constexpr int SA = KHT ? 3 : 0;                              // A ring depth (KHT); non-KHT shares the B ring index
  // Kernels with a 16-bit output (STAGED) run the epilogue through shared memory and TMA stores (see conv_halo.cu): one
  // 128-row x 64-channel SWIZZLE_128B staging block per 64-channel half of the tile.  The per-tap form got it in round 2:
  // with scattered 16-byte global stores a k = 1 GEMM (attention projections: 8 K-steps per tile) spent 50 k of its 64 k
  // cycles in the epilogue (ncu, [17408 x 512] x [512 x 512]: 34 us for 6 us of MMA work).
  static constexpr int NH = (BN + 63) / 64;
  static constexpr int OUT_BYTES = STAGED ? NH * 16384 : 0;
  static constexpr int BIAS_BYTES = STAGED ? 4 * BN * 4 : 0;
  static constexpr int BUDGET = (KHT ? 225 * 1024 : 204 * 1024) - BIAS_BYTES - OUT_BYTES;
  static constexpr int SB_RAW = KHT ? (BUDGET - SA * A_BYTES) / B_STAGE_BYTES : BUDGET / (A_BYTES + B_STAGE_BYTES);
  static constexpr int SB = SB_RAW > 12 ? 12 : SB_RAW;
  static constexpr int NA = KHT ? SA : SB;                            // number of A buffers
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = NA * A_BYTES + SB * B_STAGE_BYTES + OUT_BYTES + 1024 + 1024 + BIAS_BYTES;
  static_assert(SB >= 4, "B ring too shallow");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <typename T, typename OT, int BN, bool KHT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(KHT ? TC2_KHT_THREADS : TC_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const TcArgs a) {
  constexpr bool TMA_EPI = sizeof(OT) == 2;  // staged TMA-store epilogue (16-bit output)
  using Cfg = Tc2Cfg<BN, KHT, TMA_EPI>;
  constexpr int SB = Cfg::SB, NA = Cfg::NA, NSUB = Cfg::NSUB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + NA * Cfg::A_BYTES;
  const uint32_t sOut = sB + SB * Cfg::B_STAGE_BYTES;
  const uint32_t bars = sOut + Cfg::OUT_BYTES;
  const uint32_t bfull_bar = bars, bempty_bar = bars + 8 * SB;
  const uint32_t afull_bar = bars + 16 * SB, aempty_bar = afull_bar + 8 * NA;  // used by KHT only
  const uint32_t tfull_bar = aempty_bar + 8 * NA, tempty_bar = tfull_bar + 16;
  const uint32_t rfull_bar = tempty_bar + 16;  // [4 epilogue warps]: residual tile landed (TMA_EPI only)
  const uint32_t tmem_slot = rfull_bar + 32;
  const uint32_t sbias_all = bars + 1024;      // [4 warps][BN] fp32 (TMA_EPI only)
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < SB; ++s) { mbar_init(bfull_bar + 8 * s, 2); mbar_init(bempty_bar + 8 * s, 1); }
    for (int s = 0; s < NA; ++s) { mbar_init(afull_bar + 8 * s, 2); mbar_init(aempty_bar + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, 256); }
    for (int s = 0; s < 4; ++s) mbar_init(rfull_bar + 8 * s, 1);
    fence_barrier_init();
  }
  cluster_sync_all();  // both CTAs' barriers exist before anything remote touches them
  if (warp == 1) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();  // TMEM of BOTH CTAs is allocated before the leader's first MMA can write it
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int kchunks = (a.Cin + 63) / 64;
  // groups per tile: KHT -> (kt, kw, chunk) with the kh taps inside; otherwise (tap, chunk)
  const int ngroups = (KHT ? a.nkt * a.nkw : a.k * a.k * a.k) * kchunks;
  const int nsub = KHT ? a.nsub : 1;
  const int64_t pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  // ================= TMA producers (both CTAs; warp-uniform loops, one elected lane issues) =================
  // kh-trick form: warp 0 loads the A halo stages and warp 6 the weight stages (one warp doing both, one A and nsub B loads
  // per group, did not stay ahead of the MMA warp on the short phase convs); otherwise warp 0 loads both.
  auto produce = [&](const bool do_a, const bool do_b) {
    {
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      int64_t afills = 0, bfills = 0;
      for (int64_t tile = pair0; tile < a.total_tiles; tile += npairs) {
        const int nt = (int)(tile % a.n_tiles);
        const int64_t mg = tile / a.n_tiles;
        const int n0 = ((a.probe & 2) ? 0 : nt * BN) + (int)rank * (BN / 2);
        const MTile m = decode_mtile(a, (a.probe & 2) ? (int64_t)rank : mg * 2 + rank);
        int cls = 2;
        if (KHT && a.tfold) cls = tile_fold_class(a, mg);
        const int ngroups_t = (KHT && a.tfold) ? (cls + 1) * a.nkw * kchunks : ngroups;
        const int tshift = (KHT && a.tfold) ? 2 - cls : 0;
        if constexpr (KHT) {
          // (kt, kw, channel chunk) groups as nested loops: the flat index with its divisions per group was part of what
          // kept this warp (one A and nsub B loads per group) from staying ahead of the MMA warp on the short phase convs
          const int nkt_eff = a.tfold ? cls + 1 : a.nkt;
          if (a.halo2d) {
            // 2-D halo stage (sub-pixel phases): one A load per (kt, chunk), then the nkw * nsub weight taps it feeds
            for (int kt = 0; kt < a.nkt; ++kt) {
              for (int kc = 0; kc < kchunks; ++kc) {
                if (do_a) {
                  mbar_wait(aempty_bar + 8 * sa, pa ^ 1);
                  const bool skip = (a.probe & 1) && afills >= NA;
                  ++afills;
                  if (elect_one()) {
                    if (leader) { if (skip) mbar_arrive(afull_bar + 8 * sa); else mbar_expect_tx(afull_bar + 8 * sa, 2 * a.a_tx); }
                    if (!skip) tma_load_5d_2sm(sA + sa * Cfg::A_BYTES, &tmA, afull_bar + 8 * sa, kc * 64, m.w0 + a.ow, m.h0 + a.oh, m.t + a.ot + kt, m.b);
                    if (!leader) mbar_arrive_leader(afull_bar + 8 * sa);
                  }
                  __syncwarp();
                  if (++sa == NA) { sa = 0; pa ^= 1; }
                }
                if (do_b) {
                  for (int kw = 0; kw < a.nkw; ++kw) {
                    for (int sub = 0; sub < nsub; ++sub) {
                      mbar_wait(bempty_bar + 8 * sb, pb ^ 1);
                      const bool skipb = (a.probe & 1) && bfills >= SB;
                      ++bfills;
                      if (elect_one()) {
                        if (leader) { if (skipb) mbar_arrive(bfull_bar + 8 * sb); else mbar_expect_tx(bfull_bar + 8 * sb, 2 * Cfg::B_STAGE_BYTES); }
                        if (!skipb) tma_load_3d_2sm(sB + sb * Cfg::B_STAGE_BYTES, &tmB, bfull_bar + 8 * sb, kc * 64, n0, (kt * a.nsub + sub) * a.nkw + kw);
                        if (!leader) mbar_arrive_leader(bfull_bar + 8 * sb);
                      }
                      __syncwarp();
                      if (++sb == SB) { sb = 0; pb ^= 1; }
                    }
                  }
                }
              }
            }
          }
          for (int kt = 0; kt < nkt_eff && !a.halo2d; ++kt) {
            const int wkt = a.tfold ? tfold_wgroup(cls, kt) : kt;
            for (int kw = 0; kw < a.nkw; ++kw) {
              for (int kc = 0; kc < kchunks; ++kc) {
                if (do_a) {
                  mbar_wait(aempty_bar + 8 * sa, pa ^ 1);
                  const bool skip = (a.probe & 1) && afills >= NA;
                  ++afills;
                  if (elect_one()) {
                    if (leader) { if (skip) mbar_arrive(afull_bar + 8 * sa); else mbar_expect_tx(afull_bar + 8 * sa, 2 * a.a_tx); }
                    if (!skip) tma_load_5d_2sm(sA + sa * Cfg::A_BYTES, &tmA, afull_bar + 8 * sa, kc * 64, m.w0 + a.ow + kw, m.h0 + a.oh, m.t + a.ot + tshift + kt, m.b);
                    if (!leader) mbar_arrive_leader(afull_bar + 8 * sa);
                  }
                  __syncwarp();
                  if (++sa == NA) { sa = 0; pa ^= 1; }
                }
#pragma unroll
                for (int sub = 0; sub < NSUB; ++sub) {
                  if (sub >= nsub || !do_b) break;
                  mbar_wait(bempty_bar + 8 * sb, pb ^ 1);
                  const bool skipb = (a.probe & 1) && bfills >= SB;
                  ++bfills;
                  if (elect_one()) {
                    if (leader) { if (skipb) mbar_arrive(bfull_bar + 8 * sb); else mbar_expect_tx(bfull_bar + 8 * sb, 2 * Cfg::B_STAGE_BYTES); }
                    if (!skipb) tma_load_3d_2sm(sB + sb * Cfg::B_STAGE_BYTES, &tmB, bfull_bar + 8 * sb, kc * 64, n0, (wkt * a.nsub + sub) * a.nkw + kw);
                    if (!leader) mbar_arrive_leader(bfull_bar + 8 * sb);
                  }
                  __syncwarp();
                  if (++sb == SB) { sb = 0; pb ^= 1; }
                }
              }
            }
          }
        }
        for (int g = 0; g < ngroups_t && !KHT; ++g) {   // per-tap form: one stage = the A tile and the weight tile of one (tap, chunk)
          const int kc = g % kchunks, tg = g / kchunks;
          const int kt = tg / (a.k * a.k), kh = (tg / a.k) % a.k, kw = tg % a.k;
          mbar_wait(bempty_bar + 8 * sb, pb ^ 1);
          const bool skipb = (a.probe & 1) && bfills >= SB;
          ++bfills;
          if (elect_one()) {
            if (leader) {
              if (skipb) mbar_arrive(bfull_bar + 8 * sb);
              else mbar_expect_tx(bfull_bar + 8 * sb, 2 * (Cfg::B_STAGE_BYTES + Cfg::A_BYTES));
            }
            if (!skipb) {
              tma_load_5d_2sm(sA + sb * Cfg::A_BYTES, &tmA, bfull_bar + 8 * sb, kc * 64, m.w0 * a.sw + kw, m.h0 * a.sh + kh, m.t * a.st + kt, m.b);
              tma_load_3d_2sm(sB + sb * Cfg::B_STAGE_BYTES, &tmB, bfull_bar + 8 * sb, kc * 64, n0, (kt * a.k + kh) * a.k + kw);
            }
            if (!leader) mbar_arrive_leader(bfull_bar + 8 * sb);
          }
          __syncwarp();
          if (++sb == SB) { sb = 0; pb ^= 1; }
        }
        if (KHT) {
          // fused 1x1x1 conv_shortcut: one halo stage of the block input (box origin one row above the tile) and one
          // weight stage per 64-channel chunk
          for (int kc = 0; kc < a.sc_chunks; ++kc) {
            if (do_a) {
              mbar_wait(aempty_bar + 8 * sa, pa ^ 1);
              if (elect_one()) {
                if (leader) mbar_expect_tx(afull_bar + 8 * sa, 2 * a.a_tx);
                tma_load_5d_2sm(sA + sa * Cfg::A_BYTES, &tmX, afull_bar + 8 * sa, kc * 64, m.w0, m.h0 - 1, m.t, m.b);
                if (!leader) mbar_arrive_leader(afull_bar + 8 * sa);
              }
              __syncwarp();
              if (++sa == NA) { sa = 0; pa ^= 1; }
            }
            if (do_b) {
              mbar_wait(bempty_bar + 8 * sb, pb ^ 1);
              if (elect_one()) {
                if (leader) mbar_expect_tx(bfull_bar + 8 * sb, 2 * Cfg::B_STAGE_BYTES);
                tma_load_3d_2sm(sB + sb * Cfg::B_STAGE_BYTES, &tmW, bfull_bar + 8 * sb, kc * 64, n0, 0);
                if (!leader) mbar_arrive_leader(bfull_bar + 8 * sb);
              }
              __syncwarp();
              if (++sb == SB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  };
  if (warp == 0) {
    produce(true, !KHT);
  } else if (KHT && warp == 6) {
    produce(false, true);
  } else if (warp == 1) {
    if (leader) {
      // ================= MMA issuer (leader CTA only; warp-uniform loops, one elected lane issues) =================
      constexpr uint32_t idesc = make_idesc_m256(BN, TcFmt<T>::fmt);
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      int iter = 0;
      for (int64_t tile = pair0; tile < a.total_tiles; tile += npairs, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        int ngroups_t = ngroups;
        if (KHT && a.tfold) ngroups_t = (tile_fold_class(a, (int64_t)((uint32_t)tile / (uint32_t)a.n_tiles)) + 1) * a.nkw * kchunks;
        if constexpr (KHT) {
          // Per A stage: one wait, then per kh tap one wait and four MMAs whose descriptors are the stage bases plus
          // compile-time offsets (tap kh = + kh halo rows of TW voxels; K = 16 slice k = + 32 B).  The tap count per stage
          // (3, or 2 for a sub-pixel phase) is a template argument of the loop: with the run-time `nsub` the loop took ~185
          // instructions per stage and the 8 MMAs of a phase stage (1024 clocks) did not hide it (ncu: tensor pipe 62 %).
          const uint64_t adesc0 = make_kmajor_sw128_desc(sA);
          const uint64_t bdesc0 = make_kmajor_sw128_desc(sB);
          const uint32_t row16 = (uint32_t)(a.TW * 128) >> 4;
          auto run_groups = [&](auto ns_c) {
            constexpr int NS = decltype(ns_c)::value;
#pragma unroll 1
            for (int g = 0; g < ngroups_t; ++g) {
              mbar_wait(afull_bar + 8 * sa, pa);
              const uint64_t ad = adesc0 + (uint64_t)((uint32_t)sa * (uint32_t)(Cfg::A_BYTES >> 4));
              const uint32_t first = g != 0 ? 1u : 0u;
#pragma unroll
              for (int sub = 0; sub < NS; ++sub) {
                mbar_wait(bfull_bar + 8 * sb, pb);
                tc_fence_after();
                if (elect_one()) {
                  const uint64_t as = ad + (uint64_t)((uint32_t)sub * row16);
                  const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)sb * (uint32_t)(Cfg::B_STAGE_BYTES >> 4));
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16_2sm(d_tmem, as + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (sub | k) != 0 ? 1u : first);
                  umma_commit_2sm(bempty_bar + 8 * sb);
                }
                __syncwarp();
                if (++sb == SB) { sb = 0; pb ^= 1; }
              }
              if (elect_one()) umma_commit_2sm(aempty_bar + 8 * sa);
              __syncwarp();
              if (++sa == NA) { sa = 0; pa ^= 1; }
            }
          };
          if (a.halo2d) {
            // sub-pixel phase with a 2-D halo stage: 2 x 2 (kh, kw) taps per stage; tap (kh, kw) = + (kh * PITCH + kw) rows of
            // 128 B with the 8-row-group stride PITCH * 128 B (PITCH = TW + 1 columns; the swizzle follows absolute address bits)
            const uint32_t pitch = (uint32_t)(a.TW + a.nkw - 1);
            const uint64_t hdesc0 = (uint64_t)((sA >> 4) & 0x3FFF) | ((uint64_t)(((pitch * 128u) >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            const int nstage = a.nkt * kchunks;
#pragma unroll 1
            for (int g = 0; g < nstage; ++g) {
              mbar_wait(afull_bar + 8 * sa, pa);
              const uint64_t ad = hdesc0 + (uint64_t)((uint32_t)sa * (uint32_t)(Cfg::A_BYTES >> 4));
              const uint32_t first = g != 0 ? 1u : 0u;
#pragma unroll
              for (int kw = 0; kw < 2; ++kw) {
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                  mbar_wait(bfull_bar + 8 * sb, pb);
                  tc_fence_after();
                  if (elect_one()) {
                    const uint64_t as = ad + (uint64_t)(((uint32_t)sub * pitch + (uint32_t)kw) * 8u);
                    const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)sb * (uint32_t)(Cfg::B_STAGE_BYTES >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      umma_f16_2sm(d_tmem, as + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (kw | sub | k) != 0 ? 1u : first);
                    umma_commit_2sm(bempty_bar + 8 * sb);
                  }
                  __syncwarp();
                  if (++sb == SB) { sb = 0; pb ^= 1; }
                }
              }
              if (elect_one()) umma_commit_2sm(aempty_bar + 8 * sa);
              __syncwarp();
              if (++sa == NA) { sa = 0; pa ^= 1; }
            }
          }
          else if (nsub == 3) run_groups(std::integral_constant<int, 3>{});
          else if (nsub == 2) run_groups(std::integral_constant<int, 2>{});
          else run_groups(std::integral_constant<int, 1>{});
        }
        for (int g = 0; g < ngroups_t && !KHT; ++g) {   // per-tap form: A and B of a (tap, chunk) share the stage index
          mbar_wait(bfull_bar + 8 * sb, pb);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = make_kmajor_sw128_desc(sA + sb * Cfg::A_BYTES);
            const uint64_t bdesc = make_kmajor_sw128_desc(sB + sb * Cfg::B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (g | k) != 0);
            umma_commit_2sm(bempty_bar + 8 * sb);
          }
          __syncwarp();
          if (++sb == SB) { sb = 0; pb ^= 1; }
        }
        if (KHT) {
          for (int kc = 0; kc < a.sc_chunks; ++kc) {  // fused shortcut: the tile rows start one row into the stage
            mbar_wait(afull_bar + 8 * sa, pa);
            mbar_wait(bfull_bar + 8 * sb, pb);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t adesc = make_kmajor_sw128_desc(sA + sa * Cfg::A_BYTES + a.TW * 128);
              const uint64_t bdesc = make_kmajor_sw128_desc(sB + sb * Cfg::B_STAGE_BYTES);
              const int rem = a.sc_cin - kc * 64;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k * 16 < rem) umma_f16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
              umma_commit_2sm(bempty_bar + 8 * sb);
              umma_commit_2sm(aempty_bar + 8 * sa);
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; pb ^= 1; }
            if (++sa == NA) { sa = 0; pa ^= 1; }
          }
        }
        if (elect_one()) umma_commit_2sm(tfull_bar + 8 * acc);
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ================= epilogue warps (both CTAs, own TMEM) =================
    const int q = warp & 3;
    int iter = 0;
    if constexpr (TMA_EPI) {
      // TMEM -> registers -> (+bias, +residual tile fetched by TMA) -> swizzled staging rows -> one TMA store per warp
      // and 64-channel half; GroupNorm partials by the halving tree into per-warp fp64 register accumulators that are
      // flushed when the (batch item, n-tile) changes.  Same scheme as conv_halo.cu.
      constexpr int NH = Cfg::NH;
      const int hq = (32 * q) / a.TW;                                  // first tile row of this warp's 32 voxels
      const int wq = a.TW > 32 ? (32 * q) % a.TW : 0;                  // ... and their first column (tiles wider than 32 voxels)
      const int hh = (32 * q + lane) / a.TW, ww = (32 * q + lane) % a.TW;
      const uint32_t rbar = rfull_bar + 8 * q;
      const uint32_t stage_w = sOut + q * 4096;
      const uint32_t sbias = sbias_all + q * (BN * 4);
      int bias_nt = -1;
      uint32_t rph = 0;
      double gacc[BN / 32];
#pragma unroll
      for (int j = 0; j < BN / 32; ++j) gacc[j] = 0.0;
      int gb = -1, gnt = 0;
      float lacc[64];
#pragma unroll
      for (int e = 0; e < 64; ++e) lacc[e] = 0.f;
      auto gn_flush = [&]() {
        if (a.gn_part == nullptr || gb < 0) return;
        if (epi_lane_acc(BN, a.gn_cpg)) {
          epi_flush_lanes<BN>(lacc, a.gn_cpg, gnt * BN / a.gn_cpg, a.gn_groups,
                              a.gn_part + ((int64_t)gb * a.gn_rows + blockIdx.x * 4 + q) * a.gn_groups * 2, lane);
          return;
        }
        const int V = 2 * (32 / a.gn_cpg);
        const int per = 32 / V;
        if (lane % per == 0) {
          const int idx = lane / per;
          double* row = a.gn_part + ((int64_t)gb * a.gn_rows + blockIdx.x * 4 + q) * a.gn_groups * 2;
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) {
            const int grp = (gnt * BN + 32 * j) / a.gn_cpg + (idx >> 1);
            if (grp < a.gn_groups) row[grp * 2 + (idx & 1)] += gacc[j];  // private slot: plain RMW
          }
        }
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) gacc[j] = 0.0;
      };
      for (int64_t tile = pair0; tile < a.total_tiles; tile += npairs, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        const int nt = (int)(tile % a.n_tiles);
        const int64_t mg = tile / a.n_tiles;
        const int n0 = nt * BN;
        const MTile m = decode_mtile(a, mg * 2 + rank);
        const bool live = m.valid && !(a.probe & 4);
        if (live) {
          if (m.b != gb || nt != gnt) { gn_flush(); gb = m.b; gnt = nt; }
          if (lane == 0) {  // staging rows are free once the previous TMA stores have read them
            bulk_wait_read0();
            if (a.res) {
              mbar_expect_tx(rbar, NH * 4096);
#pragma unroll
              for (int hf = 0; hf < NH; ++hf)
                tma_load_5d(stage_w + hf * 16384, &tmR, rbar, n0 + hf * 64, m.w0 + wq, m.h0 + hq, m.t, m.b);
            }
          }
          __syncwarp();
        }
        mbar_wait(tfull_bar + 8 * acc, acc_phase);
        tc_fence_after();
        if (live) {
          const bool valid = (m.h0 + hh) < a.Ho && (m.w0 + ww) < a.Wo;
          if (a.res) { mbar_wait(rbar, rph); rph ^= 1u; }
          const uint32_t t_cols = tmem_base + (uint32_t)(acc * BN);
          if (nt != bias_nt) { epi_load_bias<BN>(a.bias, n0, a.Cout, sbias, lane); bias_nt = nt; }
          HYVAE_EPI_TILE_SWITCH(T, BN, a.gn_part ? a.gn_cpg : 0, t_cols, q, lane, stage_w, sbias, n0, a.Cout, a.bias != nullptr,
                                a.res != nullptr, a.round_like_ref != 0, valid, gacc, lacc)
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int hf = 0; hf < NH; ++hf)
              if (n0 + hf * 64 < a.Cout)
                tma_store_5d(&tmY, stage_w + hf * 16384, n0 + hf * 64, m.w0 + wq, m.h0 + hq, m.t, m.b);
            bulk_commit();
          }
        }
        tc_fence_before();
        if (leader) mbar_arrive(tempty_bar + 8 * acc);
        else mbar_arrive_leader(tempty_bar + 8 * acc);
      }
      gn_flush();
      if (lane == 0) bulk_wait0();
    } else {
      for (int64_t tile = pair0; tile < a.total_tiles; tile += npairs, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        const int nt = (int)(tile % a.n_tiles);
        const int64_t mg = tile / a.n_tiles;
        mbar_wait(tfull_bar + 8 * acc, acc_phase);
        tc_fence_after();
        const MTile m = decode_mtile(a, mg * 2 + rank);
        if (m.valid && !(a.probe & 4)) epilogue_mtile<T, OT, BN>(a, m, q, lane, tmem_base + (uint32_t)(acc * BN), nt * BN);
        tc_fence_before();
        if (leader) mbar_arrive(tempty_bar + 8 * acc);
        else mbar_arrive_leader(tempty_bar + 8 * acc);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

template <typename T, typename OT, int BN, int MT>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, cudaStream_t stream) {
  using Cfg = TcCfg<BN, MT>;
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    if (cudaFuncSetAttribute(conv_tc_kernel<T, OT, BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
      return fail(HYVAE_ECUDA, "conv_tc: cannot opt in to %d bytes of shared memory", Cfg::SMEM_BYTES);
    attr_once.done();
  }
  int64_t grid = a.total_tiles < num_sms() ? a.total_tiles : num_sms();
  conv_tc_kernel<T, OT, BN, MT><<<(unsigned)grid, TC_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, a);
  return check_launch("conv3d_causal_tc");
}

// Tensor map of the OUTPUT lattice of a conv (or of its residual): logical dims (C, W, H, T, B) with the element strides
// of TcArgs, so a phase of the upsample decomposition (stride-2 scatter into y) is a dense box over a strided view.
static int encode_out_map(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, int64_t off, int C, int W, int H, int T, int B,
                          int64_t sW, int64_t sH, int64_t sT, int64_t sB, int TW) {
  EncodeTiledFn encode = get_encode_fn();
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sT * 2, (cuuint64_t)sB * 2};
  const int bw = TW > 32 ? 32 : TW;
  cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)(32 / bw), 1, 1};  // one warp's 32 voxels of a TH x TW tile
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(tm, dt, 5, (char*)const_cast<void*>(base) + off * 2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(out lattice) failed with %d", (int)r);
}

template <typename T, typename OT, int BN, bool KHT>
static int launch_tc2(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, cudaStream_t stream,
                      const CUtensorMap* tmX = nullptr, const CUtensorMap* tmW = nullptr) {
  using Cfg = Tc2Cfg<BN, KHT, sizeof(OT) == 2>;
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    if (cudaFuncSetAttribute(conv_tc2_kernel<T, OT, BN, KHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
      return fail(HYVAE_ECUDA, "conv_tc2: cannot opt in to %d bytes of shared memory", Cfg::SMEM_BYTES);
    attr_once.done();
  }
  const int64_t max_pairs = conv_sms() / 2;
  const int64_t pairs = a.total_tiles < max_pairs ? a.total_tiles : max_pairs;
  CUtensorMap tmY = tmA, tmR = tmA;  // placeholders unless the staged TMA-store epilogue is compiled in
  if (sizeof(OT) == 2) {
    const CUtensorMapDataType dt = TcFmt<T>::fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    if (int e = encode_out_map(&tmY, dt, a.y, a.yoff, a.Cout, a.Wo, a.Ho, a.To, a.B, a.ysW, a.ysH, a.ysT, a.ysB, a.TW)) return e;
    tmR = tmY;
    if (a.res)
      if (int e = encode_out_map(&tmR, dt, a.res, a.roff, a.Cout, a.Wo, a.Ho, a.To, a.B, a.rsW, a.rsH, a.rsT, a.rsB, a.TW)) return e;
  }
  conv_tc2_kernel<T, OT, BN, KHT><<<(unsigned)(2 * pairs), KHT ? TC2_KHT_THREADS : TC_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmY, tmR, tmX ? *tmX : tmA, tmW ? *tmW : tmB, a);
  return check_launch("conv3d_causal_tc (2-CTA)");
}

}  // namespace hyvae

using namespace hyvae;

// tile shape: TH x TW = 128 output voxels of one frame, chosen to minimise padded area
static void pick_tile(int Ho, int Wo, int sh, int sw, int* TH, int* TW) {
  int best_tw = 16; int64_t best_area = -1;
  for (int tw = 8; tw <= 128; tw <<= 1) {
    int th = 128 / tw;
    if (tw * sw > 256 || th * sh > 256) continue;
    int64_t area = (int64_t)((Ho + th - 1) / th) * th * ((Wo + tw - 1) / tw) * tw;
    if (best_area < 0 || area < best_area || (area == best_area && tw == 16)) { best_area = area; best_tw = tw; }
  }
  *TW = best_tw; *TH = 128 / best_tw;
}

extern "C" int64_t hyvae_conv3d_tc_gn_rows(void) { return (int64_t)gn_partial_rows(); }
extern "C" int64_t hyvae_gn_partials_doubles(int32_t B, int32_t groups) {
  return B <= 0 || groups <= 0 ? 0 : gn_warp_rows_doubles(B, groups) + gn_cta_rows_doubles(B, groups) + 2;
}

static int conv_tc_entry(const hyvae_vol* x, const void* w, const float* bias, const hyvae_vol* residual,
                         const hyvae_vol* y, int32_t k, int32_t st, int32_t sh, int32_t sw,
                         int32_t round_like_ref, int32_t variant, double* gn_partials, int32_t gn_groups,
                         const hyvae_vol* sc_x, const void* sc_w, void* stream);

extern "C" int hyvae_conv3d_causal_tc(const hyvae_vol* x, const void* w, const float* bias, const hyvae_vol* residual,
                                      const hyvae_vol* y, int32_t k, int32_t st, int32_t sh, int32_t sw,
                                      int32_t round_like_ref, int32_t variant, double* gn_partials, int32_t gn_groups,
                                      void* stream) {
  return conv_tc_entry(x, w, bias, residual, y, k, st, sh, sw, round_like_ref, variant, gn_partials, gn_groups, nullptr, nullptr, stream);
}

// conv2 of a ResnetBlockCausal3D whose skip path is a 1x1x1 conv_shortcut (unet_causal_3d_blocks.py:407-415):
// y = conv3x3x3(x) + sc_w * sc_x + bias, where bias is the SUM of both convs' biases.  The shortcut runs as extra K
// chunks of the same accumulator, so its output tensor is never written or re-read.
extern "C" int hyvae_conv3d_causal_tc_shortcut(const hyvae_vol* x, const void* w, const float* bias, const hyvae_vol* sc_x,
                                               const void* sc_w, const hyvae_vol* y, double* gn_partials, int32_t gn_groups,
                                               int32_t w_has_fold, void* stream) {
  if (int e = check_vol(sc_x, "sc_x")) return e;
  HYVAE_CHECK_ARG(sc_w != nullptr, "sc_w is null");
  return conv_tc_entry(x, w, bias, nullptr, y, 3, 1, 1, 1, 0, w_has_fold ? 0x100 : 0, gn_partials, gn_groups, sc_x, sc_w, stream);
}

static int conv_tc_entry(const hyvae_vol* x, const void* w, const float* bias, const hyvae_vol* residual,
                         const hyvae_vol* y, int32_t k, int32_t st, int32_t sh, int32_t sw,
                         int32_t round_like_ref, int32_t variant, double* gn_partials, int32_t gn_groups,
                         const hyvae_vol* sc_x, const void* sc_w, void* stream) {
  if (int e = check_vol(x, "x")) return e;
  if (int e = check_vol(y, "y")) return e;
  HYVAE_CHECK_ARG(w != nullptr, "w is null");
  HYVAE_CHECK_ARG(k == 1 || k == 3, "kernel size %d not supported", k);
  HYVAE_CHECK_ARG(x->dtype == HYVAE_BF16 || x->dtype == HYVAE_F16, "tensor-core conv needs bf16/f16 activations");
  HYVAE_CHECK_ARG((x->dtype == y->dtype || y->dtype == HYVAE_F32) && x->B == y->B, "x / y dtype or batch mismatch");
  HYVAE_CHECK_ARG(x->dtype == y->dtype || residual == nullptr, "residual needs y in x's dtype");
  HYVAE_CHECK_ARG(x->pt == k - 1 && x->ph == k / 2 && x->pw == k / 2, "x must carry the halo (%d,%d,%d), has (%d,%d,%d)", k - 1, k / 2, k / 2, x->pt, x->ph, x->pw);
  HYVAE_CHECK_ARG(x->C % 8 == 0 && y->C % 8 == 0, "Cin and Cout must be multiples of 8 (Cin=%d Cout=%d)", x->C, y->C);
  HYVAE_CHECK_ARG(st >= 1 && sh >= 1 && sw >= 1 && sh <= 2 && sw <= 2, "stride (%d,%d,%d) not supported by the tensor-core path", st, sh, sw);
  HYVAE_CHECK_ARG(y->T == (x->T - 1) / st + 1 && y->H == (x->H - 1) / sh + 1 && y->W == (x->W - 1) / sw + 1, "y dims do not match the conv output");
  HYVAE_CHECK_ARG(((uintptr_t)x->data & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y->data & 15) == 0, "pointers must be 16-byte aligned");
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");

  Vol vx = make_vol(x), vy = make_vol(y);
  TcArgs a;
  a.y = y->data; a.bias = bias;
  a.ysB = vy.sB; a.ysT = vy.sT; a.ysH = vy.sH; a.ysW = vy.sW; a.yoff = vy.at(0, 0, 0, 0);
  if (residual) {
    if (int e = check_vol(residual, "residual")) return e;
    HYVAE_CHECK_ARG(residual->dtype == y->dtype && residual->B == y->B && residual->T == y->T && residual->H == y->H &&
                    residual->W == y->W && residual->C == y->C, "residual shape mismatch");
    Vol vr = make_vol(residual);
    a.res = residual->data; a.rsB = vr.sB; a.rsT = vr.sT; a.rsH = vr.sH; a.rsW = vr.sW; a.roff = vr.at(0, 0, 0, 0);
  } else {
    a.res = nullptr; a.rsB = a.rsT = a.rsH = a.rsW = a.roff = 0;
  }
  a.B = y->B; a.To = y->T; a.Ho = y->H; a.Wo = y->W; a.Cin = x->C; a.Cout = y->C;
  a.k = k; a.st = st; a.sh = sh; a.sw = sw; a.round_like_ref = round_like_ref;
  { const char* pe = getenv("HYVAE_TC_PROBE"); a.probe = pe ? atoi(pe) : 0; }  // measurement only: results are garbage when set
  a.nkt = a.nkw = a.nsub = 3; a.ot = a.oh = a.ow = 0; a.a_tx = 18 * 1024; a.halo2d = 0;
  a.sc_chunks = a.sc_cin = 0;
  // variant bit 8: `w` holds 45 tap slices, the 27 of the conv followed by the 18 folded first-frame taps (tfold_class);
  // used by the halo and kh-trick kernels for stride-1 3x3x3 convs, ignored (first 27 slices) by every other kernel
  const bool w_has_fold = (variant & 0x100) != 0 && k == 3 && st == 1 && sh == 1 && sw == 1;
  // variant bit 9: kw-packed thin input (hyvae_ncthw_to_vol_kw3): x's 16 channels are (kw, c) of a <= 5-channel source and
  // `w` is [9 = (kt, kh)][Cout][16]; only the thin halo kernel takes it
  const bool kwpack = (variant & 0x200) != 0;
  variant &= 0xff;
  const int w_taps = kwpack ? 9 : (w_has_fold ? 45 : k * k * k);
  a.tfold = 0;

  // ---- halo kernel (conv_halo.cu): every stride-1 3x3x3 conv with Cout <= 128 and a 16-bit output (variant 5 forces it)
  const bool halo_ok = k == 3 && st == 1 && sh == 1 && sw == 1 && y->C <= 128 && y->dtype == x->dtype &&
                       (gn_partials == nullptr || (gn_groups > 0 && y->C % gn_groups == 0 && y->C / gn_groups >= 2));
  if (sc_x != nullptr) {
    HYVAE_CHECK_ARG(k == 3 && st == 1 && sh == 1 && sw == 1 && y->C > 64 && y->dtype == x->dtype && variant == 0 &&
                    (gn_partials == nullptr || (gn_groups > 0 && y->C % gn_groups == 0 && y->C / gn_groups >= 2)),
                    "fused shortcut needs a stride-1 3x3x3 conv with Cout > 64 and a 16-bit output");
    HYVAE_CHECK_ARG(sc_x->dtype == x->dtype && sc_x->B == y->B && sc_x->T == y->T && sc_x->H == y->H && sc_x->W == y->W && sc_x->C % 8 == 0,
                    "shortcut input must have y's extent and x's dtype");
    HYVAE_CHECK_ARG(((uintptr_t)sc_x->data & 15) == 0 && ((uintptr_t)sc_w & 15) == 0, "pointers must be 16-byte aligned");
  }
  // ---- stacked-tap kernel (conv_stack.cu): the decoder's conv_out (Cout stored as 8), variant 7 forces it
  if ((variant == 7 || (variant == 0 && halo_ok)) && y->C == 8 && x->C % 64 == 0 && residual == nullptr && gn_partials == nullptr && sc_x == nullptr) {
    HaloArgs h;
    h.bias = bias; h.B = y->B; h.To = y->T; h.Ho = y->H; h.Wo = y->W; h.Cin = x->C; h.Cout = y->C;
    h.tiles_h = (y->H + 15) / 16; h.groups_w = (y->W + 15) / 16;
    h.total = (int64_t)y->B * y->T * h.tiles_h * h.groups_w;
    h.has_res = 0; h.round_like_ref = 0; h.sc_chunks = h.sc_cin = 0; h.gn_part = nullptr; h.gn_groups = h.gn_cpg = h.gn_rows = 0; h.probe = a.probe; h.tfold = 0; h.kwpack = 0;
    const CUtensorMapDataType dt = x->dtype == HYVAE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap tmA, tmB;
    {
      cuuint64_t dims[5] = {(cuuint64_t)x->C, (cuuint64_t)vx.Wp(), (cuuint64_t)vx.Hp(), (cuuint64_t)vx.Tp(), (cuuint64_t)x->B};
      cuuint64_t strides[4] = {(cuuint64_t)vx.sW * 2, (cuuint64_t)vx.sH * 2, (cuuint64_t)vx.sT * 2, (cuuint64_t)vx.sB * 2};
      cuuint32_t box[5] = {64, 18, 18, 1, 1};
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      CUresult r = encode(&tmA, dt, 5, x->data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(A stack) failed with %d", (int)r);
    }
    {  // packed weights [27][8][Cin] == [kt][(kh,kw,c) = 72][Cin]; rows 72..79 of the box are zero fill
      cuuint64_t dims[3] = {(cuuint64_t)x->C, 72, 3};
      cuuint64_t strides[2] = {(cuuint64_t)x->C * 2, (cuuint64_t)x->C * 72 * 2};
      cuuint32_t box[3] = {64, 80, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = encode(&tmB, dt, 3, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(B stack) failed with %d", (int)r);
    }
    char tag[56];
    snprintf(tag, sizeof(tag), "k3 %d->%d %dx%dx%dx%d s111 stacked taps", x->C, y->C, y->B, y->T, y->H, y->W);
    ProfScope prof(PC_CONV_TC, 2.0 * (double)y->B * y->T * y->H * y->W * y->C * x->C * 27, stream, tag);
    return launch_conv_stack(x->dtype, tmA, tmB, h, y->data, vy.sB, vy.sT, vy.sH, vy.sW, vy.at(0, 0, 0, 0), (cudaStream_t)stream);
  }
  if (variant == 5 || variant == 6 || (variant == 0 && halo_ok)) {  // 5 = force the 1-CTA form, 6 = force the CTA-pair form
    HYVAE_CHECK_ARG(k == 3 && st == 1 && sh == 1 && sw == 1 && y->C <= 128 && y->dtype == x->dtype, "halo kernel: needs k=3, stride 1, Cout <= 128, 16-bit output");
    const int bn = y->C > 64 ? 128 : (y->C > 32 ? 64 : 32), mt = 2;
    HaloArgs h;
    h.tiles_h = (y->H + 15) / 16; h.groups_w = (y->W + 8 * mt - 1) / (8 * mt);
    h.total = (int64_t)y->B * y->T * h.tiles_h * h.groups_w;
    // CTA pairs (M = 256 MMAs, half of the weight tile per CTA) for the wide tile once there is work for every pair
    const bool pair = bn == 128 && (variant == 6 || (variant == 0 && h.total >= num_sms()));
    // thin-Cin form (conv_in: 3 channels stored as 16): 32-byte rows, SWIZZLE_32B, resident weights
    const bool thin = x->C == 16 && bn == 128 && sc_x == nullptr && variant != 5;
    const int krow = thin ? 16 : 64;
    const CUtensorMapSwizzle swz = thin ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    int twh, thh, taps_per_b, brows;
    halo_geometry(bn, mt, pair, thin, &twh, &thh, &taps_per_b, &brows);
    h.bias = bias; h.B = y->B; h.To = y->T; h.Ho = y->H; h.Wo = y->W; h.Cin = x->C; h.Cout = y->C;
    h.has_res = residual != nullptr; h.round_like_ref = round_like_ref;
    h.sc_cin = sc_x ? sc_x->C : 0; h.sc_chunks = (h.sc_cin + 63) / 64;
    h.gn_part = gn_partials; h.gn_groups = gn_groups; h.gn_cpg = 0; h.gn_rows = gn_partial_rows(); h.probe = a.probe;
    h.tfold = (w_has_fold && !thin) ? 1 : 0;
    h.kwpack = kwpack ? 1 : 0;
    HYVAE_CHECK_ARG(!kwpack || thin, "kw-packed input needs the thin halo kernel (Cin stored as 16, 64 < Cout <= 128)");
    if (gn_partials) {
      HYVAE_CHECK_ARG(gn_groups > 0 && y->C % gn_groups == 0, "gn_groups=%d does not divide Cout=%d", gn_groups, y->C);
      h.gn_cpg = y->C / gn_groups;
      HYVAE_CHECK_ARG(h.gn_cpg >= 2 && h.gn_cpg <= 32 && (h.gn_cpg & (h.gn_cpg - 1)) == 0, "halo kernel: fused GroupNorm statistics need Cout/groups in {2,...,32} (got %d)", h.gn_cpg);
    }
    const CUtensorMapDataType dt = x->dtype == HYVAE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap tmA, tmB, tmY, tmR;
    {
      cuuint64_t dims[5] = {(cuuint64_t)x->C, (cuuint64_t)vx.Wp(), (cuuint64_t)vx.Hp(), (cuuint64_t)vx.Tp(), (cuuint64_t)x->B};
      cuuint64_t strides[4] = {(cuuint64_t)vx.sW * 2, (cuuint64_t)vx.sH * 2, (cuuint64_t)vx.sT * 2, (cuuint64_t)vx.sB * 2};
      cuuint32_t box[5] = {(cuuint32_t)krow, (cuuint32_t)twh, (cuuint32_t)thh, 1, 1};
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      CUresult r = encode(&tmA, dt, 5, x->data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(A halo) failed with %d", (int)r);
    }
    {
      cuuint64_t dims[3] = {(cuuint64_t)x->C, (cuuint64_t)y->C, (cuuint64_t)w_taps};
      cuuint64_t strides[2] = {(cuuint64_t)x->C * 2, (cuuint64_t)x->C * y->C * 2};
      cuuint32_t box[3] = {(cuuint32_t)krow, (cuuint32_t)brows, (cuuint32_t)taps_per_b};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = encode(&tmB, dt, 3, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swz, thin ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
      cuuint32_t box[5] = {64, (cuuint32_t)twh, (cuuint32_t)thh, 1, 1};
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      CUresult r = encode(&tmX, dt, 5, (char*)sc_x->data + vs.at(0, 0, 0, 0) * 2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(shortcut x) failed with %d", (int)r);
      cuuint64_t wd[3] = {(cuuint64_t)sc_x->C, (cuuint64_t)y->C, 1};
      cuuint64_t ws[2] = {(cuuint64_t)sc_x->C * 2, (cuuint64_t)sc_x->C * y->C * 2};
      cuuint32_t wb[3] = {64, (cuuint32_t)brows, 1};
      cuuint32_t we[3] = {1, 1, 1};
      r = encode(&tmW, dt, 3, const_cast<void*>(sc_w), wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(shortcut w) failed with %d", (int)r);
    }
    char tag[56];
    snprintf(tag, sizeof(tag), "k3 %d->%d %dx%dx%dx%d s111 BN%d halo%s%s%s", x->C, y->C, y->B, y->T, y->H, y->W, bn, pair ? "2" : "", sc_x ? "+sc" : "", thin ? " thin" : "");
    const double vox = (double)y->B * y->T * y->H * y->W;
    const double fold_saved = h.tfold ? 2.0 * (double)y->B * y->H * y->W * y->C * x->C * 27.0 * (y->T >= 2 ? 1.0 : 2.0 / 3.0) : 0.0;
    const double work = 2.0 * vox * y->C * (x->C * 27.0 + (sc_x ? sc_x->C : 0));
    ProfScope prof(PC_CONV_TC, work, stream, tag, work - fold_saved);  // executed: frames 0 / 1 run 1 / 2 of their 3 frame taps
    return launch_halo(x->dtype, bn, mt, pair, thin, tmA, tmB, tmY, tmR, tmX, tmW, h, (cudaStream_t)stream);
  }

  HYVAE_CHECK_ARG(!kwpack, "kw-packed input is only supported by the thin halo kernel (stride-1 3x3x3, Cin stored as 16, 64 < Cout <= 128)");
  pick_tile(y->H, y->W, sh, sw, &a.TH, &a.TW);
  // variant: 0 = auto, 1 = 1-CTA kernel with MT=1, 2 = 1-CTA kernel, 3 = CTA-pair kernel without the kh trick
  const int BN_sel = y->C > 128 ? 256 : (y->C > 64 ? 128 : (y->C > 32 ? 64 : 32));
  bool kht = false;
  // Cout <= 128: an N=128 MMA keeps the shared-memory port saturated with operand reads (A 4 KB + B per 64 cycles),
  // and measured faster on the 1-CTA kernel with two m-tiles per weight tile than on the pair kernel (profiles/).
  const int64_t mt_plain = (int64_t)y->B * y->T * ((y->H + a.TH - 1) / a.TH) * ((y->W + a.TW - 1) / a.TW);
  const bool prefer_1cta = (variant == 0) && BN_sel <= 128 && mt_plain >= 2 * (int64_t)num_sms();
  if (((variant == 0 && !prefer_1cta) || variant == 4) && k == 3 && st == 1 && sh == 1 && sw == 1 && BN_sel >= 64) {
    // 16 x 8 tiles: accept up to 15 % more padded area than the best 128-voxel tile shape
    const int64_t area_best = (int64_t)((y->H + a.TH - 1) / a.TH) * a.TH * ((y->W + a.TW - 1) / a.TW) * a.TW;
    // kh-trick tiles: 16 rows x 8 columns, or 8 x 16 when that wastes less (ragged 18 / 36 / 72-row tiles of the 720p split)
    const int64_t area_168 = (int64_t)((y->H + 15) / 16) * 16 * ((y->W + 7) / 8) * 8;
    const int64_t area_816 = (int64_t)((y->H + 7) / 8) * 8 * ((y->W + 15) / 16) * 16;
    const bool t816 = area_816 < area_168;
    const int kth = t816 ? 8 : 16, ktw = t816 ? 16 : 8;
    const int64_t area_kht = t816 ? area_816 : area_168;
    const int64_t mt_kht = (int64_t)y->B * y->T * ((y->H + kth - 1) / kth) * ((y->W + ktw - 1) / ktw);
    // (the staged epilogue of the kh-trick kernel reduces GroupNorm partials for >= 2 channels per group)
    const bool gn_ok = gn_partials == nullptr || (gn_groups > 0 && y->C % gn_groups == 0 && y->C / gn_groups >= 2);
    if (area_kht * 100 <= area_best * 115 && mt_kht >= 2 && gn_ok) { kht = true; a.TH = kth; a.TW = ktw; a.a_tx = (kth + 2) * ktw * 128; }
  }
  a.tiles_h = (y->H + a.TH - 1) / a.TH; a.tiles_w = (y->W + a.TW - 1) / a.TW;
  const int BN = BN_sel;
  a.n_tiles = (y->C + BN - 1) / BN;
  a.m_tiles = (int64_t)y->B * y->T * a.tiles_h * a.tiles_w;
  a.m_tiles_per_b = (int64_t)y->T * a.tiles_h * a.tiles_w;
  a.gn_part = gn_partials; a.gn_groups = gn_groups; a.gn_cpg = 0; a.gn_rows = gn_partial_rows();
  if (gn_partials) {
    HYVAE_CHECK_ARG(gn_groups > 0 && y->C % gn_groups == 0, "gn_groups=%d does not divide Cout=%d", gn_groups, y->C);
    a.gn_cpg = y->C / gn_groups;
    HYVAE_CHECK_ARG(a.gn_cpg <= 32 && (a.gn_cpg & (a.gn_cpg - 1)) == 0, "fused GroupNorm statistics need Cout/groups in {1,2,4,8,16,32} (got %d)", a.gn_cpg);
  }
  // variant: 0 = auto (CTA-pair kernel whenever there are >= 2 m-tiles), 1 = 1-CTA kernel with MT=1, 2 = 1-CTA kernel
  const bool two_cta = ((variant == 0 && !prefer_1cta) || variant == 3 || variant == 4) && a.m_tiles >= 2 && BN >= 64;
  if (!two_cta && kht) return fail(HYVAE_EINVAL, "internal: kh-trick tile shape chosen without the CTA-pair kernel");
  // 1-CTA kernel: two m-tiles per CTA tile for the narrow-N layers once there is enough work to fill the chip
  const int MT = two_cta ? 2 : ((BN <= 128 && a.m_tiles >= 2 * (int64_t)num_sms() && variant != 1) ? 2 : 1);
  a.total_tiles = ((a.m_tiles + MT - 1) / MT) * a.n_tiles;

  const CUtensorMapDataType dt = x->dtype == HYVAE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmB, tmX, tmW;
  if (sc_x != nullptr) {
    if (!(two_cta && kht)) return fail(HYVAE_EUNSUPPORTED, "fused shortcut needs the halo or the kh-trick pair kernel for this shape");
    a.sc_cin = sc_x->C; a.sc_chunks = (sc_x->C + 63) / 64;
    Vol vs = make_vol(sc_x);
    cuuint64_t dims[5] = {(cuuint64_t)sc_x->C, (cuuint64_t)sc_x->W, (cuuint64_t)sc_x->H, (cuuint64_t)sc_x->T, (cuuint64_t)sc_x->B};
    cuuint64_t strides[4] = {(cuuint64_t)vs.sW * 2, (cuuint64_t)vs.sH * 2, (cuuint64_t)vs.sT * 2, (cuuint64_t)vs.sB * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)a.TW, (cuuint32_t)(a.TH + 2), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmX, dt, 5, (char*)sc_x->data + vs.at(0, 0, 0, 0) * 2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(shortcut x) failed with %d", (int)r);
    cuuint64_t wd[3] = {(cuuint64_t)sc_x->C, (cuuint64_t)y->C, 1};
    cuuint64_t ws[2] = {(cuuint64_t)sc_x->C * 2, (cuuint64_t)sc_x->C * y->C * 2};
    cuuint32_t wb[3] = {64, (cuuint32_t)(BN / 2), 1};
    cuuint32_t we[3] = {1, 1, 1};
    r = encode(&tmW, dt, 3, const_cast<void*>(sc_w), wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(shortcut w) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)x->C, (cuuint64_t)vx.Wp(), (cuuint64_t)vx.Hp(), (cuuint64_t)vx.Tp(), (cuuint64_t)x->B};
    cuuint64_t strides[4] = {(cuuint64_t)vx.sW * 2, (cuuint64_t)vx.sH * 2, (cuuint64_t)vx.sT * 2, (cuuint64_t)vx.sB * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)(a.TW * sw), (cuuint32_t)(kht ? a.TH + 2 : a.TH * sh), 1, 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1, 1};
    CUresult r = encode(&tmA, dt, 5, x->data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)x->C, (cuuint64_t)y->C, (cuuint64_t)w_taps};
    cuuint64_t strides[2] = {(cuuint64_t)x->C * 2, (cuuint64_t)x->C * y->C * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)(two_cta ? BN / 2 : BN), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmB, dt, 3, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  cudaStream_t s = (cudaStream_t)stream;
  char tag[56];
  snprintf(tag, sizeof(tag), "k%d %d->%d %dx%dx%dx%d s%d%d%d BN%d %s%d%s", k, x->C, y->C, y->B, y->T, y->H, y->W, st, sh, sw, BN,
           two_cta ? (kht ? "2ctaK" : "2cta") : "MT", MT, sc_x ? "+sc" : "");
  a.tfold = (w_has_fold && two_cta && kht) ? 1 : 0;
  const double work = 2.0 * (double)y->B * y->T * y->H * y->W * y->C * ((double)x->C * k * k * k + (sc_x ? sc_x->C : 0));
  const double fold_saved = a.tfold ? 2.0 * (double)y->B * y->H * y->W * y->C * x->C * 27.0 * (y->T >= 2 ? 1.0 : 2.0 / 3.0) : 0.0;
  ProfScope prof(PC_CONV_TC, work, stream, tag, work - fold_saved);
  const CUtensorMap* pX = sc_x ? &tmX : nullptr;
  const CUtensorMap* pW = sc_x ? &tmW : nullptr;
#define HYVAE_TC_LAUNCH(T, OT)                                                                              \
  switch (BN) {                                                                                             \
    case 256: return launch_tc<T, OT, 256, 1>(tmA, tmB, a, s);                                              \
    case 128: return MT == 2 ? launch_tc<T, OT, 128, 2>(tmA, tmB, a, s) : launch_tc<T, OT, 128, 1>(tmA, tmB, a, s); \
    case 64: return MT == 2 ? launch_tc<T, OT, 64, 2>(tmA, tmB, a, s) : launch_tc<T, OT, 64, 1>(tmA, tmB, a, s);    \
    default: return MT == 2 ? launch_tc<T, OT, 32, 2>(tmA, tmB, a, s) : launch_tc<T, OT, 32, 1>(tmA, tmB, a, s);    \
  }
  const bool f32out = (y->dtype == HYVAE_F32);
#define HYVAE_TC2_LAUNCH(T, OT)                                                                                        \
  switch (BN) {                                                                                                        \
    case 256: return kht ? launch_tc2<T, OT, 256, true>(tmA, tmB, a, s, pX, pW) : launch_tc2<T, OT, 256, false>(tmA, tmB, a, s); \
    case 128: return kht ? launch_tc2<T, OT, 128, true>(tmA, tmB, a, s, pX, pW) : launch_tc2<T, OT, 128, false>(tmA, tmB, a, s); \
    default: return kht ? launch_tc2<T, OT, 64, true>(tmA, tmB, a, s, pX, pW) : launch_tc2<T, OT, 64, false>(tmA, tmB, a, s);    \
  }
  if (two_cta) {
    if (x->dtype == HYVAE_BF16) {
      if (f32out) { HYVAE_TC2_LAUNCH(__nv_bfloat16, float) } else { HYVAE_TC2_LAUNCH(__nv_bfloat16, __nv_bfloat16) }
    } else {
      if (f32out) { HYVAE_TC2_LAUNCH(__half, float) } else { HYVAE_TC2_LAUNCH(__half, __half) }
    }
  }
#undef HYVAE_TC2_LAUNCH
  if (x->dtype == HYVAE_BF16) {
    if (f32out) { HYVAE_TC_LAUNCH(__nv_bfloat16, float) } else { HYVAE_TC_LAUNCH(__nv_bfloat16, __nv_bfloat16) }
  } else {
    if (f32out) { HYVAE_TC_LAUNCH(__half, float) } else { HYVAE_TC_LAUNCH(__half, __half) }
  }
#undef HYVAE_TC_LAUNCH
}

// ---------------------------------------------------------------------------------- sub-pixel phases of upsample + conv
// UpsampleCausal3D.forward (unet_causal_3d_blocks.py:152-175) = nearest x2 (frame 0 not duplicated in T) followed by a
// 3x3x3 CausalConv3d.  Every high-res tap of output voxel (t', h', w') reads low-res voxel x[f(t')][h'>>1][w'>>1]-ish, so
// per output parity ("phase") the 27 taps collapse onto 2x2x2 (T upsampled) or 3x2x2 (T not upsampled) low-res taps
// whose weights are sums of the original taps:
//   H (same for W): h' = 2i   -> x[i-1]*W0 + x[i]*(W1+W2);      h' = 2i+1 -> x[i]*(W0+W1) + x[i+1]*W2
//   T (up_t == 2):  t' = 2j   -> x[j-1]*W0 + x[j]*(W1+W2);      t' = 2j-1 -> x[j-1]*(W0+W1) + x[j]*W2   (j >= 1)
// with x[-1] = x[0], x[H] = x[H-1] (the replicate halo of the LOW-res volume reproduces the replicate padding of the
// upsampled one).  One call computes one phase: a 2(3)x2x2-tap conv over the low-res volume whose output rows are
// scattered with stride 2 into y.  3.4x (2.25x) fewer MACs than convolving the upsampled tensor, and the 8x larger
// upsampled tensor is never materialised.  Runs on the kh-trick CTA-pair kernel.
extern "C" int hyvae_conv3d_upphase_tc(const hyvae_vol* x, const void* w, const float* bias, const hyvae_vol* y,
                                       int32_t up_t, int32_t pt, int32_t ph, int32_t pw, double* gn_partials,
                                       int32_t gn_groups, void* stream) {
  if (int e = check_vol(x, "x")) return e;
  if (int e = check_vol(y, "y")) return e;
  HYVAE_CHECK_ARG(w != nullptr, "w is null");
  HYVAE_CHECK_ARG((up_t == 1 || up_t == 2) && (pt == 0 || (pt == 1 && up_t == 2)) && (ph == 0 || ph == 1) && (pw == 0 || pw == 1),
                  "bad phase (up_t=%d pt=%d ph=%d pw=%d)", up_t, pt, ph, pw);
  HYVAE_CHECK_ARG((x->dtype == HYVAE_BF16 || x->dtype == HYVAE_F16) && x->dtype == y->dtype && x->B == y->B, "needs 16-bit x and y of one dtype");
  const int nkt = up_t == 2 ? 2 : 3;
  HYVAE_CHECK_ARG(x->pt == nkt - 1 && x->ph == 1 && x->pw == 1, "x must carry the halo (%d,1,1), has (%d,%d,%d)", nkt - 1, x->pt, x->ph, x->pw);
  HYVAE_CHECK_ARG(y->T == (up_t == 2 ? 2 * x->T - 1 : x->T) && y->H == 2 * x->H && y->W == 2 * x->W, "y dims do not match the upsampled conv output");
  HYVAE_CHECK_ARG(x->C % 8 == 0 && y->C % 8 == 0 && y->C >= 64, "Cin %% 8, Cout %% 8 and Cout >= 64 required (Cin=%d Cout=%d)", x->C, y->C);
  HYVAE_CHECK_ARG(((uintptr_t)x->data & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y->data & 15) == 0, "pointers must be 16-byte aligned");
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const int To = (up_t == 2 && pt == 1) ? x->T - 1 : x->T;
  if (To <= 0) return HYVAE_OK;  // a single frame has no odd output frames

  Vol vx = make_vol(x), vy = make_vol(y);
  TcArgs a;
  a.y = y->data; a.bias = bias; a.res = nullptr; a.rsB = a.rsT = a.rsH = a.rsW = a.roff = 0;
  a.ysB = vy.sB; a.ysT = vy.sT * (up_t == 2 ? 2 : 1); a.ysH = vy.sH * 2; a.ysW = vy.sW * 2;
  a.yoff = vy.at(0, (up_t == 2 && pt == 1) ? 1 : 0, ph, pw);
  a.B = y->B; a.To = To; a.Ho = x->H; a.Wo = x->W; a.Cin = x->C; a.Cout = y->C;
  a.k = 3; a.st = a.sh = a.sw = 1; a.round_like_ref = 0;
  { const char* pe = getenv("HYVAE_TC_PROBE"); a.probe = pe ? atoi(pe) : 0; }
  a.nkt = nkt; a.nkw = 2; a.nsub = 2; a.ot = (up_t == 2 && pt == 1) ? 1 : 0; a.oh = ph; a.ow = pw;
  // 2-D halo stage {64 ch, 9, 17}: the two kw taps read the same stage one column apart (HYVAE_PHASE_HALO2D=0: one {64, 8, 17} stage per kw)
  { const char* e = getenv("HYVAE_PHASE_HALO2D"); a.halo2d = (e && e[0] == '0') ? 0 : 1; }
  a.a_tx = (a.halo2d ? 9 : 8) * 17 * 128;
  a.sc_chunks = a.sc_cin = 0;
  a.tfold = 0;
  a.TH = 16; a.TW = 8;
  a.tiles_h = (x->H + 15) / 16; a.tiles_w = (x->W + 7) / 8;
  const int BN = y->C > 128 ? 256 : (y->C > 64 ? 128 : 64);
  a.n_tiles = (y->C + BN - 1) / BN;
  a.m_tiles = (int64_t)y->B * To * a.tiles_h * a.tiles_w;
  a.m_tiles_per_b = (int64_t)To * a.tiles_h * a.tiles_w;
  a.total_tiles = ((a.m_tiles + 1) / 2) * a.n_tiles;
  a.gn_part = gn_partials; a.gn_groups = gn_groups; a.gn_cpg = 0; a.gn_rows = gn_partial_rows();
  if (gn_partials) {
    HYVAE_CHECK_ARG(gn_groups > 0 && y->C % gn_groups == 0, "gn_groups=%d does not divide Cout=%d", gn_groups, y->C);
    a.gn_cpg = y->C / gn_groups;
    HYVAE_CHECK_ARG(a.gn_cpg >= 2 && a.gn_cpg <= 32 && (a.gn_cpg & (a.gn_cpg - 1)) == 0, "fused GroupNorm statistics need Cout/groups in {2,4,8,16,32} (got %d)", a.gn_cpg);
  }
  const CUtensorMapDataType dt = x->dtype == HYVAE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[5] = {(cuuint64_t)x->C, (cuuint64_t)vx.Wp(), (cuuint64_t)vx.Hp(), (cuuint64_t)vx.Tp(), (cuuint64_t)x->B};
    cuuint64_t strides[4] = {(cuuint64_t)vx.sW * 2, (cuuint64_t)vx.sH * 2, (cuuint64_t)vx.sT * 2, (cuuint64_t)vx.sB * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)(a.halo2d ? 9 : 8), 17, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmA, dt, 5, x->data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(A phase) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)x->C, (cuuint64_t)y->C, (cuuint64_t)(nkt * 4)};
    cuuint64_t strides[2] = {(cuuint64_t)x->C * 2, (cuuint64_t)x->C * y->C * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)(BN / 2), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&tmB, dt, 3, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(B phase) failed with %d", (int)r);
  }
  cudaStream_t s = (cudaStream_t)stream;
  char tag[56];
  snprintf(tag, sizeof(tag), "up%d22 p%d%d%d %d->%d lo %dx%dx%dx%d BN%d", up_t, pt, ph, pw, x->C, y->C, y->B, To, x->H, x->W, BN);
  // algorithmic work = what the reference executes for these output voxels: 27 taps at high resolution
  ProfScope prof(PC_CONV_TC, 2.0 * (double)y->B * To * x->H * x->W * y->C * x->C * 27, stream, tag,
                 2.0 * (double)y->B * To * x->H * x->W * y->C * x->C * nkt * 4);
#define HYVAE_UP_LAUNCH(T)                                                           \
  switch (BN) {                                                                      \
    case 256: return launch_tc2<T, T, 256, true>(tmA, tmB, a, s);                    \
    case 128: return launch_tc2<T, T, 128, true>(tmA, tmB, a, s);                    \
    default: return launch_tc2<T, T, 64, true>(tmA, tmB, a, s);                      \
  }
  if (x->dtype == HYVAE_BF16) { HYVAE_UP_LAUNCH(__nv_bfloat16) } else { HYVAE_UP_LAUNCH(__half) }
#undef HYVAE_UP_LAUNCH
}
