// Stride-1 3x3x3 CausalConv3d for Cout <= 128 on tcgen05/TMEM with a FULL 2-D halo stage in shared memory.
//
// Reference semantics: F.pad(replicate) + nn.Conv3d, unet_causal_3d_blocks.py:73-75 (+ residual :415).
//
// Why a second kernel: with N = Cout <= 128 an MMA of 128 voxels x 128 channels x K=16 lasts 64 cycles and reads
// 8 KB of operands from shared memory, i.e. the tensor core already uses the whole 128 B/clk shared-memory port; every
// byte TMA writes into shared memory and every scattered global store of the epilogue competes with it
// (profiles/r01_probe_conv_a.txt: 1.60 PFLOP/s with loads and epilogue disabled, 1.04 with them).  So this kernel
//   * loads the A operand ONCE per (frame tap kt, 64-channel chunk): an 18-row x (8*MT+2)-column halo patch
//     {64 ch, TWH, 18} that feeds all nine (kh, kw) taps of MT adjacent 16x8-voxel m-tiles.  Tap (kh, kw) of m-tile i
//     is the same stage at byte offset (kh*PITCH + kw + 8*i)*128 with the 8-row-group stride (SBO) = PITCH*128:
//     L2 -> SM operand traffic drops from 96 to ~41 B/clk/SM;
//   * runs the epilogue through shared memory: TMEM -> registers -> (+bias, +residual tile fetched by TMA) ->
//     swizzled staging rows -> one TMA store per warp and 64-channel half, instead of 32-way scattered 16-byte stores;
//   * reduces the GroupNorm partial sums with a recursive-halving shuffle tree (16 instead of 80 shuffles per 32
//     columns) and keeps the per-warp fp64 accumulators in registers until the batch item changes;
//   * PAIR = true (Cout = 128): two CTAs of a cluster issue ONE tcgen05.mma.cta_group::2 (M = 256) per K = 16 step and
//     m-tile slot; each CTA stages its own halo patch and only HALF of the weight tile, so the operand reads of the
//     tensor core drop from 128 to 96 B/clk/SM and the weight writes from 32 to 16 B/clk/SM — the sum of all
//     shared-memory traffic (121 B/clk) then fits under the port (it is 168 B/clk for the 1-CTA form).
// Roles (64 + MT * 128 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), then FOUR epilogue warps PER M-TILE SLOT
// (warps 2..5 drain slot 0, warps 6..9 slot 1): thin layers (conv_in: 9 MMAs per 128-channel output tile) and the frame-0 / 1 tiles of
// the temporal fold are bound by the epilogue, which one group of four warps ran at ~4 k clocks per tile.  The producer and MMA loops are
// warp-uniform with one elected lane issuing.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "conv_internal.h"
#include "tcgen05.cuh"

namespace hyvae {

constexpr int HALO_THREADS = 320;   // MT = 2 everywhere: 2 + 4 * MT warps

// THIN (Cin <= 16, e.g. conv_in 3 -> 128 with the input stored as 16 channels): rows are 32 bytes (one K = 16 MMA slice)
// with SWIZZLE_32B instead of 128-byte rows that would be 7/8 zero fill: a halo stage is 10 KB, so the ring is deep enough
// to hide the TMA latency behind the short (9 taps x 64 clk) steps, and ALL 27 weight taps stay resident in shared memory.
template <int BN, int MT, bool PAIR, bool THIN = false> struct HaloCfg {
  static constexpr int ROWB = THIN ? 32 : 128;                       // bytes per voxel row of an operand stage
  static constexpr int TWH = 8 * MT + 2, THH = 18;                  // halo patch: columns x rows
  static constexpr int PITCH = TWH;                                  // smem rows per halo row: dense
  static constexpr int A_TX = TWH * THH * ROWB;                     // bytes TMA delivers per A stage
  static constexpr int A_BYTES = (A_TX + 1023) / 1024 * 1024;
  static constexpr int TB = THIN ? 27 : (BN >= 128 ? 1 : 3);         // taps per B stage (THIN: all of them, loaded once)
  static constexpr int BROWS = PAIR ? BN / 2 : BN;                   // weight rows this CTA stages
  static constexpr int B_TAP_BYTES = BROWS * ROWB;
  static constexpr int B_BYTES = TB * B_TAP_BYTES;
  static constexpr int NH = (BN + 63) / 64;                          // 64-channel halves of the output tile
  static constexpr int OUT_BYTES = MT * NH * 16384;                  // one staging tile per m-tile slot: a tile's TMA store is only
                                                                     // waited for when its slot comes round again, MT tiles later
  static constexpr int BUDGET = 227 * 1024 - 6144;                   // minus alignment slack, barrier block, per-warp bias copies (8 warps)
  static constexpr int NA_THIN_RAW = (BUDGET - OUT_BYTES - (B_BYTES + 1023) / 1024 * 1024) / A_BYTES;
  static constexpr int NA = THIN ? (NA_THIN_RAW > 8 ? 8 : NA_THIN_RAW) : 2;
  static constexpr int NB_RAW = (BUDGET - NA * A_BYTES - OUT_BYTES) / B_BYTES;
  static constexpr int NB = THIN ? 1 : (NB_RAW > 12 ? 12 : NB_RAW);
  static constexpr int B_RING_BYTES = (NB * B_BYTES + 1023) / 1024 * 1024;
  static constexpr int ACC_COLS = MT * BN;
  static constexpr int TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
  static constexpr int SMEM_BYTES = NA * A_BYTES + B_RING_BYTES + OUT_BYTES + 6144;
  static_assert(MT == 2, "one group of four epilogue warps per m-tile slot: HALO_THREADS assumes MT == 2");
  static_assert(THIN || NB >= 3, "B ring too shallow");
  static_assert(!THIN || NA >= 4, "A ring too shallow");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns must be a power of two <= 512");
};

// K-major SWIZZLE_128B descriptor with an arbitrary 8-row-group stride (SBO) and a start address that is only
// 128-byte aligned.  The hardware applies the 128B swizzle to ABSOLUTE shared-memory address bits (measured: results are
// bit-identical to the aligned kernels with the base-offset field left 0, and wrong with it set), so any row of a
// TMA-written SWIZZLE_128B region may be row 0 of an operand, and 8-row groups may start at any row.
__device__ __forceinline__ uint64_t make_halo_desc(uint32_t addr, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// same for 32-byte rows (SWIZZLE_32B, layout type 6): one row = one K = 16 slice, 8-row groups of 256 B when dense
__device__ __forceinline__ uint64_t make_sw32_desc(uint32_t addr, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
}

struct HGroup { int b, t, h0, w0; };
// A group index >= a.total decodes to b >= B: TMA zero-fills its loads and the epilogue skips it (odd group count, PAIR)
__device__ __forceinline__ HGroup decode_group(const HaloArgs& a, int64_t g, int mt_cols) {
  HGroup r;
  const int gw = (int)(g % a.groups_w); g /= a.groups_w;
  const int th = (int)(g % a.tiles_h); g /= a.tiles_h;
  r.t = (int)(g % a.To); r.b = (int)(g / a.To);
  r.h0 = th * 16; r.w0 = gw * mt_cols;
  return r;
}

template <typename T, int BN, int MT, bool PAIR, bool THIN>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const HaloArgs a) {
  using Cfg = HaloCfg<BN, MT, PAIR, THIN>;
  constexpr int NA = Cfg::NA, NB = Cfg::NB, NH = Cfg::NH, PITCH = Cfg::PITCH, TB = Cfg::TB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + NA * Cfg::A_BYTES;
  const uint32_t sOut = sB + Cfg::B_RING_BYTES;
  const uint32_t bars = sOut + Cfg::OUT_BYTES;
  const uint32_t afull = bars, aempty = afull + 8 * NA;
  const uint32_t bfull = aempty + 8 * NA, bempty = bfull + 8 * NB;
  const uint32_t tfull = bempty + 8 * NB, tempty = tfull + 16;
  const uint32_t rfull = tempty + 16;                       // [4 * MT warps]
  const uint32_t tmem_slot = rfull + 8 * 4 * MT;
  const uint32_t sbias_all = bars + 1024;                   // [4 * MT warps][BN] fp32
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr uint32_t NCTA = PAIR ? 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmY);
    if (a.has_res) tma_prefetch_desc(&tmR);
    if (a.sc_chunks) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW); }
    for (int s = 0; s < NA; ++s) { mbar_init(afull + 8 * s, NCTA); mbar_init(aempty + 8 * s, 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(bfull + 8 * s, NCTA); mbar_init(bempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + 8 * s, 1); mbar_init(tempty + 8 * s, 128 * MT * NCTA); }
    for (int s = 0; s < 4 * MT; ++s) mbar_init(rfull + 8 * s, 1);
    fence_barrier_init();
  }
  if constexpr (PAIR) {
    cluster_sync_all();  // both CTAs' barriers exist before anything remote touches them
    if (warp == 1) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();  // TMEM of BOTH CTAs is allocated before the leader's first MMA can write it
  } else {
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int kchunks = (a.Cin + 63) / 64;
  // steps of a group: (kt, kc) of the 3x3x3 conv, then — fused 1x1x1 conv_shortcut of the resnet block
  // (unet_causal_3d_blocks.py:407-415) — one step per 64-channel chunk of the block input, whose halo stage is read
  // at the centre tap only and multiplied with the shortcut weights into the same accumulator
  // (first-frame temporal fold: a unit of output frame 0 / 1 has 1 / 2 frame taps instead of 3, see tfold_class)
  // work units: a group of MT m-tiles per CTA; a CTA pair walks two consecutive groups (2u, 2u+1) per unit
  const int64_t unit0 = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int64_t ustride = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int64_t units = PAIR ? (a.total + 1) / 2 : a.total;
  auto group_of = [&](int64_t u) -> int64_t { return PAIR ? 2 * u + rank : u; };
  // fold class of a unit from the output frame(s) of its group(s): 32-bit divisions only — this also runs on the MMA
  // issuing warp, where decode_group's 64-bit div/mod chain per unit drained the tensor pipe's queue (measured: -3 %)
  const uint32_t gpf = (uint32_t)(a.tiles_h * a.groups_w);  // groups per output frame
  auto frame_of = [&](int64_t g) -> int { return (int)(((uint32_t)g / gpf) % (uint32_t)a.To); };
  auto unit_cls = [&](int64_t u) -> int {
    if (!a.tfold) return 2;
    if constexpr (PAIR) {
      const int64_t g0 = 2 * u, g1 = 2 * u + 1;
      return tfold_class(1, frame_of(g0), g0 < a.total, frame_of(g1), g1 < a.total);
    } else {
      return tfold_class(1, frame_of(u), true, 0, false);
    }
  };

  if (warp == 0) {
    // ================= TMA producer (warp-uniform loops, one elected lane issues) =================
    int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
    auto issue_A = [&](int64_t u, int step, int cls) {
      const HGroup m = decode_group(a, group_of(u), 8 * MT);
      const int steps_main = (cls + 1) * kchunks;
      const bool sc = step >= steps_main;
      const int kt = step / kchunks, kc = sc ? step - steps_main : step % kchunks;
      // shortcut input: logical (unpadded) coordinates, the 1-voxel rim of the box is never read
      const CUtensorMap* map = sc ? &tmX : &tmA;
      const int cw = sc ? m.w0 - 1 : m.w0, ch = sc ? m.h0 - 1 : m.h0, ct = sc ? m.t : m.t + (2 - cls) + kt;
      mbar_wait(aempty + 8 * sa, pa ^ 1);
      if (elect_one()) {
        if constexpr (PAIR) {
          if (leader) mbar_expect_tx(afull + 8 * sa, 2 * Cfg::A_TX);
          tma_load_5d_2sm(sA + sa * Cfg::A_BYTES, map, afull + 8 * sa, kc * 64, cw, ch, ct, m.b);
          if (!leader) mbar_arrive_leader(afull + 8 * sa);
        } else {
          mbar_expect_tx(afull + 8 * sa, Cfg::A_TX);
          tma_load_5d(sA + sa * Cfg::A_BYTES, map, afull + 8 * sa, kc * 64, cw, ch, ct, m.b);
        }
      }
      __syncwarp();
      if (++sa == NA) { sa = 0; pa ^= 1; }
    };
    if constexpr (THIN) {
      // all 27 weight taps once, then a deep ring of 10 KB halo stages (one per frame tap kt)
      if (elect_one()) {
        if constexpr (PAIR) {
          if (leader) mbar_expect_tx(bfull, 2 * Cfg::B_BYTES);
          tma_load_3d_2sm(sB, &tmB, bfull, 0, (int)rank * Cfg::BROWS, 0);
          if (!leader) mbar_arrive_leader(bfull);
        } else {
          mbar_expect_tx(bfull, Cfg::B_BYTES);
          tma_load_3d(sB, &tmB, bfull, 0, 0, 0);
        }
      }
      __syncwarp();
      for (int64_t u = unit0; u < units; u += ustride)
        for (int step = 0; step < 3; ++step) issue_A(u, step, 2);
    }
    bool first = true;
    for (int64_t u = unit0; u < units && !THIN; u += ustride) {
      const int cls = unit_cls(u);
      const int cls_next = (u + ustride < units) ? unit_cls(u + ustride) : 2;
      const int steps_main = (cls + 1) * kchunks, steps_per_group = steps_main + a.sc_chunks;
      for (int step = 0; step < steps_per_group; ++step) {
        if (first) { issue_A(u, step, cls); first = false; }
        const bool sc = step >= steps_main;
        const int kt = step / kchunks, kc = sc ? step - steps_main : step % kchunks;
        const int ntg = sc ? 1 : 9 / TB;  // weight stages of this step (the shortcut has a single tap; TB == 1 there)
        const CUtensorMap* wmap = sc ? &tmW : &tmB;
#pragma unroll 1
        for (int tg = 0; tg < ntg; ++tg) {
          if (tg == ntg / 2) {  // prefetch the next A halo while the MMA works through this one
            if (step + 1 < steps_per_group) issue_A(u, step + 1, cls);
            else if (u + ustride < units) issue_A(u + ustride, 0, cls_next);
          }
          const int wtap = sc ? 0 : tfold_wgroup(cls, kt) * 9 + tg * TB;
          mbar_wait(bempty + 8 * sb, pb ^ 1);
          if (elect_one()) {
            if constexpr (PAIR) {
              if (leader) mbar_expect_tx(bfull + 8 * sb, 2 * Cfg::B_BYTES);
              tma_load_3d_2sm(sB + sb * Cfg::B_BYTES, wmap, bfull + 8 * sb, kc * 64, (int)rank * Cfg::BROWS, wtap);
              if (!leader) mbar_arrive_leader(bfull + 8 * sb);
            } else {
              mbar_expect_tx(bfull + 8 * sb, Cfg::B_BYTES);
              tma_load_3d(sB + sb * Cfg::B_BYTES, wmap, bfull + 8 * sb, kc * 64, 0, wtap);
            }
          }
          __syncwarp();
          if (++sb == NB) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ================= MMA issuer (leader CTA; warp-uniform loops, one elected lane issues) =================
      constexpr uint32_t idesc = PAIR ? make_idesc_m256(BN, TcFmt<T>::fmt) : make_idesc(BN, TcFmt<T>::fmt);
      auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t accum) {
        if constexpr (PAIR) umma_f16_2sm(d, ad, bd, idesc, accum); else umma_f16(d, ad, bd, idesc, accum);
      };
      auto commit = [&](uint32_t bar) { if constexpr (PAIR) umma_commit_2sm(bar); else umma_commit(bar); };
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      int iter = 0;
      if constexpr (THIN) mbar_wait(bfull, 0);  // the resident weights
      for (int64_t u = unit0; u < units && THIN; u += ustride, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        for (int kt = 0; kt < 3; ++kt) {
          mbar_wait(afull + 8 * sa, pa);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_stage = sA + sa * Cfg::A_BYTES;
            if (a.kwpack) {  // the kw taps ride along the channel axis: three (kh) MMAs per frame tap, centre column
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
                const uint64_t bdesc = make_sw32_desc(sB + (uint32_t)((kt * 3 + kh) * Cfg::B_TAP_BYTES), 256);
#pragma unroll
                for (int i = 0; i < MT; ++i)
                  mma(d_tmem + i * BN, make_sw32_desc(a_stage + (uint32_t)((kh * PITCH + 1 + 8 * i) * 32), PITCH * 32), bdesc, (kt | kh) != 0);
              }
            } else {
#pragma unroll
              for (int tap9 = 0; tap9 < 9; ++tap9) {
                const int kh = tap9 / 3, kw = tap9 - 3 * kh;
                const uint64_t bdesc = make_sw32_desc(sB + (uint32_t)((kt * 9 + tap9) * Cfg::B_TAP_BYTES), 256);
#pragma unroll
                for (int i = 0; i < MT; ++i)
                  mma(d_tmem + i * BN, make_sw32_desc(a_stage + (uint32_t)((kh * PITCH + kw + 8 * i) * 32), PITCH * 32), bdesc, (kt | tap9) != 0);
              }
            }
            commit(aempty + 8 * sa);
          }
          __syncwarp();
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
        if (elect_one()) commit(tfull + 8 * acc);
        __syncwarp();
      }
      for (int64_t u = unit0; u < units && !THIN; u += ustride, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        const int steps_main = (unit_cls(u) + 1) * kchunks, steps_per_group = steps_main + a.sc_chunks;
        for (int step = 0; step < steps_per_group; ++step) {
          mbar_wait(afull + 8 * sa, pa);
          const uint32_t a_stage = sA + sa * Cfg::A_BYTES;
          // K = 16 slices that hold real channels in this 64-channel chunk (thin layers: conv_in has Cin = 8)
          const bool sc = step >= steps_main;
          const int rem = sc ? a.sc_cin - (step - steps_main) * 64 : a.Cin - (step % kchunks) * 64;
          const int nk = rem >= 64 ? 4 : (rem + 15) / 16;
          const int ntg = sc ? 1 : 9 / TB;
#pragma unroll 1
          for (int tg = 0; tg < ntg; ++tg) {
            mbar_wait(bfull + 8 * sb, pb);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int tt = 0; tt < TB; ++tt) {
                if (sc && tt > 0) break;
                const int tap9 = sc ? 4 : tg * TB + tt;  // shortcut: centre tap (kh, kw) = (1, 1)
                const int kh = tap9 / 3, kw = tap9 - 3 * kh;
                const uint64_t bdesc = make_kmajor_sw128_desc(sB + sb * Cfg::B_BYTES + tt * Cfg::B_TAP_BYTES);
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                  const uint64_t adesc = make_halo_desc(a_stage + (uint32_t)((kh * PITCH + kw + 8 * i) * 128), PITCH * 128);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (k < nk) mma(d_tmem + i * BN, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), (step | tap9 | k) != 0);
                }
              }
              commit(bempty + 8 * sb);
            }
            __syncwarp();
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
          if (elect_one()) commit(aempty + 8 * sa);
          __syncwarp();
          if (++sa == NA) { sa = 0; pa ^= 1; }
        }
        if (elect_one()) commit(tfull + 8 * acc);
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue warps (every CTA: its own MT m-tiles, its own TMEM) =================
    const int q = warp & 3;  // TMEM lane quarter = rows 32q .. 32q+31 of every m-tile = tile rows 4q .. 4q+3
    const int ew = warp - 2, slot = ew >> 2;  // this warp drains m-tile slot `slot` of every work unit
    const int hh = 4 * q + (lane >> 3), ww = lane & 7;
    double gacc[BN / 32];
#pragma unroll
    for (int j = 0; j < BN / 32; ++j) gacc[j] = 0.0;
    int gb = -1;       // batch item the accumulators belong to
    uint32_t rph = 0;  // phase of this warp's residual barrier
    const uint32_t rbar = rfull + 8 * ew;
    const uint32_t sbias = sbias_all + ew * (BN * 4);
    epi_load_bias<BN>(a.bias, 0, a.Cout, sbias, lane);
    const uint32_t stage_q = sOut + q * 4096;  // this warp's 32 rows of a staging tile (slot i: + i * NH * 16384, half hf: + hf * 16384)
    float lacc[64];
#pragma unroll
    for (int e = 0; e < 64; ++e) lacc[e] = 0.f;
    auto gn_flush = [&]() {
      if (a.gn_part == nullptr || gb < 0) return;
      if (epi_lane_acc(BN, a.gn_cpg)) {
        epi_flush_lanes<BN>(lacc, a.gn_cpg, 0, a.gn_groups, a.gn_part + ((int64_t)gb * a.gn_rows + blockIdx.x * 8 + ew) * a.gn_groups * 2, lane);
        return;
      }
      const int V = 2 * (32 / a.gn_cpg);
      const int per = 32 / V;  // lanes holding the same value
      if (lane % per == 0) {
        const int idx = lane / per;
        double* row = a.gn_part + ((int64_t)gb * a.gn_rows + blockIdx.x * 8 + ew) * a.gn_groups * 2;
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) {
          const int grp = (32 * j) / a.gn_cpg + (idx >> 1);
          if (grp < a.gn_groups) row[grp * 2 + (idx & 1)] += gacc[j];  // private slot: plain RMW
        }
      }
#pragma unroll
      for (int j = 0; j < BN / 32; ++j) gacc[j] = 0.0;
    };
    int iter = 0;
    for (int64_t u = unit0; u < units; u += ustride, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const HGroup m = decode_group(a, group_of(u), 8 * MT);
      const bool live = m.b < a.B && !(a.probe & 4);
      if (live && m.b != gb) { gn_flush(); gb = m.b; }
      // The staging rows of this warp are reused by every m-tile: they are free once the previous TMA stores have read
      // them.  The residual tile of the first m-tile is fetched while the MMAs of this group are still running.
      // A warp owns one m-tile slot: its previous TMA store (one unit ago) is the only bulk group that read the staging rows.
      auto stage_acquire = [&](int i) {
        if (lane == 0) {
          bulk_wait_read<0>();
          if (a.has_res) {
            mbar_expect_tx(rbar, NH * 4096);
#pragma unroll
            for (int hf = 0; hf < NH; ++hf)
              tma_load_5d(stage_q + (i * NH + hf) * 16384, &tmR, rbar, hf * 64, m.w0 + 8 * i, m.h0 + 4 * q, m.t, m.b);
          }
        }
        __syncwarp();
      };
      const int i = slot;
      const bool mine = live && m.w0 + 8 * i < a.Wo;   // warp-uniform
      if (mine) stage_acquire(i);
      mbar_wait(tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const bool row_ok = (m.h0 + hh) < a.Ho;
      if (mine) {
        const uint32_t stage_w = stage_q + i * NH * 16384;
        const bool valid = row_ok && (m.w0 + 8 * i + ww) < a.Wo;
        if (a.has_res) { mbar_wait(rbar, rph); rph ^= 1u; }
        const uint32_t t_cols = tmem_base + (uint32_t)(acc * Cfg::ACC_COLS + i * BN);
        HYVAE_EPI_TILE_SWITCH(T, BN, a.gn_part ? a.gn_cpg : 0, t_cols, q, lane, stage_w, sbias, 0, a.Cout, a.bias != nullptr,
                              a.has_res != 0, a.round_like_ref != 0, valid, gacc, lacc)
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int hf = 0; hf < NH; ++hf)
            if (hf * 64 < a.Cout)
              tma_store_5d(&tmY, stage_w + hf * 16384, hf * 64, m.w0 + 8 * i, m.h0 + 4 * q, m.t, m.b);
          bulk_commit();
        }
      }
      tc_fence_before();
      if constexpr (PAIR) {
        if (leader) mbar_arrive(tempty + 8 * acc); else mbar_arrive_leader(tempty + 8 * acc);
      } else {
        mbar_arrive(tempty + 8 * acc);
      }
    }
    gn_flush();
    if (lane == 0) bulk_wait0();
  }

  tc_fence_before();
  if constexpr (PAIR) {
    cluster_sync_all();  // no CTA of the pair exits (or frees TMEM) while the other may still signal it
    if (warp == 1) { tc_fence_after(); tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base); }
  } else {
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<Cfg::TMEM_COLS>(tmem_base); }
  }
}

template <typename T, int BN, int MT, bool PAIR, bool THIN = false>
static int launch_halo_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
                         const CUtensorMap& tmX, const CUtensorMap& tmW, const HaloArgs& a, cudaStream_t stream) {
  using Cfg = HaloCfg<BN, MT, PAIR, THIN>;
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    if (cudaFuncSetAttribute(conv_halo_kernel<T, BN, MT, PAIR, THIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
      return fail(HYVAE_ECUDA, "conv_halo: cannot opt in to %d bytes of shared memory", Cfg::SMEM_BYTES);
    attr_once.done();
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (PAIR) {
    const int64_t units = (a.total + 1) / 2, max_pairs = conv_sms() / 2;
    cfg.gridDim = dim3((unsigned)(2 * (units < max_pairs ? units : max_pairs)));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3((unsigned)(a.total < num_sms() ? a.total : num_sms()));
  }
  cfg.blockDim = dim3(HALO_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  if (cudaLaunchKernelEx(&cfg, conv_halo_kernel<T, BN, MT, PAIR, THIN>, tmA, tmB, tmY, tmR, tmX, tmW, a) != cudaSuccess)
    return fail(HYVAE_ECUDA, "conv_halo: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return check_launch("conv3d_causal_tc (halo)");
}

// geometry the host needs for the tensor maps: A box {64 (16 thin), twh, thh}; B box {64 (16), brows, taps_per_b}
void halo_geometry(int bn, int mt, bool pair, bool thin, int* twh, int* thh, int* taps_per_b, int* brows) {
  *twh = 8 * mt + 2; *thh = 18; *taps_per_b = thin ? 27 : (bn >= 128 ? 1 : 3); *brows = pair ? bn / 2 : bn;
}

int launch_halo(int dtype, int bn, int mt, bool pair, bool thin, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
                const CUtensorMap& tmR, const CUtensorMap& tmX, const CUtensorMap& tmW, const HaloArgs& a, cudaStream_t stream) {
#define HYVAE_HALO_CASE(T)                                                                                       \
  if (thin && bn == 128 && mt == 2) return pair ? launch_halo_t<T, 128, 2, true, true>(tmA, tmB, tmY, tmR, tmX, tmW, a, stream)    \
                                                : launch_halo_t<T, 128, 2, false, true>(tmA, tmB, tmY, tmR, tmX, tmW, a, stream);  \
  if (thin) return fail(HYVAE_EUNSUPPORTED, "conv_halo: the thin-Cin form is only built for BN=128");            \
  if (bn == 128 && mt == 2) return pair ? launch_halo_t<T, 128, 2, true>(tmA, tmB, tmY, tmR, tmX, tmW, a, stream)           \
                                        : launch_halo_t<T, 128, 2, false>(tmA, tmB, tmY, tmR, tmX, tmW, a, stream);         \
  if (bn == 64 && mt == 2 && !pair) return launch_halo_t<T, 64, 2, false>(tmA, tmB, tmY, tmR, tmX, tmW, a, stream);         \
  if (bn == 32 && mt == 2 && !pair) return launch_halo_t<T, 32, 2, false>(tmA, tmB, tmY, tmR, tmX, tmW, a, stream);
  if (dtype == HYVAE_BF16) { HYVAE_HALO_CASE(__nv_bfloat16) } else { HYVAE_HALO_CASE(__half) }
#undef HYVAE_HALO_CASE
  return fail(HYVAE_EUNSUPPORTED, "conv_halo: no instantiation for BN=%d MT=%d pair=%d", bn, mt, (int)pair);
}

}  // namespace hyvae
