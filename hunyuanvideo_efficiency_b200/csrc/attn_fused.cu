// Fused mid-block attention core on the sm_100a tensor cores: O = softmax_j(scale * Q K^T + frame-causal mask) V + bv
// in ONE kernel (flash style) — the score matrix S and the probabilities P never touch HBM.
//
// Reference semantics: diffusers==0.31.0 Attention / AttnProcessor2_0 as instantiated at unet_causal_3d_blocks.py:580-592
// and called at :661 (one head of width D = C, scale = D^-1/2), with the additive mask of
// prepare_causal_attention_mask (:38-46): row i (frame i / n_hw) sees the keys j < (i / n_hw + 1) * n_hw.
// The mask is derived from indices; KV blocks that lie entirely behind it are never loaded or multiplied.
//
// Work item = (128-query tile, half of the D value channels).  Per item, per 128-key block j:
//   S_j = Q K_j^T        tcgen05.mma, M = 128, N = 128, K = D: Q tile resident in shared memory (D/64 SWIZZLE_128B
//                        chunks of 16 KB), K_j streamed through the TMA ring; fp32 S in TMEM, double buffered
//   P_j = exp2(c (S_j - m))   4 softmax warps, one query row per thread (TMEM lane = row, no shuffles): running row
//                        maximum m with LAZY rescaling (O and the row sum l are only rescaled when the maximum grew
//                        by more than 2^8; P <= 256 stays far inside fp16 range), P rounded to the operand type and
//                        written into a swizzled shared-memory tile = the K-major A operand of the next MMA
//   O  += P_j V_j        tcgen05.mma, M = 128, N = min(D, 256), K = 128: V^T chunks through the same ring; O in TMEM
// Epilogue: O / l + bv -> 16-bit -> global.  TMEM: O = 256 columns, S = 2 x 128 columns (all 512).  With D = 512 the
// O accumulator of a full row does not fit next to S, hence the split of the value channels into two items, each of
// which recomputes S (QK^T is 2/3 of an item's MACs; the alternative — S and P through HBM, 1.8 GB per canonical
// tile — is what this kernel replaces).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = softmax/epilogue.
// Persistent CTAs; items sorted by cost (number of visible key blocks, descending) and dealt in a snake order over the
// CTAs, so the assignment is static (bit-reproducible) and balanced within a few percent.
#include <cuda.h>

#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "conv_internal.h"
#include "tcgen05.cuh"

namespace hyvae {

struct AttnArgs {
  void* o;          // [L][D], 16-bit
  const float* bv;  // [D] fp32 (may be null)
  int L, n_hw, n_qt, n_items;
  float c;          // scale * log2(e)
};

constexpr int ATTN_THREADS = 192;
constexpr int SLOT_BYTES = 16384;  // one TMA box {64 elements, 128 rows}: 128 rows of 128 bytes, SWIZZLE_128B

template <int D> struct AttnCfg {
  static_assert(D == 128 || D == 256 || D == 512, "head width must be 128, 256 or 512");
  static constexpr int QCH = D / 64;                // 64-channel chunks of a Q / K row
  static constexpr int DVH = D > 256 ? 256 : D;     // value channels (O columns) per work item
  static constexpr int NSPLIT = D / DVH;
  static constexpr int VS = DVH / 128;              // ring slots per 64-key chunk of V^T
  static constexpr int Q_BYTES = QCH * SLOT_BYTES;
  static constexpr int P_BYTES = 2 * SLOT_BYTES;    // 128 queries x 128 keys, two 64-key swizzle atoms columns
  static constexpr int NSLOT_RAW = ((227 * 1024 - 2048 - Q_BYTES - P_BYTES) / SLOT_BYTES) & ~1;
  static constexpr int NSLOT = NSLOT_RAW >= 8 ? 8 : 4;
  static constexpr int SMEM_BYTES = Q_BYTES + P_BYTES + NSLOT * SLOT_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
  static_assert(NSLOT_RAW >= 4 && SMEM_BYTES <= 227 * 1024, "shared-memory budget");
};

struct AttnItem { int r0, half, nblk, lim_min; };

// wave w of the snake deal: even waves run left to right over the CTAs, odd waves right to left.
// MC (multicast pair): the unit that is dealt is a PAIR of adjacent query tiles (cluster rank r takes tile 2 qp + r); both
// CTAs walk the key blocks of the later tile in lockstep, because each of them loads half of every K / V^T box and
// multicasts it to both.  Blocks the earlier tile cannot see come out fully masked (P = 0) there.
template <int NSPLIT, bool MC>
__device__ __forceinline__ bool attn_item(const AttnArgs& a, int w, AttnItem& it) {
  const int G = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x, b = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int i = w * G + ((w & 1) ? (G - 1 - b) : b);
  if (i >= a.n_items) return false;
  it.half = i % NSPLIT;
  int rlast;
  if (MC) {
    const int qp = (a.n_qt + 1) / 2 - 1 - i / NSPLIT;
    it.r0 = (2 * qp + (int)cluster_ctarank()) * 128;
    rlast = min(2 * qp * 128 + 255, a.L - 1);
  } else {
    it.r0 = (a.n_qt - 1 - i / NSPLIT) * 128;  // late query tiles see the most keys: they go first
    rlast = min(it.r0 + 127, a.L - 1);
  }
  const int lim_max = (rlast / a.n_hw + 1) * a.n_hw;
  it.lim_min = (min(it.r0, a.L - 1) / a.n_hw + 1) * a.n_hw;
  it.nblk = (lim_max + 127) / 128;
  return true;
}

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <typename T, int D, bool MC>
__global__ void __launch_bounds__(ATTN_THREADS, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnArgs a) {
  using Cfg = AttnCfg<D>;
  constexpr int NSLOT = Cfg::NSLOT, QCH = Cfg::QCH, DVH = Cfg::DVH, VS = Cfg::VS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sP = sQ + Cfg::Q_BYTES;
  const uint32_t ring = sP + Cfg::P_BYTES;
  const uint32_t bars = ring + NSLOT * SLOT_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * NSLOT;
  const uint32_t q_full = bars + 16 * NSLOT, q_free = q_full + 8;
  const uint32_t s_full = q_free + 8, s_free = s_full + 16;
  const uint32_t p_full = s_free + 16, p_free = p_full + 8;
  const uint32_t o_full = p_free + 8, o_free = o_full + 8;
  const uint32_t tmem_slot = o_free + 8;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    // MC: a slot is refilled by BOTH CTAs' multicasts, so it is free only when both CTAs' MMAs have read it
    for (int s = 0; s < NSLOT; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, MC ? 2 : 1); }
    mbar_init(q_full, 1); mbar_init(q_free, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(s_full + 8 * s, 1); mbar_init(s_free + 8 * s, 128); }
    mbar_init(p_full, 128); mbar_init(p_free, 1);
    mbar_init(o_full, 1); mbar_init(o_free, 128);
    fence_barrier_init();
  } else if (warp == 1) {
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // the peer's barriers exist before anything remote (multicast TMA, commit) touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t t_o = tmem_base, t_s = tmem_base + 256;  // O: columns [0, DVH); S buffers: [256, 384), [384, 512)
  const int G = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t crank = MC ? cluster_ctarank() : 0u;

  if (warp == 0) {
    // ================= TMA producer: the ring order is exactly the MMA issue order =================
    uint32_t cnt = 0;
    int it = 0;
    for (int w = 0; w * G < a.n_items; ++w) {
      AttnItem im;
      if (!attn_item<Cfg::NSPLIT, MC>(a, w, im)) continue;
      mbar_wait(q_free, (uint32_t)((it & 1) ^ 1));  // every S MMA of the previous item has read the old Q tile
      if (elect_one()) {
        mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
        for (int kc = 0; kc < QCH; ++kc) tma_load_2d(sQ + kc * SLOT_BYTES, &tmQ, q_full, kc * 64, im.r0);
      }
      __syncwarp();
      auto load = [&](const CUtensorMap* tm, int c0, int c1) {
        const uint32_t slot = cnt % NSLOT, ph = (cnt / NSLOT) & 1;
        mbar_wait(empty_bar + 8 * slot, ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(full_bar + 8 * slot, SLOT_BYTES);  // MC: 8 KB from this CTA's load + 8 KB from the peer's
          if constexpr (MC)  // `tm` has a 64-row box: this CTA fetches rows [64 r, 64 r + 64) of the box for both CTAs
            tma_load_2d_mc(ring + slot * SLOT_BYTES + crank * (SLOT_BYTES / 2), tm, full_bar + 8 * slot, c0, c1 + (int)crank * 64, (uint16_t)3);
          else
            tma_load_2d(ring + slot * SLOT_BYTES, tm, full_bar + 8 * slot, c0, c1);
        }
        __syncwarp();
        ++cnt;
      };
      auto load_k = [&](int j) {
        for (int kc = 0; kc < QCH; ++kc) load(&tmK, kc * 64, j * 128);
      };
      auto load_v = [&](int j) {
        for (int c = 0; c < 2; ++c)
          for (int v = 0; v < VS; ++v) load(&tmV, j * 128 + c * 64, im.half * DVH + v * 128);
      };
      load_k(0);
      if (im.nblk > 1) load_k(1);
      for (int j = 0; j < im.nblk; ++j) {
        load_v(j);
        if (j + 2 < im.nblk) load_k(j + 2);
      }
      ++it;
    }
  } else if (warp == 1) {
    // ================= MMA issuer: S(0) S(1) PV(0) S(2) PV(1) ... — S(j+1) runs while the softmax works on S(j) =================
    constexpr uint32_t idesc_s = make_idesc(128, TcFmt<T>::fmt), idesc_o = make_idesc(DVH, TcFmt<T>::fmt);
    uint32_t cnt = 0, gb_s = 0, gb_p = 0;
    int it = 0;
    for (int w = 0; w * G < a.n_items; ++w) {
      AttnItem im;
      if (!attn_item<Cfg::NSPLIT, MC>(a, w, im)) continue;
      mbar_wait(q_full, (uint32_t)(it & 1));
      tc_fence_after();
      auto issue_s = [&](int j) {
        const uint32_t buf = gb_s & 1, ph = (gb_s >> 1) & 1;
        mbar_wait(s_free + 8 * buf, ph ^ 1);  // the softmax warps have read the previous contents of this S buffer
        tc_fence_after();
        const uint32_t d = t_s + buf * 128;
        for (int kc = 0; kc < QCH; ++kc) {
          const uint32_t slot = cnt % NSLOT, sph = (cnt / NSLOT) & 1;
          mbar_wait(full_bar + 8 * slot, sph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = make_kmajor_sw128_desc(sQ + kc * SLOT_BYTES);
            const uint64_t bdesc = make_kmajor_sw128_desc(ring + slot * SLOT_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_s, (kc | k) != 0);
            if constexpr (MC) umma_commit_mc(empty_bar + 8 * slot, (uint16_t)3); else umma_commit(empty_bar + 8 * slot);
          }
          __syncwarp();
          ++cnt;
        }
        if (elect_one()) {
          umma_commit(s_full + 8 * buf);
          if (j == im.nblk - 1) umma_commit(q_free);
        }
        __syncwarp();
        ++gb_s;
      };
      auto issue_pv = [&](int j) {
        mbar_wait(p_full, gb_p & 1);                                   // P_j is in shared memory (and O rescaled if needed)
        if (j == 0) mbar_wait(o_free, (uint32_t)((it & 1) ^ 1));       // the previous item's epilogue has read O
        tc_fence_after();
        for (int c = 0; c < 2; ++c) {
          const uint32_t slot = cnt % NSLOT, sph = (cnt / NSLOT) & 1;  // slot is even and VS <= 2: no wrap inside a chunk
#pragma unroll
          for (int v = 0; v < VS; ++v) mbar_wait(full_bar + 8 * (slot + v), sph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = make_kmajor_sw128_desc(sP + c * SLOT_BYTES);
            const uint64_t bdesc = make_kmajor_sw128_desc(ring + slot * SLOT_BYTES);  // DVH rows: VS consecutive slots
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(t_o, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_o, (j | c | k) != 0);
#pragma unroll
            for (int v = 0; v < VS; ++v) {
              if constexpr (MC) umma_commit_mc(empty_bar + 8 * (slot + v), (uint16_t)3); else umma_commit(empty_bar + 8 * (slot + v));
            }
          }
          __syncwarp();
          cnt += VS;
        }
        if (elect_one()) {
          umma_commit(p_free);
          if (j == im.nblk - 1) umma_commit(o_full);
        }
        __syncwarp();
        ++gb_p;
      };
      issue_s(0);
      if (im.nblk > 1) issue_s(1);
      for (int j = 0; j < im.nblk; ++j) {
        issue_pv(j);
        if (j + 2 < im.nblk) issue_s(j + 2);
      }
      ++it;
    }
  } else {
    // ================= softmax + epilogue warps: thread = query row = TMEM lane =================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const float c = a.c;
    uint32_t gb = 0;
    int it = 0;
    T* od = reinterpret_cast<T*>(a.o);
    for (int w = 0; w * G < a.n_items; ++w) {
      AttnItem im;
      if (!attn_item<Cfg::NSPLIT, MC>(a, w, im)) continue;
      const int grow = im.r0 + row;
      const int lim_r = grow < a.L ? (grow / a.n_hw + 1) * a.n_hw : a.L;
      float m_run = 0.f, l = 0.f;
      for (int j = 0; j < im.nblk; ++j, ++gb) {
        const uint32_t buf = gb & 1, ph = (gb >> 1) & 1;
        mbar_wait(s_full + 8 * buf, ph);
        tc_fence_after();
        const uint32_t ts = t_s + t_lane + buf * 128;
        const bool masked = j * 128 + 128 > im.lim_min;  // warp-uniform: some row of the tile has hidden keys in this block
        const int nvalid = lim_r - j * 128;              // this row sees the block's columns [0, nvalid)
        // ---- pass 1: row maximum of the visible scores
        float mx = -INFINITY;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t v[32];
          tmem_ld32(ts + ch * 32, v);
          tmem_ld_wait();
          if (masked) {
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, (ch * 32 + e < nvalid) ? __uint_as_float(v[e]) : -INFINITY);
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(v[e]));
          }
        }
        // lazy rescaling: keep the stale maximum unless the new one is more than 2^8 above it
        float fac = 1.f;
        bool resc = false;
        if (j == 0) {
          m_run = mx;  // column 0 is visible to every row, so mx is finite
        } else if ((mx - m_run) * c > 8.f) {
          fac = ex2_approx((m_run - mx) * c);
          m_run = mx;
          l *= fac;
          resc = true;
        }
        const float mc = m_run * c;
        // ---- pass 2: P = exp2(c S - c m), row sum, rounded to the operand type
        uint32_t preg[64];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t v[32];
          tmem_ld32(ts + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float p0 = ex2_approx(fmaf(__uint_as_float(v[e]), c, -mc));
            float p1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), c, -mc));
            if (masked) {
              if (ch * 32 + e >= nvalid) p0 = 0.f;
              if (ch * 32 + e + 1 >= nvalid) p1 = 0.f;
            }
            l += p0 + p1;
            preg[ch * 16 + e / 2] = pack2<T>(p0, p1);
          }
        }
        tc_fence_before();
        mbar_arrive(s_free + 8 * buf);  // S buffer may be overwritten by S(j + 2)
        // ---- P tile (and O, if it must be rescaled) may only be touched once PV(j - 1) has completed
        mbar_wait(p_free, (gb & 1) ^ 1);
        if (__any_sync(0xffffffffu, resc)) {
          tc_fence_after();
#pragma unroll 1
          for (int ch = 0; ch < DVH / 32; ++ch) {
            uint32_t v[32];
            tmem_ld32(t_o + t_lane + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * fac);
            tmem_st32(t_o + t_lane + ch * 32, v);
          }
          tmem_st_wait();
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const uint4 val = make_uint4(preg[4 * u], preg[4 * u + 1], preg[4 * u + 2], preg[4 * u + 3]);
          sts128(sP + (uint32_t)(u >> 3) * SLOT_BYTES + (uint32_t)row * 128 + ((uint32_t)((u & 7) ^ (row & 7)) << 4), val);
        }
        fence_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        mbar_arrive(p_full);
      }
      // ---- epilogue: O / l + bv
      mbar_wait(o_full, (uint32_t)(it & 1));
      tc_fence_after();
      const float inv = 1.f / l;
      const int col0 = im.half * DVH;
#pragma unroll 1
      for (int ch = 0; ch < DVH / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(t_o + t_lane + ch * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]) * inv;
        if (a.bv != nullptr) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bv + col0 + ch * 32 + e));
            f[e] += b.x; f[e + 1] += b.y; f[e + 2] += b.z; f[e + 3] += b.w;
          }
        }
        if (grow < a.L) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            Vec8<T> o; o.set(&f[g * 8]);
            o.store(od + (int64_t)grow * D + col0 + ch * 32 + g * 8);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(o_free);
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn attn_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major [rows][cols] 16-bit matrix, box {64 columns, box_rows rows}; rows / columns beyond the matrix are zero-filled
static int encode_rows_map(CUtensorMap* tm, CUtensorMapDataType dt, const void* base, int64_t rows, int64_t cols, const char* what,
                           int box_rows = 128) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = attn_encode_fn()(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
}

template <typename T, int D, bool MC>
static int launch_attn(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, AttnArgs a, cudaStream_t stream) {
  using Cfg = AttnCfg<D>;
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    if (cudaFuncSetAttribute(attn_fused_kernel<T, D, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
      return fail(HYVAE_ECUDA, "attn_fused: cannot opt in to %d bytes of shared memory", Cfg::SMEM_BYTES);
    attr_once.done();
  }
  if (!MC) {
    a.n_items = a.n_qt * Cfg::NSPLIT;
    const int grid = a.n_items < num_sms() ? a.n_items : num_sms();
    attn_fused_kernel<T, D, false><<<(unsigned)grid, ATTN_THREADS, Cfg::SMEM_BYTES, stream>>>(tmQ, tmK, tmV, a);
    return check_launch("attn_block_causal");
  }
  // multicast pair form: clusters of two CTAs, one PAIR of query tiles per item
  a.n_items = ((a.n_qt + 1) / 2) * Cfg::NSPLIT;
  const int pairs = a.n_items < num_sms() / 2 ? a.n_items : num_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs)); cfg.blockDim = dim3(ATTN_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, attn_fused_kernel<T, D, true>, tmQ, tmK, tmV, a) != cudaSuccess) { /* reported by check_launch */ }
  return check_launch("attn_block_causal (multicast pair)");
}

}  // namespace hyvae

using namespace hyvae;

extern "C" int hyvae_attn_block_causal(const void* q, const void* k, const void* vt, const float* bv, void* o, int32_t dtype,
                                       int64_t L, int32_t n_hw, int32_t D, float scale, void* stream) {
  HYVAE_CHECK_ARG(q && k && vt && o, "attn_block_causal: null pointer");
  HYVAE_CHECK_ARG(dtype == HYVAE_BF16 || dtype == HYVAE_F16, "attn_block_causal needs bf16/f16 operands");
  HYVAE_CHECK_ARG(L > 0 && n_hw > 0 && L % n_hw == 0 && L < (1 << 30), "attn_block_causal: L=%lld must be a positive multiple of n_hw=%d",
                  (long long)L, n_hw);
  HYVAE_CHECK_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)vt | (uintptr_t)o | (uintptr_t)bv) & 15) == 0, "pointers must be 16-byte aligned");
  if (!(D == 128 || D == 256 || D == 512) || L % 8 != 0 || !hyvae_device_supports_tc())
    return fail(HYVAE_EUNSUPPORTED, "attn_block_causal: D=%d / L=%lld not covered by the fused kernel (D in {128,256,512}, L %% 8 == 0)", D, (long long)L);
  if (!attn_encode_fn()) return fail(HYVAE_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const CUtensorMapDataType dt = dtype == HYVAE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  // HYVAE_ATTN_MULTICAST=1 (experiment, default off): CTA pairs that multicast the K / V^T boxes to each other, halving the
  // L2 -> SM TMA traffic (3.9 GB per launch at L = 17408, profiles/r01_ncu_attn_fused_512.txt).  Measured: correct, but only
  // 0.428 -> 0.415 ms — the kernel is not bound by L2 bandwidth but by the shared-memory port (an N = 128 MMA reads its
  // 8 KB of operands in 64 cycles = the whole port, and the TMA fill of the ring competes with it), which multicast does
  // not relieve; cta_group::2 MMAs (half of the B operand per CTA) would.
  static const bool mc_env = [] { const char* e = getenv("HYVAE_ATTN_MULTICAST"); return e != nullptr && e[0] == '1'; }();
  const bool mc = mc_env && D == 512;  // the pair form is only instantiated for the production head width
  CUtensorMap tmQ, tmK, tmV;
  if (int e = encode_rows_map(&tmQ, dt, q, L, D, "Q")) return e;
  if (int e = encode_rows_map(&tmK, dt, k, L, D, "K", mc ? 64 : 128)) return e;
  if (int e = encode_rows_map(&tmV, dt, vt, D, L, "V^T", mc ? 64 : 128)) return e;
  AttnArgs a;
  a.o = o; a.bv = bv; a.L = (int)L; a.n_hw = n_hw; a.n_qt = (int)((L + 127) / 128); a.n_items = 0;
  a.c = scale * 1.4426950408889634f;
  char tag[56];
  snprintf(tag, sizeof(tag), "attn fused L=%lld n_hw=%d D=%d", (long long)L, n_hw, D);
  ProfScope prof(PC_ATTN, 4.0 * (double)L * (double)L * D, stream, tag);  // dense SDPA flops, not halved for causality
  cudaStream_t st = (cudaStream_t)stream;
  if (mc) {
    if (dtype == HYVAE_F16) return launch_attn<__half, 512, true>(tmQ, tmK, tmV, a, st);
    return launch_attn<__nv_bfloat16, 512, true>(tmQ, tmK, tmV, a, st);
  }
  if (dtype == HYVAE_F16) {
    if (D == 512) return launch_attn<__half, 512, false>(tmQ, tmK, tmV, a, st);
    if (D == 256) return launch_attn<__half, 256, false>(tmQ, tmK, tmV, a, st);
    return launch_attn<__half, 128, false>(tmQ, tmK, tmV, a, st);
  }
  if (D == 512) return launch_attn<__nv_bfloat16, 512, false>(tmQ, tmK, tmV, a, st);
  if (D == 256) return launch_attn<__nv_bfloat16, 256, false>(tmQ, tmK, tmV, a, st);
  return launch_attn<__nv_bfloat16, 128, false>(tmQ, tmK, tmV, a, st);
}
