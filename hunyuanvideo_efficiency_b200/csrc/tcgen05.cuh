// tcgen05 / TMEM / TMA / mbarrier PTX wrappers shared by the tensor-core kernels of libhyvae.so (sm_100a).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace hyvae {

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("hyvae conv_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// One lane of a converged warp.  The producer / MMA roles run their loops WARP-UNIFORMLY and predicate only the
// issue instructions with this: loop state, shared-memory addresses and descriptors then live in uniform registers and
// UTMALDG / UTCHMMA take them directly (a role written as `if (lane == 0) { loops }` makes every operand divergent and
// costs an R2UR + ELECT + BRA.U.ANY waterfall of ~25 instructions per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// multicast form: the box lands at the same CTA-relative offset in every CTA of `cta_mask`, and each of them gets the
// complete_tx on its own barrier at `bar`'s offset
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 1-CTA MMAs whose completion must be seen by both CTAs of a cluster pair (a shared-memory slot that a peer's multicast
// TMA refills): arrives on `bar`'s offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 registers per thread -> 32 lanes x 32 consecutive fp32 columns (the inverse of tmem_ld32)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major) | [32,46) SBO >> 4 = 1024 B
//   between 8-row core-matrix groups | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: fp32 accumulate, A/B both K-major, M = 128.
//   [4,6) D fmt (1 = f32) | [7,10) A fmt | [10,13) B fmt (0 = f16, 1 = bf16) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc(int n, int ab_fmt) {
  return (1u << 4) | ((uint32_t)ab_fmt << 7) | ((uint32_t)ab_fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ---------------------------------------------------------------------------------- TMA-store epilogue helpers
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the latest N committed bulk groups have finished READING their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Sum V per-lane values over the 32 lanes with recursive halving: afterwards every lane holds the warp total of
// value index lane * V / 32 (V - 1 + log2(32 / V) shuffles instead of 5 * V).  Fixed tree => bit-reproducible.
template <int V>
__device__ __forceinline__ float halving_reduce(float (&v)[V], int lane) {
  int o = 16;
#pragma unroll
  for (int h = V / 2; h >= 1; h >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
#pragma unroll
  for (; o >= 1; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
  return v[0];
}

// GroupNorm partial of one 32-column chunk: per group of CPG channels (sum, sum of squares) over this warp's rows.
template <int CPG>
__device__ __forceinline__ float gn_chunk_reduce(const float* f, bool valid, int lane) {
  constexpr int V = 2 * (32 / CPG);
  float v[V];
#pragma unroll
  for (int g = 0; g < 32 / CPG; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int c = 0; c < CPG; ++c) { const float u = valid ? f[g * CPG + c] : 0.f; s += u; q = fmaf(u, u, q); }
    v[2 * g] = s; v[2 * g + 1] = q;
  }
  return halving_reduce<V>(v, lane);
}

__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}

// Staged epilogue of ONE 128-voxel x BN-channel accumulator tile for this warp's 32 rows (TMEM lane quarter q):
// TMEM -> registers -> + bias (per-warp copy in shared memory: with 227 KB of shared memory there is no L1 left, a
// global bias load was an L2 round trip per 32 columns) -> + residual (already in the staging rows, put there by TMA)
// -> 16-bit -> swizzled staging rows (the caller issues the TMA store) -> GroupNorm partials by the halving tree into
// the warp's fp64 register accumulators.  CPG (channels per group, 0 = no statistics) is a template parameter and the
// caller switches on it ONCE per tile: with the switch inside the unrolled column loop all five variants were
// interleaved in the instruction stream and the epilogue stalled on instruction fetch (ncu: stall_no_inst).
template <typename T, int BN, int CPG>
__device__ __forceinline__ void epi_tile(uint32_t t_cols, int q, int lane, uint32_t stage_w, uint32_t sbias, int n0, int Cout,
                                         bool has_bias, bool has_res, bool rlr, bool valid, double (&gacc)[BN / 32], float (&lacc)[64]) {
#pragma unroll
  for (int j = 0; j < BN / 32; ++j) {
    uint32_t v[32];
    tmem_ld32(t_cols + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 32), v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
    const int nc = n0 + j * 32;
    if (nc < Cout) {  // warp-uniform
      const uint32_t srow = stage_w + (j >> 1) * 16384 + lane * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (has_bias) {  // columns >= Cout read zeros (the per-warp copy is zero padded)
          const float4 b0 = lds_f4(sbias + (uint32_t)(j * 32 + c * 8) * 4), b1 = lds_f4(sbias + (uint32_t)(j * 32 + c * 8 + 4) * 4);
          f[c * 8 + 0] += b0.x; f[c * 8 + 1] += b0.y; f[c * 8 + 2] += b0.z; f[c * 8 + 3] += b0.w;
          f[c * 8 + 4] += b1.x; f[c * 8 + 5] += b1.y; f[c * 8 + 6] += b1.z; f[c * 8 + 7] += b1.w;
        }
        const uint32_t sa16 = srow + ((uint32_t)((((j & 1) * 4 + c) ^ (lane & 7))) << 4);
        if (has_res) {
          Vec8<T> r; r.v = lds128(sa16);
          float rf[8]; r.get(rf);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[c * 8 + e] = (rlr ? rnd<T>(f[c * 8 + e]) : f[c * 8 + e]) + rf[e];
        }
        Vec8<T> o; o.set(&f[c * 8]);
        sts128(sa16, o.v);
      }
      if constexpr (CPG > 0 && 2 * BN / (CPG > 0 ? CPG : 1) <= 64) {
        // per-LANE fp32 partial sums of this row, kept in registers across tiles; lanes are only combined when the
        // caller flushes (epi_flush_lanes): no shuffle and no fp64 add on the per-tile path
#pragma unroll
        for (int g = 0; g < 32 / CPG; ++g) {
          float sm = 0.f, sq = 0.f;
#pragma unroll
          for (int c = 0; c < CPG; ++c) { const float u = valid ? f[g * CPG + c] : 0.f; sm += u; sq = fmaf(u, u, sq); }
          lacc[(j * (32 / CPG) + g) * 2] += sm;
          lacc[(j * (32 / CPG) + g) * 2 + 1] += sq;
        }
      } else if constexpr (CPG > 0) {
        gacc[j] += (double)gn_chunk_reduce<(CPG > 0 ? CPG : 1)>(f, valid, lane);
      }
    }
  }
}

// Flush of the per-lane GroupNorm accumulators of epi_tile: two halving trees of 32 values leave value v (= (group - first
// group of the tile) * 2 + moment) in lane v % 32 of round v / 32; the lane adds it to the warp's private fp64 row.
template <int BN>
__device__ __forceinline__ void epi_flush_lanes(float (&lacc)[64], int cpg, int first_group, int gn_groups, double* row, int lane) {
  const int nval = 2 * BN / cpg;  // <= 64
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float v[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) { v[e] = lacc[32 * r + e]; lacc[32 * r + e] = 0.f; }
    const float tot = halving_reduce<32>(v, lane);
    const int idx = 32 * r + lane, grp = first_group + (idx >> 1);
    if (idx < nval && grp < gn_groups) row[grp * 2 + (idx & 1)] += (double)tot;  // private slot: plain RMW
  }
}
__device__ __forceinline__ bool epi_lane_acc(int bn, int cpg) { return cpg > 0 && 2 * bn / cpg <= 64; }

// per-warp bias copy: BN floats at `sbias` (shared), zero beyond Cout; each lane fetches 8 consecutive values
template <int BN>
__device__ __forceinline__ void epi_load_bias(const float* bias, int n0, int Cout, uint32_t sbias, int lane) {
  if (lane * 8 < BN) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + lane * 8 + h * 4;
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias != nullptr && n < Cout) b = *reinterpret_cast<const float4*>(bias + n);  // Cout is a multiple of 8
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sbias + (uint32_t)(lane * 8 + h * 4) * 4), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
    }
  }
  __syncwarp();
}

#define HYVAE_EPI_TILE_SWITCH(T, BN, cpg, ...)                         \
  switch (cpg) {                                                       \
    case 0: epi_tile<T, BN, 0>(__VA_ARGS__); break;                    \
    case 2: epi_tile<T, BN, 2>(__VA_ARGS__); break;                    \
    case 4: epi_tile<T, BN, 4>(__VA_ARGS__); break;                    \
    case 8: epi_tile<T, BN, 8>(__VA_ARGS__); break;                    \
    case 16: epi_tile<T, BN, 16>(__VA_ARGS__); break;                  \
    default: epi_tile<T, BN, 32>(__VA_ARGS__); break;                  \
  }

// ---------------------------------------------------------------------------------- first-frame temporal fold
// A stride-1 causal conv pads TWO copies of frame 0 in front (unet_causal_3d_blocks.py:68,74), so output frame 0 sees the
// same frame under all three kt taps and output frame 1 sees frame 0 under kt = 0 and 1.  With the packed weights extended
// by two folded tap groups ([27..35] = W[kt=0]+W[1]+W[2], [36..44] = W[0]+W[1], summed in fp32 and rounded once) a tile of
// output frame 0 needs ONE frame tap and a tile of frame 1 needs TWO: 1/T of the layer's MACs and operand loads are
// never issued (2-11 % at T = 65 ... 9).  Class of a tile = min(t, 2); class 2 is the plain 3-tap schedule, which is also
// valid for frames 0 and 1 (the halo holds the replicated frames) and is what a CTA pair runs when its two tiles differ.
__device__ __forceinline__ int tfold_class(int tfold, int t0, bool v0, int t1, bool v1) {
  if (!tfold) return 2;
  const int c0 = t0 < 2 ? t0 : 2, c1 = t1 < 2 ? t1 : 2;
  if (v0 && v1) return c0 == c1 ? c0 : 2;
  return v0 ? c0 : (v1 ? c1 : 2);
}
// weight tap group (x 9 taps) of folded frame tap ktp, and the padded input frame it reads relative to t: t + (2 - cls) + ktp
__device__ __forceinline__ int tfold_wgroup(int cls, int ktp) { return cls == 2 ? ktp : (cls == 1 ? (ktp == 0 ? 4 : 2) : 3); }

// ---------------------------------------------------------------------------------- CTA-pair (cta_group::2) forms
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even CTA of the pair
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t local_bar) {  // arrive on the barrier at the same offset in cluster rank 0
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(local_bar)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {  // arrives on `bar`'s offset in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m256(int n, int ab_fmt) {
  return (1u << 4) | ((uint32_t)ab_fmt << 7) | ((uint32_t)ab_fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}


}  // namespace hyvae
