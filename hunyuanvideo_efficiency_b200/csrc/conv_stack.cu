// Stride-1 3x3x3 CausalConv3d with a THIN output (Cout <= 8: the decoder's conv_out, 128 -> 3 stored as 8) on tcgen05.
//
// Reference semantics: F.pad(replicate) + nn.Conv3d, unet_causal_3d_blocks.py:73-75 (vae.py:226,292).
//
// With N = 8 (padded to 32) the tap-by-tap implicit GEMM of conv_halo.cu re-reads the 4 KB A operand from shared
// memory for every 16-cycle MMA: it is bound by A reads at 28 % tensor-pipe utilisation (profiles/r01_ncu_*).
// Here the nine (kh, kw) taps are STACKED ALONG N instead: per frame tap kt and 64-channel chunk ONE MMA set
//   Z[p][(kh,kw,c)] += X_halo[p][:] * W[kt][(kh,kw,c)][:]        (M = the 18 x 18 halo voxels p, N = 72 -> 80)
// multiplies every halo voxel with all nine tap matrices at once, and the epilogue forms
//   y[h][w][c] = bias[c] + sum_{kh,kw} Z[(h+kh)*18 + (w+kw)][(kh,kw,c)]
// from a shared-memory copy of Z in a FIXED order (bit-reproducible; no atomics).  MMA count per 256 outputs drops
// from 432 to 72 (N = 80), A reads from shared memory 6x.  The packed weights [27][8][Cin] are already
// [kt][(kh,kw,c)][Cin], so no repacking is needed: rows 72..79 of the weight box are TMA zero fill.
// Roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue.
#include <cuda.h>

#include "common.cuh"
#include "conv_internal.h"
#include "tcgen05.cuh"

namespace hyvae {

namespace stack {
constexpr int THREADS = 192;
constexpr int NHALO = 18 * 18;                 // halo voxels of a 16 x 16 output tile
constexpr int A_TX = NHALO * 128;              // one 64-channel halo stage
constexpr int A_BYTES = (A_TX + 1023) / 1024 * 1024;  // 328 rows; the third m-tile reads on into the next buffer: those
                                                      // rows (>= 324) only produce accumulator rows that are never used
constexpr int NA = 2;
constexpr int BN = 80;                         // stacked N: 9 taps x 8 channels = 72, rounded up to a multiple of 16
constexpr int B_BYTES = BN * 128, B_STRIDE = 11 * 1024, NB = 3;
constexpr int ZPITCH = 76;                     // floats per Z row: 16-byte accesses of consecutive rows hit distinct banks
constexpr int Z_BYTES = (NHALO * ZPITCH * 4 + 1023) / 1024 * 1024;
constexpr int ACC_COLS = 3 * BN;               // 240 fp32 columns per accumulator buffer
constexpr int SMEM_BYTES = NA * A_BYTES + NB * B_STRIDE + Z_BYTES + 2048;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
}  // namespace stack

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // the 4 epilogue warps
__device__ __forceinline__ void sts_f4(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(stack::THREADS, 1)
conv_stack_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloArgs a, T* __restrict__ y,
                  int64_t ysB, int64_t ysT, int64_t ysH, int64_t ysW, int64_t yoff) {
  using namespace stack;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + NA * A_BYTES;
  const uint32_t sZ = sB + NB * B_STRIDE;
  const uint32_t bars = sZ + Z_BYTES;
  const uint32_t afull = bars, aempty = afull + 8 * NA;
  const uint32_t bfull = aempty + 8 * NA, bempty = bfull + 8 * NB;
  const uint32_t tfull = bempty + 8 * NB, tempty = tfull + 16;
  const uint32_t tmem_slot = tempty + 16;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    for (int s = 0; s < NA; ++s) { mbar_init(afull + 8 * s, 1); mbar_init(aempty + 8 * s, 1); }
    for (int s = 0; s < NB; ++s) { mbar_init(bfull + 8 * s, 1); mbar_init(bempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + 8 * s, 1); mbar_init(tempty + 8 * s, 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int kchunks = (a.Cin + 63) / 64;
  const int steps = 3 * kchunks;  // (kt, kc)
  auto decode = [&](int64_t g, int& b, int& t, int& h0, int& w0) {
    const int gw = (int)(g % a.groups_w); g /= a.groups_w;
    const int th = (int)(g % a.tiles_h); g /= a.tiles_h;
    t = (int)(g % a.To); b = (int)(g / a.To); h0 = th * 16; w0 = gw * 16;
  };

  if (warp == 0) {
    // ================= TMA producer =================
    int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
    for (int64_t g = blockIdx.x; g < a.total; g += gridDim.x) {
      int b, t, h0, w0; decode(g, b, t, h0, w0);
      for (int step = 0; step < steps; ++step) {
        const int kt = step / kchunks, kc = step % kchunks;
        mbar_wait(aempty + 8 * sa, pa ^ 1);
        mbar_wait(bempty + 8 * sb, pb ^ 1);
        if (elect_one()) {
          mbar_expect_tx(afull + 8 * sa, A_TX);
          tma_load_5d(sA + sa * A_BYTES, &tmA, afull + 8 * sa, kc * 64, w0, h0, t + kt, b);
          mbar_expect_tx(bfull + 8 * sb, B_BYTES);
          tma_load_3d(sB + sb * B_STRIDE, &tmB, bfull + 8 * sb, kc * 64, 0, kt);
        }
        __syncwarp();
        if (++sa == NA) { sa = 0; pa ^= 1; }
        if (++sb == NB) { sb = 0; pb ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = make_idesc(BN, TcFmt<T>::fmt);
    int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
    int iter = 0;
    for (int64_t g = blockIdx.x; g < a.total; g += gridDim.x, ++iter) {
      const int acc = iter & 1;
      mbar_wait(tempty + 8 * acc, ((iter >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      for (int step = 0; step < steps; ++step) {
        mbar_wait(afull + 8 * sa, pa);
        mbar_wait(bfull + 8 * sb, pb);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bdesc = make_kmajor_sw128_desc(sB + sb * B_STRIDE);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const uint64_t adesc = make_kmajor_sw128_desc(sA + sa * A_BYTES + i * 16384);  // rows are the halo voxels in raster order
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d_tmem + i * BN, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (step | k) != 0);
          }
          umma_commit(aempty + 8 * sa);
          umma_commit(bempty + 8 * sb);
        }
        __syncwarp();
        if (++sa == NA) { sa = 0; pa ^= 1; }
        if (++sb == NB) { sb = 0; pb ^= 1; }
      }
      if (elect_one()) umma_commit(tfull + 8 * acc);
      __syncwarp();
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3, e = q * 32 + lane;  // e: 0..127
    float bias[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bias[c] = (a.bias != nullptr && c < a.Cout) ? a.bias[c] : 0.f;
    int iter = 0;
    for (int64_t g = blockIdx.x; g < a.total; g += gridDim.x, ++iter) {
      const int acc = iter & 1;
      int b, t, h0, w0; decode(g, b, t, h0, w0);
      mbar_wait(tfull + 8 * acc, (iter >> 1) & 1);
      tc_fence_after();
      // Z rows of this thread: p = i * 128 + e  ->  shared memory (72 of the 80 columns)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int p = i * 128 + e;
        const uint32_t tc = tmem_base + (uint32_t)(acc * ACC_COLS + i * BN) + ((uint32_t)(q * 32) << 16);
        uint32_t v0[32], v1[32], v2[8];
        tmem_ld32(tc, v0);
        tmem_ld32(tc + 32, v1);
        tmem_ld8(tc + 64, v2);
        tmem_ld_wait();
        if (p < NHALO) {
          const uint32_t zr = sZ + (uint32_t)p * (ZPITCH * 4);
#pragma unroll
          for (int c = 0; c < 8; ++c) sts_f4(zr + c * 16, __uint_as_float(v0[4 * c]), __uint_as_float(v0[4 * c + 1]), __uint_as_float(v0[4 * c + 2]), __uint_as_float(v0[4 * c + 3]));
#pragma unroll
          for (int c = 0; c < 8; ++c) sts_f4(zr + 128 + c * 16, __uint_as_float(v1[4 * c]), __uint_as_float(v1[4 * c + 1]), __uint_as_float(v1[4 * c + 2]), __uint_as_float(v1[4 * c + 3]));
#pragma unroll
          for (int c = 0; c < 2; ++c) sts_f4(zr + 256 + c * 16, __uint_as_float(v2[4 * c]), __uint_as_float(v2[4 * c + 1]), __uint_as_float(v2[4 * c + 2]), __uint_as_float(v2[4 * c + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(tempty + 8 * acc);  // the accumulator is free: the MMAs of the next tile overlap the gather below
      epi_bar_sync();                 // all Z rows are in shared memory
#pragma unroll
      for (int o2 = 0; o2 < 2; ++o2) {
        const int o = e + 128 * o2, h = o >> 4, w = o & 15;
        float s[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) s[c] = bias[c];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {  // fixed order => bit-reproducible
          const int kh = tap / 3, kw = tap - 3 * kh;
          const uint32_t za = sZ + (uint32_t)((h + kh) * 18 + (w + kw)) * (ZPITCH * 4) + tap * 32;
          const float4 z0 = lds_f4(za), z1 = lds_f4(za + 16);
          s[0] += z0.x; s[1] += z0.y; s[2] += z0.z; s[3] += z0.w; s[4] += z1.x; s[5] += z1.y; s[6] += z1.z; s[7] += z1.w;
        }
        if (h0 + h < a.Ho && w0 + w < a.Wo) {
          Vec8<T> out; out.set(s);
          out.store(y + yoff + (int64_t)b * ysB + (int64_t)t * ysT + (int64_t)(h0 + h) * ysH + (int64_t)(w0 + w) * ysW);
        }
      }
      epi_bar_sync();                 // the gather is done before the next tile overwrites Z
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}

int launch_conv_stack(int dtype, const CUtensorMap& tmA, const CUtensorMap& tmB, const HaloArgs& a, void* y, int64_t ysB, int64_t ysT,
                      int64_t ysH, int64_t ysW, int64_t yoff, cudaStream_t stream) {
  const int64_t grid = a.total < num_sms() ? a.total : num_sms();
#define HYVAE_STACK_LAUNCH(T)                                                                                                        \
  {                                                                                                                                  \
    static DeviceOnce attr_once;                                                                                                    \
    if (attr_once.first()) {                                                                                                                 \
      if (cudaFuncSetAttribute(conv_stack_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, stack::SMEM_BYTES) != cudaSuccess) \
        return fail(HYVAE_ECUDA, "conv_stack: cannot opt in to %d bytes of shared memory", stack::SMEM_BYTES);                       \
      attr_once.done();                                                                                                               \
    }                                                                                                                                \
    conv_stack_kernel<T><<<(unsigned)grid, stack::THREADS, stack::SMEM_BYTES, stream>>>(tmA, tmB, a, (T*)y, ysB, ysT, ysH, ysW, yoff); \
  }
  if (dtype == HYVAE_BF16) HYVAE_STACK_LAUNCH(__nv_bfloat16) else HYVAE_STACK_LAUNCH(__half)
#undef HYVAE_STACK_LAUNCH
  return check_launch("conv3d_causal_tc (stacked taps)");
}

}  // namespace hyvae
