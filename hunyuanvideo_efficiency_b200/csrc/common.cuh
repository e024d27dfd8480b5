// Shared helpers for libhyvae.so (sm_100a).  See include/hyvae.h for the ABI.
#pragma once

#include <cstdlib>

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "../../include/hyvae.h"

namespace hyvae {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define HYVAE_CHECK_ARG(cond, ...)                                   \
  do {                                                               \
    if (!(cond)) return ::hyvae::fail(HYVAE_EINVAL, __VA_ARGS__);    \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(HYVAE_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return HYVAE_OK;
}

// ---- optional per-kernel-class timing (CUDA events on the launching stream; bench.py's roofline leg) ----
enum ProfClass { PC_CONV_TC = 0, PC_CONV_DIRECT, PC_GN_STATS, PC_GN_APPLY, PC_PAD_UPSAMPLE, PC_SOFTMAX, PC_LAYOUT, PC_BLEND,
                 PC_TEMPORAL, PC_ATTN, PC_ATTN_PROJ, PC_COUNT };
extern int g_prof_class_override;  // >= 0: conv launches are booked under this class (hyvae_profile_class_override)
struct ProfRec { cudaEvent_t a, b; int cls; double work; char tag[56]; };
extern bool g_prof_on;
extern double g_prof_exec_flops;  // tensor-core MACs*2 actually issued while profiling (<= algorithmic work: sub-pixel phases)
extern std::vector<ProfRec> g_prof;
extern std::vector<cudaEvent_t> g_prof_pool;
inline cudaEvent_t prof_event() {
  cudaEvent_t e;
  if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); }
  else cudaEventCreate(&e);
  return e;
}
// RAII: brackets the launches of one C-ABI call with two events.  `work` = algorithmic flops or bytes.
struct ProfScope {
  bool on; cudaStream_t st; ProfRec r;
  ProfScope(int cls, double work, void* stream, const char* tag = "", double executed = -1.0) : on(g_prof_on), st((cudaStream_t)stream) {
    if (on) {
      if ((cls == PC_CONV_TC || cls == PC_CONV_DIRECT) && g_prof_class_override >= 0) cls = g_prof_class_override;
      if (cls == PC_CONV_TC) g_prof_exec_flops += executed >= 0.0 ? executed : work;
      r.cls = cls; r.work = work; r.a = prof_event(); r.b = prof_event();
      snprintf(r.tag, sizeof(r.tag), "%s", tag);
      cudaEventRecord(r.a, st);
    }
  }
  ~ProfScope() { if (on) { cudaEventRecord(r.b, st); g_prof.push_back(r); } }
};

inline size_t dtype_size(int dt) { return dt == HYVAE_F32 ? 4 : 2; }

// Device-side view of hyvae_vol with strides precomputed (elements).
struct Vol {
  void* p;
  int B, T, H, W, C;
  int pt, ph, pw;
  int64_t sB, sT, sH, sW;  // physical strides in elements
  __host__ __device__ int64_t at(int b, int t, int h, int w) const {  // logical coords -> element offset of channel 0
    return (int64_t)b * sB + (int64_t)(t + pt) * sT + (int64_t)(h + ph) * sH + (int64_t)(w + pw) * sW;
  }
  __host__ __device__ int Tp() const { return T + pt; }
  __host__ __device__ int Hp() const { return H + 2 * ph; }
  __host__ __device__ int Wp() const { return W + 2 * pw; }
};

inline Vol make_vol(const hyvae_vol* v) {
  Vol o;
  o.p = v->data;
  o.B = v->B; o.T = v->T; o.H = v->H; o.W = v->W; o.C = v->C;
  o.pt = v->pt; o.ph = v->ph; o.pw = v->pw;
  o.sW = v->C;
  o.sH = (int64_t)(v->W + 2 * v->pw) * o.sW;
  o.sT = (int64_t)(v->H + 2 * v->ph) * o.sH;
  o.sB = (int64_t)(v->T + v->pt) * o.sT;
  return o;
}

inline int check_vol(const hyvae_vol* v, const char* name) {
  HYVAE_CHECK_ARG(v != nullptr && v->data != nullptr, "%s: null volume", name);
  HYVAE_CHECK_ARG(v->dtype == HYVAE_BF16 || v->dtype == HYVAE_F32 || v->dtype == HYVAE_F16, "%s: bad dtype %d", name, v->dtype);
  HYVAE_CHECK_ARG(v->B > 0 && v->T > 0 && v->H > 0 && v->W > 0 && v->C > 0, "%s: empty volume %dx%dx%dx%dx%d", name, v->B, v->T, v->H, v->W, v->C);
  HYVAE_CHECK_ARG(v->pt >= 0 && v->ph >= 0 && v->pw >= 0, "%s: negative halo", name);
  return HYVAE_OK;
}

// ---- scalar conversion ---------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// round-trip through the storage type (to reproduce the reference's rounding points)
template <typename T> __device__ __forceinline__ float rnd(float v) { return to_f<T>(from_f<T>(v)); }

// 16-byte (or 32-byte for fp32) vectors of 8 elements
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = a; *reinterpret_cast<float4*>(p + 4) = b; }
  __device__ __forceinline__ void get(float* f) const { f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w; }
  __device__ __forceinline__ void set(const float* f) { a = make_float4(f[0], f[1], f[2], f[3]); b = make_float4(f[4], f[5], f[6], f[7]); }
};
template <> struct Vec8<__nv_bfloat16> {
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void get(float* f) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  __device__ __forceinline__ void set(const float* f) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};
template <> struct Vec8<__half> {
  uint4 v;
  __device__ __forceinline__ void load(const __half* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void get(float* f) const {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
  }
  __device__ __forceinline__ void set(const float* f) {
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  }
};

// dtype dispatch
#define HYVAE_DISPATCH_DTYPE(dt, T, ...)                                \
  do {                                                                  \
    if ((dt) == HYVAE_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }   \
    else if ((dt) == HYVAE_F16) { using T = __half; __VA_ARGS__; }      \
    else { using T = float; __VA_ARGS__; }                              \
  } while (0)

// Per-DEVICE caches: function attributes (the >48 KB shared-memory opt-in) and the SM count belong to the device a call
// runs on, not to the process (a model moved to cuda:1, or a single-process multi-GPU driver).
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}
// SMs the persistent tensor-core conv kernels occupy (grid = conv_sms() CTAs, one per SM).  HYVAE_CONV_SMS < the SM count
// leaves SMs free on which the HBM-bound passes of ANOTHER tile stream (GroupNorm apply, halo fill) run concurrently with a
// conv instead of queueing behind it; even, >= 2.
inline int conv_sms() {
  static const int want = [] { const char* e = getenv("HYVAE_CONV_SMS"); return e ? atoi(e) : 0; }();
  const int n = num_sms();
  if (want < 2 || want >= n) return n;
  return want & ~1;
}
// `static DeviceOnce once; if (once.first()) { cudaFuncSetAttribute(...) ... once.done(); }`
struct DeviceOnce {
  bool set[kMaxDevices] = {};
  int dev = 0;
  bool first() { dev = current_device(); return !set[dev]; }
  void done() { set[dev] = true; }
};

}  // namespace hyvae
