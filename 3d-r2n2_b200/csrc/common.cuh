// Shared helpers for the sm_100a kernels of libr2n2_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <utility>
#include "../../include/r2n2_b200.h"

namespace r2n2 {

void set_error(const char* fmt, ...);

#define R2N2_CHECK_ARG(cond, code, ...)      \
  do {                                       \
    if (!(cond)) {                           \
      r2n2::set_error(__VA_ARGS__);          \
      return (code);                         \
    }                                        \
  } while (0)

#define R2N2_CHECK_LAUNCH(name)                                                  \
  do {                                                                           \
    cudaError_t e__ = cudaPeekAtLastError();                                     \
    if (e__ != cudaSuccess) {                                                    \
      r2n2::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return R2N2_ERR_CUDA;                                                      \
    }                                                                            \
  } while (0)

constexpr float kLeak = 0.01f;  // LeakyReLU leakiness, lib/layers.py:630

// operands of a fused conv-GRU gate epilogue (r2n2_conv_fwd_gru): mode = R2N2_GRU_*
struct GruEpi {
  int mode;
  const void* in[4];
  void* out[3];
};

// ---- 4-element vector load/store with conversion to/from float -------------------------
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) {
    return *reinterpret_cast<const float4*>(p);
  }
  static __device__ __forceinline__ void store(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
  }
};
template <>
struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
    uint2 raw = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t*>(&a);
    raw.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = raw;
  }
};

// ---- 16-byte vector of channels: 4 x fp32 or 8 x bf16, widened to float --------------------------
template <typename T>
struct Pack;
template <>
struct Pack<float> {
  static constexpr int N = 4;
  float v[4];
  static __device__ __forceinline__ Pack load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    Pack r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Pack<__nv_bfloat16> {
  static constexpr int N = 8;
  float v[8];
  static __device__ __forceinline__ Pack load(const __nv_bfloat16* p) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    Pack r;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
      float2 f = __bfloat1622float2(h);
      r.v[2 * j] = f.x; r.v[2 * j + 1] = f.y;
    }
    return r;
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      w[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <typename T>
__device__ __forceinline__ float to_float(T v);
template <>
__device__ __forceinline__ float to_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T>
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ float from_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// max(v, 0.01 v) == (v > 0 ? v : 0.01 v) for every finite v (two instructions instead of compare + multiply + select)
__device__ __forceinline__ float lrelu(float v) { return fmaxf(v, kLeak * v); }
__device__ __forceinline__ float lrelu_slope(float a) { return a > 0.f ? 1.f : kLeak; }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline int& sm_limit_ref() {
  static int limit = 0;          // r2n2_set_sm_limit (0 = none)
  return limit;
}
inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  const int lim = sm_limit_ref();
  return (lim >= 16 && lim < n) ? (lim & ~1) : n;
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// Every kernel of the library is launched with programmaticStreamSerialization: the next kernel of the stream may be
// scheduled while this one still runs (its CTAs land on SMs as ours exit), run its prologue (barrier init, TMEM
// allocation, shared-memory zero fill, descriptor fetch) and then block in griddepcontrol.wait until this grid has
// completed and its memory is visible.  Hides the launch latency + prologue of ~170 launches per training step.
// RULE: a kernel executes pdl_wait() before its FIRST global-memory access (read or write).  R2N2_PDL=0 switches the
// launch attribute off (then griddepcontrol.* are no-ops).
static __device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
static __device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("R2N2_PDL");
    on = (e && atoi(e) == 0) ? 0 : 1;
  }
  return on != 0;
}
// <<<grid, block, smem, stream>>> with the PDL attribute (and an optional cluster size along x)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_x > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)cluster_x;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// dispatch on dtype code
#define R2N2_DISPATCH_DTYPE(dtype, T, ...)                 \
  if ((dtype) == R2N2_F32) {                               \
    using T = float;                                       \
    __VA_ARGS__                                            \
  } else if ((dtype) == R2N2_BF16) {                       \
    using T = __nv_bfloat16;                               \
    __VA_ARGS__                                            \
  } else {                                                 \
    r2n2::set_error("bad dtype code %d", (int)(dtype));    \
    return R2N2_ERR_BAD_DTYPE;                             \
  }

}  // namespace r2n2
