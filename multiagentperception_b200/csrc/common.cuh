// Internal helpers shared by the libw2c translation units: error reporting, launch accounting, activation
// load/store in the two storage formats (bf16 and bf16 hi|lo planes).
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/w2c.h"

namespace w2c {

int set_error(int code, const char* fmt, ...);
void count_launch(unsigned n = 1);

// stem_tc.cu: tensor-core first layer (cout 64 / 128)
int stem3x3_tc_forward(const float* x, const float* w, const float* scale, const float* shift, void* y, int b,
                       int n_agents, int c_total, int c_first, int h, int wpx, int cout, int act, int n_split,
                       cudaStream_t stream);

int stem3x3_tc_u8_forward(const uint8_t* x, const float* lut, const float* w, const float* scale, const float* shift,
                          void* y, int b, int n_agents, int agents_total, int agent_first, int h, int wpx, int cout,
                          int act, int n_split, cudaStream_t stream);

// 7x7 stride-2 first layer of the resnet18 trunk on the tensor cores (x: fp32 NCHW views, or uint8 frames with u8 = 1)
int stem7x7_tc_forward(const void* x, const float* lut, const float* w, const float* scale, const float* shift, void* y,
                       int b, int n_agents, int c_total, int c_first, int h, int wpx, int cout, int act, int n_split,
                       int u8, cudaStream_t stream);

#define W2C_CHECK_ARG(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return ::w2c::set_error(W2C_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define W2C_CHECK_LAUNCH(what)                                                                  \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess)                                                                     \
      return ::w2c::set_error(W2C_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e__));           \
    ::w2c::count_launch();                                                                      \
  } while (0)

// storage formats (include/w2c.h): one plane (BF16, FP16) or two planes [hi | lo] per pixel (BF16X2, FP16X2)
__host__ __device__ __forceinline__ int act_planes(int act) {
  return (act == W2C_ACT_BF16X2 || act == W2C_ACT_FP16X2) ? 2 : 1;
}
__host__ __device__ __forceinline__ bool act_is_f16(int act) { return act == W2C_ACT_FP16 || act == W2C_ACT_FP16X2; }
__host__ __device__ __forceinline__ bool act_valid(int act) { return act >= W2C_ACT_BF16 && act <= W2C_ACT_FP16X2; }
// MMA passes over the planes: 1 = hi*hi; 3 = hi*hi + hi*lo + lo*hi (two-plane formats only). passes = 0 picks the
// format's default (3 for two planes).
__host__ __device__ __forceinline__ int act_passes(int act, int passes) {
  return act_planes(act) == 1 ? 1 : (passes == 1 ? 1 : 3);
}

// 16-bit storage element <-> float in either storage type (the buffers are typed __nv_bfloat16 for addressing only)
__device__ __forceinline__ float elem_to_float(__nv_bfloat16 v, bool f16) {
  return f16 ? __half2float(*reinterpret_cast<const __half*>(&v)) : __bfloat162float(v);
}
__device__ __forceinline__ __nv_bfloat16 float_to_elem(float v, bool f16) {
  if (f16) {
    const __half h = __float2half_rn(v);
    return *reinterpret_cast<const __nv_bfloat16*>(&h);
  }
  return __float2bfloat16_rn(v);
}
// a packed pair of storage elements -> two floats
__device__ __forceinline__ float2 unpack_act2(uint32_t bits, bool f16) {
  return f16 ? __half22float2(*reinterpret_cast<const __half2*>(&bits))
             : __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bits));
}

// (hi, lo) split of a pair of values in either element type: hi = rn(v), lo = rn(v - hi)
__device__ __forceinline__ void split_act2(float a, float b, bool f16, uint32_t& hi, uint32_t& lo) {
  if (f16) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
}

// value = hi (+ lo).  Pixel layout of the two-plane formats: [hi: cstride channels][lo: cstride channels].
__device__ __forceinline__ float act_load(const __nv_bfloat16* pix, int c, int cstride, int act) {
  const bool f16 = act_is_f16(act);
  float v = elem_to_float(pix[c], f16);
  if (act_planes(act) == 2) v += elem_to_float(pix[cstride + c], f16);
  return v;
}
__device__ __forceinline__ void act_store(__nv_bfloat16* pix, int c, int cstride, int act, float v) {
  const bool f16 = act_is_f16(act);
  const __nv_bfloat16 hi = float_to_elem(v, f16);
  pix[c] = hi;
  if (act_planes(act) == 2) pix[cstride + c] = float_to_elem(v - elem_to_float(hi, f16), f16);
}

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- per-device launch state. cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to a DEVICE,
// and one process may drive several (nn.DataParallel replicas, a model moved with .to('cuda:1'), train.py:177): every
// launcher keeps its one-time setup per device ordinal, never per process.
constexpr int kMaxDevices = 64;

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 || dev >= kMaxDevices ? 0 : dev;
}

struct DeviceOnce {
  std::atomic<int> done[kMaxDevices];
  DeviceOnce() {
    for (auto& d : done) d.store(0, std::memory_order_relaxed);
  }
  // Runs fn() (returning cudaError_t) once per device; racing threads may both run it, which is harmless for an
  // idempotent attribute write.
  template <typename Fn>
  int ensure(Fn fn, const char* what) {
    const int dev = current_device();
    if (done[dev].load(std::memory_order_acquire)) return W2C_OK;
    const cudaError_t e = fn();
    if (e != cudaSuccess) return set_error(W2C_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    done[dev].store(1, std::memory_order_release);
    return W2C_OK;
  }
};

// SM count of the current device (cached per device)
inline int device_sm_count() {
  static std::atomic<int> sms[kMaxDevices];
  const int dev = current_device();
  int n = sms[dev].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

}  // namespace w2c
