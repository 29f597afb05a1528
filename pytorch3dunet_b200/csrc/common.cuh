// Shared helpers for libb200unet.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200unet.h"

namespace b200 {

void set_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      b200::set_error(__VA_ARGS__);          \
      return 1;                              \
    }                                        \
  } while (0)

#define B200_CHECK_LAUNCH(name)                                                     \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      b200::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
      return 2;                                                                     \
    }                                                                               \
  } while (0)

// The 16-bit ACTIVATION / tensor-core OPERAND type.  Default bf16; building with -DB200_ACT_F16 (libb200unet_f16.so) makes it IEEE fp16
// (tcgen05 kind::f16 takes either; BASELINE configs[3] names fp16).  The name `bf16` is kept for the type throughout the kernels: read
// it as "the 16-bit storage type".  fp32 accumulation, fp32 statistics and fp32 parameters are the same in both builds.
#ifdef B200_ACT_F16
#include <cuda_fp16.h>
typedef __half bf16;
#define B200_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define B200_UMMA_FMT 0u   /* kind::f16 operand format: 0 = f16, 1 = bf16 */
__host__ __device__ __forceinline__ bf16 to_act(float x) { return __float2half_rn(x); }
__host__ __device__ __forceinline__ float from_act(bf16 x) { return __half2float(x); }
#else
typedef __nv_bfloat16 bf16;
#define B200_TMAP_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define B200_UMMA_FMT 1u
__host__ __device__ __forceinline__ bf16 to_act(float x) { return __float2bfloat16_rn(x); }
__host__ __device__ __forceinline__ float from_act(bf16 x) { return __bfloat162float(x); }
#endif

// 8 bf16 = one 16-byte vector.  Held as a uint4 so that every copy / dereference is ONE 128-bit load or store
// (a struct of four __nv_bfloat162 members is copied member-wise: four 32-bit accesses).
struct alignas(16) bf16x8 {
  uint4 u;
};

#ifdef B200_ACT_F16
__device__ __forceinline__ void unpack8(const bf16x8& p, float f[8]) {
  const uint32_t w[4] = {p.u.x, p.u.y, p.u.z, p.u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);  // .x (low half) = lo
  return *reinterpret_cast<uint32_t*>(&t);
}
#else
__device__ __forceinline__ void unpack8(const bf16x8& p, float f[8]) {
  f[0] = __uint_as_float(p.u.x << 16);
  f[1] = __uint_as_float(p.u.x & 0xffff0000u);
  f[2] = __uint_as_float(p.u.y << 16);
  f[3] = __uint_as_float(p.u.y & 0xffff0000u);
  f[4] = __uint_as_float(p.u.z << 16);
  f[5] = __uint_as_float(p.u.z & 0xffff0000u);
  f[6] = __uint_as_float(p.u.w << 16);
  f[7] = __uint_as_float(p.u.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);  // .x (low half) = lo
  return *reinterpret_cast<uint32_t*>(&t);
}
#endif
__device__ __forceinline__ bf16x8 pack8(const float f[8]) {
  bf16x8 p;
  p.u.x = pack2(f[0], f[1]);
  p.u.y = pack2(f[2], f[3]);
  p.u.z = pack2(f[4], f[5]);
  p.u.w = pack2(f[6], f[7]);
  return p;
}
__device__ __forceinline__ float bf16_round(float x) { return from_act(to_act(x)); }  // round to the 16-bit storage type

// activation and its derivative expressed through the OUTPUT y (valid for relu / leaky / elu(alpha=1))
__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
  switch (act) {
    case B200_ACT_RELU: return z > 0.f ? z : 0.f;
    case B200_ACT_LEAKY: return z > 0.f ? z : z * slope;
    case B200_ACT_ELU: return z > 0.f ? z : expm1f(z);
    default: return z;
  }
}
__device__ __forceinline__ float act_grad_from_out(float y, int act, float slope) {
  switch (act) {
    case B200_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case B200_ACT_LEAKY: return y > 0.f ? 1.f : slope;
    case B200_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    default: return 1.f;
  }
}

// Border class of a voxel coordinate along one axis: 0 = low face, 1 = interior, 2 = high face, 3 = both (dim == 1)
__host__ __device__ __forceinline__ int axis_cls(int p, int dim) {
  int lo = (p == 0), hi = (p == dim - 1);
  return lo ? (hi ? 3 : 0) : (hi ? 2 : 1);
}
// is tap t (0,1,2 -> offset -1,0,+1) in bounds for axis class c
__host__ __device__ __forceinline__ bool tap_valid(int c, int t) {
  if (t == 1) return true;
  if (t == 0) return c == 1 || c == 2;  // needs p-1 >= 0
  return c == 1 || c == 0;              // needs p+1 < dim
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// number of SMs of the CURRENT device (grids of the persistent kernels are sized from it); cached per device ordinal
int sm_count();

// cycle counter for the optional per-CTA wait statistics: compiled out (constant 0) unless the library is built with -DB200_DEBUG
__device__ __forceinline__ long long dbg_clock() {
#ifdef B200_DEBUG
  return clock64();
#else
  return 0;
#endif
}

}  // namespace b200
