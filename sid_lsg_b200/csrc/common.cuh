// Shared helpers for the sm_100a kernels of the SiD-LSG hot path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define SIDLSG_OK 0
#define SIDLSG_ERR_ARG (-1)
#define SIDLSG_ERR_CUDA (-2)
#define SIDLSG_ERR_UNSUPPORTED (-3)

// dtype codes of the C ABI
#define SIDLSG_F32 0
#define SIDLSG_BF16 1

namespace sidlsg {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <class T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024; every thread gets the result. `sh` needs 33 floats.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
    r = warp_sum(r);
    if (lane == 0) sh[32] = r;
  }
  __syncthreads();
  return sh[32];
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float silu_grad_f(float x) {
  float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace sidlsg

#define SID_DISPATCH_DTYPE(dtype, T, ...)                                  \
  do {                                                                     \
    if ((dtype) == SIDLSG_F32) { typedef float T; __VA_ARGS__; }           \
    else if ((dtype) == SIDLSG_BF16) { typedef sidlsg::bf16 T; __VA_ARGS__; } \
    else { sidlsg::set_error("bad dtype %d", (int)(dtype)); return SIDLSG_ERR_ARG; } \
  } while (0)
