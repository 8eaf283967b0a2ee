// Row softmax forward/backward for the fp32-exact attention path (scores materialised per batch chunk).
// Attention semantics: diffusers Attention / AttnProcessor: softmax(Q K^T / sqrt(d)) V, no mask
// (SURVEY.md App. A-2).  One warp per row for short rows (cross-attention, 77 keys), one block per row otherwise.
#include "common.cuh"

namespace sidlsg {

// P = softmax(scale * S) row-wise, in place allowed (P may alias S)
template <class TI, class T>
__global__ void __launch_bounds__(256)
softmax_fwd_kernel(const TI* __restrict__ S, T* __restrict__ P, long rows, int cols, long ld, float scale,
                   int warp_per_row) {
  __shared__ float sh[33];
  if (warp_per_row) {
    const int lane = threadIdx.x & 31;
    long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const TI* s = S + r * ld;
    T* p = P + r * ld;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, to_f(s[c]) * scale);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) sum += expf(to_f(s[c]) * scale - mx);
    sum = warp_sum(sum);
    float inv = 1.f / sum;
    for (int c = lane; c < cols; c += 32) p[c] = from_f<T>(expf(to_f(s[c]) * scale - mx) * inv);
  } else {
    long r = blockIdx.x;
    const TI* s = S + r * ld;
    T* p = P + r * ld;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, to_f(s[c]) * scale);
    mx = warp_max(mx);
    {
      int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      __syncthreads();
      if (lane == 0) sh[w] = mx;
      __syncthreads();
      float m2 = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : -INFINITY;
      if (w == 0) {
        m2 = warp_max(m2);
        if (lane == 0) sh[32] = m2;
      }
      __syncthreads();
      mx = sh[32];
    }
    float sum = 0.f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) sum += expf(to_f(s[c]) * scale - mx);
    sum = block_sum(sum, sh);
    float inv = 1.f / sum;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) p[c] = from_f<T>(expf(to_f(s[c]) * scale - mx) * inv);
  }
}

// dS = scale * P * (dP - sum_j dP_j P_j), in place allowed (dS may alias dP)
template <class TI, class T>
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const T* __restrict__ P, const TI* __restrict__ dP, T* __restrict__ dS, long rows, int cols,
                   long ld, float scale, int warp_per_row) {
  __shared__ float sh[33];
  if (warp_per_row) {
    const int lane = threadIdx.x & 31;
    long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const T* p = P + r * ld;
    const TI* dp = dP + r * ld;
    T* ds = dS + r * ld;
    float dot = 0.f;
    for (int c = lane; c < cols; c += 32) dot = fmaf(to_f(p[c]), to_f(dp[c]), dot);
    dot = warp_sum(dot);
    for (int c = lane; c < cols; c += 32) ds[c] = from_f<T>(scale * to_f(p[c]) * (to_f(dp[c]) - dot));
  } else {
    long r = blockIdx.x;
    const T* p = P + r * ld;
    const TI* dp = dP + r * ld;
    T* ds = dS + r * ld;
    float dot = 0.f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) dot = fmaf(to_f(p[c]), to_f(dp[c]), dot);
    dot = block_sum(dot, sh);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) ds[c] = from_f<T>(scale * to_f(p[c]) * (to_f(dp[c]) - dot));
  }
}

}  // namespace sidlsg

using namespace sidlsg;

// S / dP are fp32 (the score GEMMs always accumulate and store fp32); P / dS are in `dtype` (fp32 or bf16)
// P may alias S and dS may alias dP only when dtype is fp32.  ld = row stride (elements) of every matrix, >= cols:
// rows padded to a multiple of 8 elements keep bf16 P / dS usable as TMA operands of the tensor-core GEMMs.
extern "C" int sidlsg_softmax_fwd(const float* S, void* P, long rows, int cols, long ld, float scale, int dtype,
                                  void* stream) {
  if (rows == 0 || cols == 0) return SIDLSG_OK;
  if (ld < cols) { set_error("softmax_fwd: ld < cols"); return SIDLSG_ERR_ARG; }
  int wpr = cols <= 256;
  long blocks = wpr ? (rows + 7) / 8 : rows;
  SID_DISPATCH_DTYPE(dtype, T, (softmax_fwd_kernel<float, T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(S, (T*)P, rows, cols, ld, scale, wpr)));
  return check_launch("softmax_fwd");
}
extern "C" int sidlsg_softmax_bwd(const void* P, const float* dP, void* dS, long rows, int cols, long ld, float scale,
                                  int dtype, void* stream) {
  if (rows == 0 || cols == 0) return SIDLSG_OK;
  if (ld < cols) { set_error("softmax_bwd: ld < cols"); return SIDLSG_ERR_ARG; }
  int wpr = cols <= 256;
  long blocks = wpr ? (rows + 7) / 8 : rows;
  SID_DISPATCH_DTYPE(dtype, T, (softmax_bwd_kernel<float, T><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const T*)P, dP, (T*)dS, rows, cols, ld, scale, wpr)));
  return check_launch("softmax_bwd");
}
