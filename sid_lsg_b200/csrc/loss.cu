// Fused loss kernels of the SiD-LSG step: forward value AND input gradients in one launch.
//   fake-score loss   /root/reference/training/sid_training_loop.py:423-445
//   LSG generator loss /root/reference/training/sid_training_loop.py:508-530
// Row (= sample) semantics of the reference are kept without changing shapes: a row holding a NaN in any
// input is dropped from the sum and receives zero gradient.
//
// One thread-block CLUSTER of 8 CTAs owns a row; each CTA streams its 1/8 slice once from HBM with 128-bit
// loads for the row statistics (NaN flag, sum |x_g - y_real|), the partials are combined over distributed
// shared memory, and the slice is revisited from L1 for the loss terms and the three gradients.  HBM traffic
// is the algorithmic 6 x 4 B per element (3 reads + 3 writes); HBM-bound, below ~126 MB it is L2/latency-bound.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace sidlsg {

constexpr int LOSS_CLUSTER = 8;
constexpr int LOSS_THREADS = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float nanflag(float4 v) {
  return (isnan(v.x) || isnan(v.y) || isnan(v.z) || isnan(v.w)) ? 1.f : 0.f;
}

// combine (sum, flag) over the 8 CTAs of the cluster; every thread gets the totals
__device__ __forceinline__ void cluster_combine(float& sum, float& flag, float* sh, float* part) {
  cg::cluster_group cluster = cg::this_cluster();
  sum = block_sum(sum, sh);
  flag = block_sum(flag, sh);
  if (threadIdx.x == 0) { part[0] = sum; part[1] = flag; }
  cluster.sync();
  float s = 0.f, f = 0.f;
  for (int r = 0; r < LOSS_CLUSTER; ++r) {
    const float* rp = cluster.map_shared_rank(part, r);
    s += rp[0];
    f += rp[1];
  }
  cluster.sync();  // nobody leaves (or reuses `part`) while a peer may still be reading it
  sum = s;
  flag = f;
}

// out[0] += sum over valid rows of (e-n)^2 * scale ; out[1] += number of valid rows ; grad = 2 (e-n) scale
__global__ void __cluster_dims__(LOSS_CLUSTER, 1, 1) __launch_bounds__(LOSS_THREADS)
fake_loss_kernel(const float* __restrict__ e, const float* __restrict__ n, float* __restrict__ grad,
                 float* __restrict__ out, int CHW, float scale) {
  __shared__ float sh[33];
  __shared__ float part[2];
  const int row = blockIdx.y;
  const int nvec = CHW / 4;
  const int per = (nvec + LOSS_CLUSTER - 1) / LOSS_CLUSTER;
  const int v0 = blockIdx.x * per, v1 = min(nvec, v0 + per);
  const float* er = e + (long)row * CHW;
  const float* nr = n + (long)row * CHW;
  float flag = 0.f, dummy = 0.f;
  for (int v = v0 + threadIdx.x; v < v1; v += LOSS_THREADS) flag += nanflag(ld4(er + 4 * v));
  cluster_combine(dummy, flag, sh, part);
  const bool valid = flag == 0.f;
  float acc = 0.f;
  float* gr = grad ? grad + (long)row * CHW : nullptr;
  for (int v = v0 + threadIdx.x; v < v1; v += LOSS_THREADS) {
    float4 a = ld4(er + 4 * v), b = ld4(nr + 4 * v);
    float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    if (valid) acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
    if (gr) {
      float s2 = valid ? 2.f * scale : 0.f;
      st4(gr + 4 * v, valid ? make_float4(s2 * d.x, s2 * d.y, s2 * d.z, s2 * d.w) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    if (valid) atomicAdd(&out[0], acc * scale);
    if (valid && blockIdx.x == 0) atomicAdd(&out[1], 1.f);
  }
}

// w = max(mean|xg - yr|, 1e-5) (no gradient); l = (yr-yf) * ((yr-xg) - alpha (yr-yf)) / w
// d/dxg = -(yr-yf) s/w ; d/dyf = (2 alpha (yr-yf) - (yr-xg)) s/w ; d/dyr = ((yr-xg) + (1-2 alpha)(yr-yf)) s/w
__global__ void __cluster_dims__(LOSS_CLUSTER, 1, 1) __launch_bounds__(LOSS_THREADS)
lsg_loss_kernel(const float* __restrict__ xg, const float* __restrict__ yreal, const float* __restrict__ yfake,
                float* __restrict__ dxg, float* __restrict__ dyreal, float* __restrict__ dyfake,
                float* __restrict__ out, int CHW, float alpha, float scale) {
  __shared__ float sh[33];
  __shared__ float part[2];
  const int row = blockIdx.y;
  const int nvec = CHW / 4;
  const int per = (nvec + LOSS_CLUSTER - 1) / LOSS_CLUSTER;
  const int v0 = blockIdx.x * per, v1 = min(nvec, v0 + per);
  const long ro = (long)row * CHW;
  float flag = 0.f, asum = 0.f;
  for (int v = v0 + threadIdx.x; v < v1; v += LOSS_THREADS) {
    float4 x = ld4(xg + ro + 4 * v), r = ld4(yreal + ro + 4 * v), f = ld4(yfake + ro + 4 * v);
    flag += nanflag(x) + nanflag(r) + nanflag(f);
    asum += fabsf(x.x - r.x) + fabsf(x.y - r.y) + fabsf(x.z - r.z) + fabsf(x.w - r.w);
  }
  cluster_combine(asum, flag, sh, part);
  const bool valid = flag == 0.f;
  const float w = fmaxf(asum / (float)CHW, 1e-5f);
  const float sw = valid ? scale / w : 0.f;
  float acc = 0.f;
  const bool want_grad = dxg != nullptr;
  for (int v = v0 + threadIdx.x; v < v1; v += LOSS_THREADS) {
    float4 x = ld4(xg + ro + 4 * v), r = ld4(yreal + ro + 4 * v), f = ld4(yfake + ro + 4 * v);
    float4 gx, gr, gf;
#define LSG_TERM(c)                                              \
    {                                                            \
      float rf = r.c - f.c, rx = r.c - x.c;                      \
      float inner = rx - alpha * rf;                             \
      if (valid) acc += rf * inner;                              \
      gx.c = valid ? -rf * sw : 0.f;                             \
      gf.c = valid ? (2.f * alpha * rf - rx) * sw : 0.f;         \
      gr.c = valid ? (rx + (1.f - 2.f * alpha) * rf) * sw : 0.f; \
    }
    LSG_TERM(x) LSG_TERM(y) LSG_TERM(z) LSG_TERM(w)
#undef LSG_TERM
    if (want_grad) {
      st4(dxg + ro + 4 * v, gx);
      st4(dyreal + ro + 4 * v, gr);
      st4(dyfake + ro + 4 * v, gf);
    }
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    if (valid) atomicAdd(&out[0], acc * scale / w);
    if (valid && blockIdx.x == 0) atomicAdd(&out[1], 1.f);
  }
}

}  // namespace sidlsg

using namespace sidlsg;

// out: float[2] = {loss, valid_rows}, zeroed here.  grad may be null (value only).
extern "C" int sidlsg_fake_loss(const float* eps_hat, const float* noise, float* grad, float* out, int B, int CHW,
                                float scale, void* stream) {
  if (CHW % 4) { set_error("fake_loss: CHW=%d not a multiple of 4", CHW); return SIDLSG_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, 2 * sizeof(float), st);
  if (B == 0) return SIDLSG_OK;
  fake_loss_kernel<<<dim3(LOSS_CLUSTER, B), LOSS_THREADS, 0, st>>>(eps_hat, noise, grad, out, CHW, scale);
  return check_launch("fake_loss");
}

// gradients may all be null (value only) or all non-null
extern "C" int sidlsg_lsg_loss(const float* xg, const float* yreal, const float* yfake, float* dxg, float* dyreal,
                               float* dyfake, float* out, int B, int CHW, float alpha, float scale, void* stream) {
  if (CHW % 4) { set_error("lsg_loss: CHW=%d not a multiple of 4", CHW); return SIDLSG_ERR_ARG; }
  if ((dxg == nullptr) != (dyreal == nullptr) || (dxg == nullptr) != (dyfake == nullptr)) {
    set_error("lsg_loss: pass all three gradient buffers or none"); return SIDLSG_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, 2 * sizeof(float), st);
  if (B == 0) return SIDLSG_OK;
  lsg_loss_kernel<<<dim3(LOSS_CLUSTER, B), LOSS_THREADS, 0, st>>>(xg, yreal, yfake, dxg, dyreal, dyfake, out, CHW, alpha, scale);
  return check_launch("lsg_loss");
}
