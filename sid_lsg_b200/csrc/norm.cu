// GroupNorm(+SiLU) and LayerNorm, forward and backward, on token-major (NHWC) activations.
// HBM-bound kernels: coalesced channel-fastest access, fp32 statistics (double at the combine step).
// GroupNorm semantics: diffusers ResnetBlock2D norm1/norm2 + SiLU, Transformer2DModel.norm (eps 1e-6),
// conv_norm_out.
#include "common.cuh"

namespace sidlsg {

constexpr int GN_THREADS = 256;
constexpr int GN_MAXSLOT = 10;  // C <= 2560

// ---- GroupNorm statistics: sums[b][c][2] partial per-channel sums (double) -------------------------
template <class T>
__global__ void __launch_bounds__(GN_THREADS)
gn_partial_kernel(const T* __restrict__ x, double* __restrict__ sums, int HW, int C, int rows_per_block) {
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  float s[GN_MAXSLOT], q[GN_MAXSLOT];
#pragma unroll
  for (int i = 0; i < GN_MAXSLOT; ++i) s[i] = q[i] = 0.f;
  const T* xb = x + (long)b * HW * C;
  for (int r = r0; r < r1; ++r) {
    const T* row = xb + (long)r * C;
#pragma unroll
    for (int i = 0; i < GN_MAXSLOT; ++i) {
      int c = threadIdx.x + i * GN_THREADS;
      if (c < C) {
        float v = to_f(row[c]);
        s[i] += v;
        q[i] = fmaf(v, v, q[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < GN_MAXSLOT; ++i) {
    int c = threadIdx.x + i * GN_THREADS;
    if (c < C) {
      atomicAdd(&sums[((long)b * C + c) * 2 + 0], (double)s[i]);
      atomicAdd(&sums[((long)b * C + c) * 2 + 1], (double)q[i]);
    }
  }
}

// one thread per (b, c): group statistics -> per-channel affine  y = x*a + sh
__global__ void gn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, float* __restrict__ a, float* __restrict__ sh,
                                   int B, int C, int G, int HW, float eps) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  int b = idx / C, c = idx - b * C;
  int cpg = C / G, g = c / cpg;
  double s = 0, q = 0;
  for (int j = 0; j < cpg; ++j) {
    s += sums[((long)b * C + g * cpg + j) * 2 + 0];
    q += sums[((long)b * C + g * cpg + j) * 2 + 1];
  }
  double n = (double)HW * cpg;
  double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0) var = 0;
  float rstd = (float)(1.0 / sqrt(var + (double)eps));
  float mf = (float)mean;
  if (c == g * cpg) { mean_out[b * G + g] = mf; rstd_out[b * G + g] = rstd; }
  float ga = gamma[c];
  a[idx] = rstd * ga;
  sh[idx] = beta[c] - mf * rstd * ga;
}

template <class TI, class TO, int VEC>
__global__ void gn_apply_kernel(const TI* __restrict__ x, const float* __restrict__ a, const float* __restrict__ sh,
                                TO* __restrict__ y, long total_vec, int HWC, int C, int silu) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  long e0 = i * VEC;
  int b = (int)(e0 / HWC);
  int c0 = (int)(e0 % C);
  const float* ab = a + (long)b * C + c0;
  const float* sb = sh + (long)b * C + c0;
  TI xin[VEC];
  TO out[VEC];
  if (VEC * sizeof(TI) == 16) *(uint4*)xin = *(const uint4*)(x + e0);
  else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) xin[j] = x[e0 + j];
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    float v = fmaf(to_f(xin[j]), ab[j], sb[j]);
    if (silu) v = silu_f(v);
    out[j] = from_f<TO>(v);
  }
  if (VEC * sizeof(TO) == 16) *(uint4*)(y + e0) = *(uint4*)out;
  else if (VEC * sizeof(TO) == 32) { ((uint4*)(y + e0))[0] = ((uint4*)out)[0]; ((uint4*)(y + e0))[1] = ((uint4*)out)[1]; }
  else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) y[e0 + j] = out[j];
  }
}

// ---- GroupNorm backward ------------------------------------------------------------------------------
// partial per-(b,c) sums of dz and dz*xhat ; dz = dy * silu'(pre) (pre = x*a+sh) or dy
template <class T>
__global__ void __launch_bounds__(GN_THREADS)
gn_bwd_partial_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ a,
                      const float* __restrict__ sh, const float* __restrict__ mean, const float* __restrict__ rstd,
                      double* __restrict__ sums, int HW, int C, int G, int rows_per_block, int silu) {
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  const int cpg = C / G;
  float s[GN_MAXSLOT], q[GN_MAXSLOT], av[GN_MAXSLOT], sv[GN_MAXSLOT], mu[GN_MAXSLOT], rs[GN_MAXSLOT];
#pragma unroll
  for (int i = 0; i < GN_MAXSLOT; ++i) {
    s[i] = q[i] = 0.f;
    int c = threadIdx.x + i * GN_THREADS;
    if (c < C) {
      av[i] = a[(long)b * C + c]; sv[i] = sh[(long)b * C + c];
      mu[i] = mean[b * G + c / cpg]; rs[i] = rstd[b * G + c / cpg];
    } else { av[i] = sv[i] = mu[i] = rs[i] = 0.f; }
  }
  const long base = (long)b * HW * C;
  for (int r = r0; r < r1; ++r) {
    const long ro = base + (long)r * C;
#pragma unroll
    for (int i = 0; i < GN_MAXSLOT; ++i) {
      int c = threadIdx.x + i * GN_THREADS;
      if (c < C) {
        float xv = to_f(x[ro + c]);
        float dz = to_f(dy[ro + c]);
        if (silu) dz *= silu_grad_f(fmaf(xv, av[i], sv[i]));
        s[i] += dz;
        q[i] = fmaf(dz, (xv - mu[i]) * rs[i], q[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < GN_MAXSLOT; ++i) {
    int c = threadIdx.x + i * GN_THREADS;
    if (c < C) {
      atomicAdd(&sums[((long)b * C + c) * 2 + 0], (double)s[i]);
      atomicAdd(&sums[((long)b * C + c) * 2 + 1], (double)q[i]);
    }
  }
}

// per (b,c): coefficients of dx = A*dz + x*P + Q
__global__ void gn_bwd_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                       float* __restrict__ P, float* __restrict__ Q, int B, int C, int G, int HW) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  int b = idx / C, c = idx - b * C;
  int cpg = C / G, g = c / cpg;
  double S1 = 0, S2 = 0;
  for (int j = 0; j < cpg; ++j) {
    int cc = g * cpg + j;
    double ga = gamma[cc];
    S1 += ga * sums[((long)b * C + cc) * 2 + 0];
    S2 += ga * sums[((long)b * C + cc) * 2 + 1];
  }
  double n = (double)HW * cpg;
  double m = mean[b * G + g], r = rstd[b * G + g];
  P[idx] = (float)(-r * r * S2 / n);
  Q[idx] = (float)(-r * S1 / n + m * r * r * S2 / n);
}

// dgamma[c] (+)= sum_b sums[b,c,1] ; dbeta[c] (+)= sum_b sums[b,c,0]
__global__ void gn_bwd_param_kernel(const double* __restrict__ sums, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int B, int C, int accumulate) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0, q = 0;
  for (int b = 0; b < B; ++b) {
    s += sums[((long)b * C + c) * 2 + 0];
    q += sums[((long)b * C + c) * 2 + 1];
  }
  if (accumulate) { dgamma[c] += (float)q; dbeta[c] += (float)s; }
  else { dgamma[c] = (float)q; dbeta[c] = (float)s; }
}

template <class T, int VEC>
__global__ void gn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ a,
                                    const float* __restrict__ sh, const float* __restrict__ P,
                                    const float* __restrict__ Q, T* __restrict__ dx, long total_vec, int HWC, int C,
                                    int silu) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  long e0 = i * VEC;
  int b = (int)(e0 / HWC);
  int c0 = (int)(e0 % C);
  long o = (long)b * C + c0;
  T xin[VEC], din[VEC], out[VEC];
  *(uint4*)xin = *(const uint4*)(x + e0);
  *(uint4*)din = *(const uint4*)(dy + e0);
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    float xv = to_f(xin[j]);
    float dz = to_f(din[j]);
    float av = a[o + j];
    if (silu) dz *= silu_grad_f(fmaf(xv, av, sh[o + j]));
    out[j] = from_f<T>(fmaf(av, dz, fmaf(xv, P[o + j], Q[o + j])));
  }
  *(uint4*)(dx + e0) = *(uint4*)out;
}

// ---- LayerNorm ---------------------------------------------------------------------------------------
constexpr int LN_MAXSLOT_MAX = 40;  // C <= 1280

template <class T, int LN_MAXSLOT>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long rows, int C,
              float eps) {
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  for (long r = warp; r < rows; r += nwarps) {
    const T* row = x + r * C;
    float v[LN_MAXSLOT];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXSLOT; ++i) {
      int c = lane + 32 * i;
      v[i] = (c < C) ? to_f(row[c]) : 0.f;
      s += v[i];
    }
    float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXSLOT; ++i) {
      int c = lane + 32 * i;
      float d = (c < C) ? v[i] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    float rstd = rsqrtf(warp_sum(q) / C + eps);
    if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
    T* yr = y + r * C;
#pragma unroll
    for (int i = 0; i < LN_MAXSLOT; ++i) {
      int c = lane + 32 * i;
      if (c < C) yr[c] = from_f<T>(fmaf((v[i] - mean) * rstd, gamma[c], beta[c]));
    }
  }
}

template <class T, int LN_MAXSLOT>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ dx,
              float* __restrict__ dgamma, float* __restrict__ dbeta, long rows, int C, int want_param_grad) {
  extern __shared__ float red[];  // [2][C]
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  float dg[LN_MAXSLOT], db[LN_MAXSLOT], gm[LN_MAXSLOT];
#pragma unroll
  for (int i = 0; i < LN_MAXSLOT; ++i) {
    dg[i] = db[i] = 0.f;
    int c = lane + 32 * i;
    gm[i] = (c < C) ? gamma[c] : 0.f;
  }
  for (long r = warp; r < rows; r += nwarps) {
    const float mu = mean[r], rs = rstd[r];
    float xh[LN_MAXSLOT], g[LN_MAXSLOT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXSLOT; ++i) {
      int c = lane + 32 * i;
      if (c < C) {
        float d = to_f(dy[r * C + c]);
        xh[i] = (to_f(x[r * C + c]) - mu) * rs;
        g[i] = d * gm[i];
        dg[i] = fmaf(d, xh[i], dg[i]);
        db[i] += d;
        s1 += g[i];
        s2 = fmaf(g[i], xh[i], s2);
      } else { xh[i] = g[i] = 0.f; }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
#pragma unroll
    for (int i = 0; i < LN_MAXSLOT; ++i) {
      int c = lane + 32 * i;
      if (c < C) dx[r * C + c] = from_f<T>(rs * (g[i] - s1 - xh[i] * s2));
    }
  }
  if (!want_param_grad) return;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < LN_MAXSLOT; ++i) {
    int c = lane + 32 * i;
    if (c < C) { atomicAdd(&red[c], dg[i]); atomicAdd(&red[C + c], db[i]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&dgamma[c], red[c]);
    atomicAdd(&dbeta[c], red[C + c]);
  }
}

}  // namespace sidlsg

using namespace sidlsg;

static int gn_rows_per_block(int HW, int B) {
  int rpb = 64;
  while (rpb > 8 && (long)cdiv(HW, rpb) * B < 592) rpb >>= 1;
  return rpb;
}

// workspace: double[B*C*2] sums (zeroed here).  Outputs mean/rstd [B,G] and per-(b,c) affine a/sh [B,C].
extern "C" int sidlsg_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                    float* rstd, float* a, float* sh, double* ws, int B, int HW, int C, int G,
                                    float eps, int silu, int in_dtype, int out_dtype, void* stream) {
  if (C > GN_THREADS * GN_MAXSLOT || C % G || C % 8) { set_error("groupnorm: unsupported C=%d G=%d", C, G); return SIDLSG_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(ws, 0, sizeof(double) * 2 * (size_t)B * C, st);
  int rpb = gn_rows_per_block(HW, B);
  dim3 grid(cdiv(HW, rpb), B);
  if (in_dtype == SIDLSG_F32) gn_partial_kernel<float><<<grid, GN_THREADS, 0, st>>>((const float*)x, ws, HW, C, rpb);
  else gn_partial_kernel<bf16><<<grid, GN_THREADS, 0, st>>>((const bf16*)x, ws, HW, C, rpb);
  gn_finalize_kernel<<<cdiv((long)B * C, 256), 256, 0, st>>>(ws, gamma, beta, mean, rstd, a, sh, B, C, G, HW, eps);
  long total = (long)B * HW * C;
  if (in_dtype == SIDLSG_F32 && out_dtype == SIDLSG_F32)
    gn_apply_kernel<float, float, 4><<<cdiv(total / 4, 256), 256, 0, st>>>((const float*)x, a, sh, (float*)y, total / 4, HW * C, C, silu);
  else if (in_dtype == SIDLSG_BF16 && out_dtype == SIDLSG_BF16)
    gn_apply_kernel<bf16, bf16, 8><<<cdiv(total / 8, 256), 256, 0, st>>>((const bf16*)x, a, sh, (bf16*)y, total / 8, HW * C, C, silu);
  else if (in_dtype == SIDLSG_F32 && out_dtype == SIDLSG_BF16)
    gn_apply_kernel<float, bf16, 4><<<cdiv(total / 4, 256), 256, 0, st>>>((const float*)x, a, sh, (bf16*)y, total / 4, HW * C, C, silu);
  else { set_error("groupnorm: unsupported dtype pair"); return SIDLSG_ERR_UNSUPPORTED; }
  return check_launch("groupnorm_fwd");
}

// ws: double[B*C*2]; P,Q: float[B*C] scratch.  dgamma/dbeta fp32 (accumulate flag), may be null.
extern "C" int sidlsg_groupnorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                                    const float* rstd, const float* a, const float* sh, void* dx, float* dgamma,
                                    float* dbeta, double* ws, float* P, float* Q, int B, int HW, int C, int G,
                                    int silu, int accumulate, int dtype, void* stream) {
  if (C > GN_THREADS * GN_MAXSLOT || C % G || C % 8) { set_error("groupnorm_bwd: unsupported C=%d G=%d", C, G); return SIDLSG_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(ws, 0, sizeof(double) * 2 * (size_t)B * C, st);
  int rpb = gn_rows_per_block(HW, B);
  dim3 grid(cdiv(HW, rpb), B);
  long total = (long)B * HW * C;
  if (dtype == SIDLSG_F32) {
    gn_bwd_partial_kernel<float><<<grid, GN_THREADS, 0, st>>>((const float*)dy, (const float*)x, a, sh, mean, rstd, ws, HW, C, G, rpb, silu);
  } else {
    gn_bwd_partial_kernel<bf16><<<grid, GN_THREADS, 0, st>>>((const bf16*)dy, (const bf16*)x, a, sh, mean, rstd, ws, HW, C, G, rpb, silu);
  }
  gn_bwd_finalize_kernel<<<cdiv((long)B * C, 256), 256, 0, st>>>(ws, gamma, mean, rstd, P, Q, B, C, G, HW);
  if (dgamma && dbeta) gn_bwd_param_kernel<<<cdiv(C, 256), 256, 0, st>>>(ws, dgamma, dbeta, B, C, accumulate);
  if (dtype == SIDLSG_F32)
    gn_bwd_apply_kernel<float, 4><<<cdiv(total / 4, 256), 256, 0, st>>>((const float*)dy, (const float*)x, a, sh, P, Q, (float*)dx, total / 4, HW * C, C, silu);
  else
    gn_bwd_apply_kernel<bf16, 8><<<cdiv(total / 8, 256), 256, 0, st>>>((const bf16*)dy, (const bf16*)x, a, sh, P, Q, (bf16*)dx, total / 8, HW * C, C, silu);
  return check_launch("groupnorm_bwd");
}

extern "C" int sidlsg_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                    float* rstd, long rows, int C, float eps, int dtype, void* stream) {
  if (C > 32 * LN_MAXSLOT_MAX) { set_error("layernorm: C=%d too large", C); return SIDLSG_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((rows + 7) / 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
#define LN_FWD(T, NS) ln_fwd_kernel<T, NS><<<blocks, 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, mean, rstd, rows, C, eps)
  if (dtype == SIDLSG_F32) { if (C <= 320) LN_FWD(float, 10); else if (C <= 640) LN_FWD(float, 20); else LN_FWD(float, 40); }
  else { if (C <= 320) LN_FWD(bf16, 10); else if (C <= 640) LN_FWD(bf16, 20); else LN_FWD(bf16, 40); }
#undef LN_FWD
  return check_launch("layernorm_fwd");
}

// dgamma/dbeta: fp32, ACCUMULATED into (caller zeroes when it wants a fresh gradient); null to skip
extern "C" int sidlsg_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                                    const float* rstd, void* dx, float* dgamma, float* dbeta, long rows, int C,
                                    int dtype, void* stream) {
  if (C > 32 * LN_MAXSLOT_MAX) { set_error("layernorm_bwd: C=%d too large", C); return SIDLSG_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((rows + 7) / 8);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  int want = (dgamma && dbeta) ? 1 : 0;
  size_t sm = sizeof(float) * 2 * C;
#define LN_BWD(T, NS) ln_bwd_kernel<T, NS><<<blocks, 256, sm, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (T*)dx, dgamma, dbeta, rows, C, want)
  if (dtype == SIDLSG_F32) { if (C <= 320) LN_BWD(float, 10); else if (C <= 640) LN_BWD(float, 20); else LN_BWD(float, 40); }
  else { if (C <= 320) LN_BWD(bf16, 10); else if (C <= 640) LN_BWD(bf16, 20); else LN_BWD(bf16, 40); }
#undef LN_BWD
  return check_launch("layernorm_bwd");
}
