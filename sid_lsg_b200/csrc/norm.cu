// GroupNorm(+SiLU) and LayerNorm, forward and backward, on token-major (NHWC) activations.
// HBM-bound kernels: coalesced channel-fastest access, fp32 statistics (double at the combine step).
// GroupNorm semantics: diffusers ResnetBlock2D norm1/norm2 + SiLU, Transformer2DModel.norm (eps 1e-6),
// conv_norm_out.
#include "common.cuh"
#include <stdlib.h>

namespace sidlsg {

constexpr int GN_THREADS = 256;
constexpr int GN_MAXC = 2560;
constexpr int GN_MAXCHUNKS = 64;

template <class T> struct GnVec;   // elements per 16-byte access
template <> struct GnVec<float> { static constexpr int n = 4; };
template <> struct GnVec<bf16> { static constexpr int n = 8; };

template <class T, int V>
__device__ __forceinline__ void gn_load(const T* p, float* out) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < V; ++i) out[i] = to_f(e[i]);
}
template <class T, int V>
__device__ __forceinline__ void gn_store(T* p, const float* in) {
  uint4 raw;
  T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
  for (int i = 0; i < V; ++i) e[i] = from_f<T>(in[i]);
  *reinterpret_cast<uint4*>(p) = raw;
}

// Launch geometry shared by the statistics kernels: blockDim = (VX, RY); a thread owns V consecutive channels of
// channel slab `it` (C is covered by `iters` slabs of VX*V channels) and walks rows r0+y, r0+y+RY, ...
struct GnGeom { int vx, ry, iters, chunks, rows_per_chunk; };

static GnGeom gn_geom(int B, int HW, int C, int V) {
  GnGeom g;
  int vpr = C / V;
  g.iters = (vpr + GN_THREADS - 1) / GN_THREADS;
  g.vx = (vpr + g.iters - 1) / g.iters;
  g.ry = GN_THREADS / g.vx;
  if (g.ry < 1) g.ry = 1;
  if (g.ry > HW) g.ry = HW;
  static int waves = -1;                            // SIDLSG_GN_WAVES (default 8: 2.4 -> 2.6-3.1 TB/s on the 64x64 levels): CTAs per SM the grid aims at
  if (waves < 0) { const char* e = getenv("SIDLSG_GN_WAVES"); waves = e ? atoi(e) : 8; if (waves < 1) waves = 1; }
  int want = (waves * 148 + B - 1) / B;
  int maxc = HW / (g.ry * 2);                       // at least 2 rows per thread
  if (maxc < 1) maxc = 1;
  g.chunks = want < maxc ? want : maxc;
  if (g.chunks > GN_MAXCHUNKS) g.chunks = GN_MAXCHUNKS;
  if (g.chunks < 1) g.chunks = 1;
  g.rows_per_chunk = (HW + g.chunks - 1) / g.chunks;
  g.chunks = (HW + g.rows_per_chunk - 1) / g.rows_per_chunk;
  return g;
}

// ---- forward statistics: partial[b][chunk][c] = (sum x, sum x^2) over the chunk's rows, fp32 ---------------------
template <class T, int V>
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const T* __restrict__ x, float2* __restrict__ partial, int HW, int C, int vpr, int iters,
                int rows_per_chunk) {
  extern __shared__ float red[];   // [2][RY][VX*V]
  const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int tx = threadIdx.x, ty = threadIdx.y, VX = blockDim.x, RY = blockDim.y;
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  const T* xb = x + (long)b * HW * C;
  for (int it = 0; it < iters; ++it) {
    const int v = it * VX + tx;
    float s[V], q[V];
#pragma unroll
    for (int i = 0; i < V; ++i) s[i] = q[i] = 0.f;
    if (v < vpr) {
      for (int r = r0 + ty; r < r1; r += RY) {
        float e[V];
        gn_load<T, V>(xb + (long)r * C + v * V, e);
#pragma unroll
        for (int i = 0; i < V; ++i) { s[i] += e[i]; q[i] = fmaf(e[i], e[i], q[i]); }
      }
    }
    const int W = VX * V;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) { red[ty * W + tx * V + i] = s[i]; red[(RY + ty) * W + tx * V + i] = q[i]; }
    __syncthreads();
    for (int c = ty * VX + tx; c < W; c += VX * RY) {
      float ss = 0.f, qq = 0.f;
      for (int y = 0; y < RY; ++y) { ss += red[y * W + c]; qq += red[(RY + y) * W + c]; }
      const int ch = it * W + c;
      if (ch < C) partial[((long)b * chunks + chunk) * C + ch] = make_float2(ss, qq);
    }
  }
}

// one warp per (b, group): combine chunks x channels in double -> mean / rstd, then the per-channel affine
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float2* __restrict__ partial, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ mean_out, float* __restrict__ rstd_out, float* __restrict__ a,
                   float* __restrict__ sh, int B, int C, int G, int HW, int chunks, float eps) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * G) return;
  const int b = w / G, g = w - b * G, cpg = C / G;
  // a lane owns whole chunks (its cpg channels are contiguous: no index division, loads issued back to back with two
  // independent accumulator pairs) - the previous element-strided loop was a chain of 20 dependent L2 round trips and
  // made this 10-25 us of pure latency per GroupNorm call
  double s = 0, q = 0;
  for (int ch = lane; ch < chunks; ch += 32) {
    const float2* pp = partial + ((long)b * chunks + ch) * C + g * cpg;
    double s0 = 0, q0 = 0, s1 = 0, q1 = 0;
    int j = 0;
#pragma unroll 4
    for (; j + 1 < cpg; j += 2) {
      const float2 v0 = pp[j], v1 = pp[j + 1];
      s0 += v0.x; q0 += v0.y; s1 += v1.x; q1 += v1.y;
    }
    if (j < cpg) { const float2 v0 = pp[j]; s0 += v0.x; q0 += v0.y; }
    s += s0 + s1; q += q0 + q1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  const double n = (double)HW * cpg;
  const double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float mf = (float)mean;
  if (lane == 0) { mean_out[w] = mf; rstd_out[w] = rstd; }
  for (int j = lane; j < cpg; j += 32) {
    const int c = g * cpg + j;
    const float ga = gamma[c];
    a[(long)b * C + c] = rstd * ga;
    sh[(long)b * C + c] = beta[c] - mf * rstd * ga;
  }
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

// y = [silu](x * a[b,c] + sh[b,c]) with the statistics kernels' geometry: a thread owns V consecutive channels (its
// coefficients live in registers) and walks the rows of its chunk, so the loop body is one 16-byte load, V FMAs and
// one 16-byte store; consecutive threads touch consecutive 16-byte pieces of the row-major tensor.
template <class T, int V>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_rows_kernel(const T* __restrict__ x, const float* __restrict__ a, const float* __restrict__ sh,
                     T* __restrict__ y, int HW, int C, int vpr, int iters, int rows_per_chunk, int silu) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tx = threadIdx.x, ty = threadIdx.y, VX = blockDim.x, RY = blockDim.y;
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  const long base = (long)b * HW * C;
  for (int it = 0; it < iters; ++it) {
    const int v = it * VX + tx;
    if (v >= vpr) continue;
    float av[V], sv[V];
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      const float4 a4 = *reinterpret_cast<const float4*>(a + (long)b * C + v * V + i);
      const float4 s4 = *reinterpret_cast<const float4*>(sh + (long)b * C + v * V + i);
      av[i] = a4.x; av[i + 1] = a4.y; av[i + 2] = a4.z; av[i + 3] = a4.w;
      sv[i] = s4.x; sv[i + 1] = s4.y; sv[i + 2] = s4.z; sv[i + 3] = s4.w;
    }
#pragma unroll 2
    for (int r = r0 + ty; r < r1; r += RY) {
      float e[V];
      gn_load<T, V>(x + base + (long)r * C + v * V, e);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float t = fmaf(e[i], av[i], sv[i]);
        e[i] = silu ? silu_f(t) : t;
      }
      gn_store<T, V>(y + base + (long)r * C + v * V, e);
    }
  }
}

// ---- backward ---------------------------------------------------------------------------------------------------
// partial[b][chunk][c] = (sum dz, sum dz*xhat); dz = dy * silu'(x*a+sh) (or dy)
template <class T, int V>
__global__ void __launch_bounds__(GN_THREADS)
gn_bwd_stats_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ a,
                    const float* __restrict__ sh, const float* __restrict__ mean, const float* __restrict__ rstd,
                    float2* __restrict__ partial, int HW, int C, int G, int vpr, int iters, int rows_per_chunk, int silu) {
  extern __shared__ float red[];
  const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int tx = threadIdx.x, ty = threadIdx.y, VX = blockDim.x, RY = blockDim.y;
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  const int cpg = C / G;
  const long base = (long)b * HW * C;
  for (int it = 0; it < iters; ++it) {
    const int v = it * VX + tx;
    float s[V], q[V], av[V], sv[V], mu[V], rs[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s[i] = q[i] = 0.f;
      const int c = v * V + i;
      if (v < vpr) {
        av[i] = a[(long)b * C + c]; sv[i] = sh[(long)b * C + c];
        mu[i] = mean[b * G + c / cpg]; rs[i] = rstd[b * G + c / cpg];
      } else { av[i] = sv[i] = mu[i] = rs[i] = 0.f; }
    }
    if (v < vpr) {
      for (int r = r0 + ty; r < r1; r += RY) {
        float xe[V], de[V];
        gn_load<T, V>(x + base + (long)r * C + v * V, xe);
        gn_load<T, V>(dy + base + (long)r * C + v * V, de);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float dz = de[i];
          if (silu) dz *= silu_grad_f(fmaf(xe[i], av[i], sv[i]));
          s[i] += dz;
          q[i] = fmaf(dz, (xe[i] - mu[i]) * rs[i], q[i]);
        }
      }
    }
    const int W = VX * V;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) { red[ty * W + tx * V + i] = s[i]; red[(RY + ty) * W + tx * V + i] = q[i]; }
    __syncthreads();
    for (int c = ty * VX + tx; c < W; c += VX * RY) {
      float ss = 0.f, qq = 0.f;
      for (int y = 0; y < RY; ++y) { ss += red[y * W + c]; qq += red[(RY + y) * W + c]; }
      const int ch = it * W + c;
      if (ch < C) partial[((long)b * chunks + chunk) * C + ch] = make_float2(ss, qq);
    }
  }
}

// one warp per (b, group): coefficients of dx = A*dz + x*P + Q
__global__ void __launch_bounds__(256)
gn_bwd_finalize_kernel(const float2* __restrict__ partial, const float* __restrict__ gamma,
                       const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ P,
                       float* __restrict__ Q, int B, int C, int G, int HW, int chunks) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * G) return;
  const int b = w / G, g = w - b * G, cpg = C / G;
  double S1 = 0, S2 = 0;
  for (int ch = lane; ch < chunks; ch += 32) {       // lane = chunk, as in gn_finalize_kernel
    const float2* pp = partial + ((long)b * chunks + ch) * C + g * cpg;
    const float* gp = gamma + g * cpg;
    double a0 = 0, b0 = 0, a1 = 0, b1 = 0;
    int j = 0;
#pragma unroll 4
    for (; j + 1 < cpg; j += 2) {
      const float2 v0 = pp[j], v1 = pp[j + 1];
      const double g0 = gp[j], g1 = gp[j + 1];
      a0 += g0 * v0.x; b0 += g0 * v0.y; a1 += g1 * v1.x; b1 += g1 * v1.y;
    }
    if (j < cpg) { const float2 v0 = pp[j]; const double g0 = gp[j]; a0 += g0 * v0.x; b0 += g0 * v0.y; }
    S1 += a0 + a1; S2 += b0 + b1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { S1 += __shfl_xor_sync(0xffffffffu, S1, o); S2 += __shfl_xor_sync(0xffffffffu, S2, o); }
  const double n = (double)HW * cpg;
  const double m = mean[w], r = rstd[w];
  const float Pv = (float)(-r * r * S2 / n), Qv = (float)(-r * S1 / n + m * r * r * S2 / n);
  for (int j = lane; j < cpg; j += 32) {
    P[(long)b * C + g * cpg + j] = Pv;
    Q[(long)b * C + g * cpg + j] = Qv;
  }
}

// dgamma[c] += sum_{b,chunk} partial.y ; dbeta[c] += sum partial.x.  blockIdx.y strides over the (b, chunk) rows so
// the reduction is spread over SMs; slices meet through fp32 atomics (the caller zeroes dgamma/dbeta first when
// it does not accumulate).
__global__ void gn_bwd_param_kernel(const float2* __restrict__ partial, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int rows, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f, q = 0.f;
  for (int i = blockIdx.y; i < rows; i += gridDim.y) {
    const float2 v = partial[(long)i * C + c];
    s += v.x; q += v.y;
  }
  atomicAdd(dgamma + c, q);
  atomicAdd(dbeta + c, s);
}

// dx = a * dz + x * P + Q, dz = dy * silu'(x*a+sh) (or dy); same row-walking geometry as gn_apply_rows_kernel
template <class T, int V>
__global__ void __launch_bounds__(GN_THREADS)
gn_bwd_apply_rows_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ a,
                         const float* __restrict__ sh, const float* __restrict__ P, const float* __restrict__ Q,
                         T* __restrict__ dx, int HW, int C, int vpr, int iters, int rows_per_chunk, int silu) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tx = threadIdx.x, ty = threadIdx.y, VX = blockDim.x, RY = blockDim.y;
  const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
  const long base = (long)b * HW * C;
  for (int it = 0; it < iters; ++it) {
    const int v = it * VX + tx;
    if (v >= vpr) continue;
    float av[V], sv[V], pv[V], qv[V];
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      const long o = (long)b * C + v * V + i;
      const float4 a4 = *reinterpret_cast<const float4*>(a + o), s4 = *reinterpret_cast<const float4*>(sh + o);
      const float4 p4 = *reinterpret_cast<const float4*>(P + o), q4 = *reinterpret_cast<const float4*>(Q + o);
      av[i] = a4.x; av[i + 1] = a4.y; av[i + 2] = a4.z; av[i + 3] = a4.w;
      sv[i] = s4.x; sv[i + 1] = s4.y; sv[i + 2] = s4.z; sv[i + 3] = s4.w;
      pv[i] = p4.x; pv[i + 1] = p4.y; pv[i + 2] = p4.z; pv[i + 3] = p4.w;
      qv[i] = q4.x; qv[i + 1] = q4.y; qv[i + 2] = q4.z; qv[i + 3] = q4.w;
    }
#pragma unroll 2
    for (int r = r0 + ty; r < r1; r += RY) {
      float xe[V], de[V];
      gn_load<T, V>(x + base + (long)r * C + v * V, xe);
      gn_load<T, V>(dy + base + (long)r * C + v * V, de);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float dz = de[i];
        if (silu) dz *= silu_grad_f(fmaf(xe[i], av[i], sv[i]));
        de[i] = fmaf(av[i], dz, fmaf(xe[i], pv[i], qv[i]));
      }
      gn_store<T, V>(dx + base + (long)r * C + v * V, de);
    }
  }
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

// ---- LayerNorm, vectorised (C % V == 0): lane l owns the 16-byte pieces l, l+32, ... of a row, so a warp reads a row
// in 512-byte requests instead of 64-byte ones and gamma / beta live in registers across the warp's rows -----------------
template <class T, int NV>
__global__ void __launch_bounds__(256)
ln_fwd_vec_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                  T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, long rows, int C,
                  float eps) {
  constexpr int V = GnVec<T>::n;
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  const int pieces = C / V;
  float gm[NV][V], bt[NV][V];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int p = lane + 32 * i;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      gm[i][j] = p < pieces ? gamma[p * V + j] : 0.f;
      bt[i][j] = p < pieces ? beta[p * V + j] : 0.f;
    }
  }
  const float invC = 1.f / C;
  for (long r = warp; r < rows; r += nwarps) {
    const T* row = x + r * C;
    float v[NV][V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int p = lane + 32 * i;
      if (p < pieces) gn_load<T, V>(row + p * V, v[i]);
      else {
#pragma unroll
        for (int j = 0; j < V; ++j) v[i][j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < V; ++j) s += v[i][j];
    }
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < pieces) {
#pragma unroll
        for (int j = 0; j < V; ++j) { const float dd = v[i][j] - mean; q = fmaf(dd, dd, q); }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invC + eps);
    if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
    T* yr = y + r * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int p = lane + 32 * i;
      if (p < pieces) {
        float o[V];
#pragma unroll
        for (int j = 0; j < V; ++j) o[j] = fmaf((v[i][j] - mean) * rstd, gm[i][j], bt[i][j]);
        gn_store<T, V>(yr + p * V, o);
      }
    }
  }
}

template <class T, int NV>
__global__ void __launch_bounds__(256)
ln_bwd_vec_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma,
                  const float* __restrict__ mean, const float* __restrict__ rstd, T* __restrict__ dx,
                  float* __restrict__ dgamma, float* __restrict__ dbeta, long rows, int C, int want_param_grad) {
  extern __shared__ float red[];  // [2][C]
  constexpr int V = GnVec<T>::n;
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  const int pieces = C / V;
  float gm[NV][V], dg[NV][V], db[NV][V];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int p = lane + 32 * i;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      dg[i][j] = db[i][j] = 0.f;
      gm[i][j] = p < pieces ? gamma[p * V + j] : 0.f;
    }
  }
  const float invC = 1.f / C;
  for (long r = warp; r < rows; r += nwarps) {
    const float mu = mean[r], rs = rstd[r];
    float xh[NV][V], g[NV][V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int p = lane + 32 * i;
      if (p < pieces) {
        float de[V];
        gn_load<T, V>(dy + r * C + p * V, de);
        gn_load<T, V>(x + r * C + p * V, xh[i]);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          xh[i][j] = (xh[i][j] - mu) * rs;
          g[i][j] = de[j] * gm[i][j];
          dg[i][j] = fmaf(de[j], xh[i][j], dg[i][j]);
          db[i][j] += de[j];
          s1 += g[i][j];
          s2 = fmaf(g[i][j], xh[i][j], s2);
        }
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) xh[i][j] = g[i][j] = 0.f;
      }
    }
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int p = lane + 32 * i;
      if (p < pieces) {
        float o[V];
#pragma unroll
        for (int j = 0; j < V; ++j) o[j] = rs * (g[i][j] - s1 - xh[i][j] * s2);
        gn_store<T, V>(dx + r * C + p * V, o);
      }
    }
  }
  if (!want_param_grad) return;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int p = lane + 32 * i;
    if (p < pieces) {
#pragma unroll
      for (int j = 0; j < V; ++j) { atomicAdd(&red[p * V + j], dg[i][j]); atomicAdd(&red[C + p * V + j], db[i][j]); }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&dgamma[c], red[c]);
    atomicAdd(&dbeta[c], red[C + c]);
  }
}

}  // namespace sidlsg

using namespace sidlsg;

// scratch bytes the GroupNorm entry points need for `ws` (per-chunk fp32 partial sums)
extern "C" long sidlsg_groupnorm_ws_bytes(int B, int HW, int C, int dtype) {
  if (B <= 0 || HW <= 0 || C <= 0) return 16;
  GnGeom g = gn_geom(B, HW, C, dtype == SIDLSG_F32 ? 4 : 8);
  return (long)sizeof(float2) * B * g.chunks * C;
}

template <class T>
static void gn_launch_stats(const T* x, float2* ws, int B, int HW, int C, cudaStream_t st) {
  constexpr int V = GnVec<T>::n;
  GnGeom g = gn_geom(B, HW, C, V);
  size_t sm = sizeof(float) * 2 * g.ry * g.vx * V;
  gn_stats_kernel<T, V><<<dim3(g.chunks, B), dim3(g.vx, g.ry), sm, st>>>(x, ws, HW, C, C / V, g.iters, g.rows_per_chunk);
}

// ws: sidlsg_groupnorm_ws_bytes(B,HW,C,in_dtype) bytes.  Outputs mean/rstd [B,G] and the per-(b,c) affine a/sh [B,C].
extern "C" int sidlsg_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                    float* rstd, float* a, float* sh, void* ws, int B, int HW, int C, int G,
                                    float eps, int silu, int in_dtype, int out_dtype, void* stream) {
  if (C > GN_MAXC || C % G || C % 8) { set_error("groupnorm: unsupported C=%d G=%d", C, G); return SIDLSG_ERR_ARG; }
  if (B == 0 || HW == 0) return SIDLSG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float2* part = (float2*)ws;
  GnGeom g = gn_geom(B, HW, C, in_dtype == SIDLSG_F32 ? 4 : 8);
  if (in_dtype == SIDLSG_F32) gn_launch_stats<float>((const float*)x, part, B, HW, C, st);
  else gn_launch_stats<bf16>((const bf16*)x, part, B, HW, C, st);
  gn_finalize_kernel<<<cdiv((long)B * G * 32, 256), 256, 0, st>>>(part, gamma, beta, mean, rstd, a, sh, B, C, G, HW, g.chunks, eps);
  long total = (long)B * HW * C;
  const dim3 agrid(g.chunks, B), ablock(g.vx, g.ry);
  if (in_dtype == SIDLSG_F32 && out_dtype == SIDLSG_F32)
    gn_apply_rows_kernel<float, 4><<<agrid, ablock, 0, st>>>((const float*)x, a, sh, (float*)y, HW, C, C / 4, g.iters, g.rows_per_chunk, silu);
  else if (in_dtype == SIDLSG_BF16 && out_dtype == SIDLSG_BF16)
    gn_apply_rows_kernel<bf16, 8><<<agrid, ablock, 0, st>>>((const bf16*)x, a, sh, (bf16*)y, HW, C, C / 8, g.iters, g.rows_per_chunk, silu);
  else if (in_dtype == SIDLSG_F32 && out_dtype == SIDLSG_BF16)
    gn_apply_kernel<float, bf16, 4><<<cdiv(total / 4, 256), 256, 0, st>>>((const float*)x, a, sh, (bf16*)y, total / 4, HW * C, C, silu);
  else { set_error("groupnorm: unsupported dtype pair"); return SIDLSG_ERR_UNSUPPORTED; }
  return check_launch("groupnorm_fwd");
}

// ws as above; P,Q: float[B*C] scratch.  dgamma/dbeta fp32 (accumulate flag), may be null.
extern "C" int sidlsg_groupnorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                                    const float* rstd, const float* a, const float* sh, void* dx, float* dgamma,
                                    float* dbeta, void* ws, float* P, float* Q, int B, int HW, int C, int G,
                                    int silu, int accumulate, int dtype, void* stream) {
  if (C > GN_MAXC || C % G || C % 8) { set_error("groupnorm_bwd: unsupported C=%d G=%d", C, G); return SIDLSG_ERR_ARG; }
  if (B == 0 || HW == 0) return SIDLSG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float2* part = (float2*)ws;
  long total = (long)B * HW * C;
  const int V = dtype == SIDLSG_F32 ? 4 : 8;
  GnGeom g = gn_geom(B, HW, C, V);
  size_t sm = sizeof(float) * 2 * g.ry * g.vx * V;
  dim3 grid(g.chunks, B), block(g.vx, g.ry);
  if (dtype == SIDLSG_F32)
    gn_bwd_stats_kernel<float, 4><<<grid, block, sm, st>>>((const float*)dy, (const float*)x, a, sh, mean, rstd, part, HW, C, G, C / 4, g.iters, g.rows_per_chunk, silu);
  else
    gn_bwd_stats_kernel<bf16, 8><<<grid, block, sm, st>>>((const bf16*)dy, (const bf16*)x, a, sh, mean, rstd, part, HW, C, G, C / 8, g.iters, g.rows_per_chunk, silu);
  gn_bwd_finalize_kernel<<<cdiv((long)B * G * 32, 256), 256, 0, st>>>(part, gamma, mean, rstd, P, Q, B, C, G, HW, g.chunks);
  if (dgamma && dbeta) {
    if (!accumulate) {
      cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st);
      cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st);
    }
    const int rows = B * g.chunks;
    const int ysplit = rows < 64 ? rows : 64;
    gn_bwd_param_kernel<<<dim3(cdiv(C, 128), ysplit), 128, 0, st>>>(part, dgamma, dbeta, rows, C);
  }
  (void)total;
  if (dtype == SIDLSG_F32)
    gn_bwd_apply_rows_kernel<float, 4><<<grid, block, 0, st>>>((const float*)dy, (const float*)x, a, sh, P, Q, (float*)dx, HW, C, C / 4, g.iters, g.rows_per_chunk, silu);
  else
    gn_bwd_apply_rows_kernel<bf16, 8><<<grid, block, 0, st>>>((const bf16*)dy, (const bf16*)x, a, sh, P, Q, (bf16*)dx, HW, C, C / 8, g.iters, g.rows_per_chunk, silu);
  return check_launch("groupnorm_bwd");
}

extern "C" int sidlsg_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                    float* rstd, long rows, int C, float eps, int dtype, void* stream) {
  if (C > 32 * LN_MAXSLOT_MAX) { set_error("layernorm: C=%d too large", C); return SIDLSG_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((rows + 7) / 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  {
    // vectorised kernels: slots = ceil(C / V / 32) in {1, 2, 3, 5, 10}
    const int V = dtype == SIDLSG_F32 ? 4 : 8;
    const int nv = (C % V) ? 0 : (C / V + 31) / 32;
#define LN_FWD_V(T, NV) { ln_fwd_vec_kernel<T, NV><<<blocks, 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, mean, rstd, rows, C, eps); return check_launch("layernorm_fwd"); }
    if (nv >= 1 && nv <= 10 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
      if (dtype == SIDLSG_F32) {
        if (nv <= 1) LN_FWD_V(float, 1) else if (nv <= 2) LN_FWD_V(float, 2) else if (nv <= 3) LN_FWD_V(float, 3)
        else if (nv <= 5) LN_FWD_V(float, 5) else LN_FWD_V(float, 10)
      } else {
        if (nv <= 1) LN_FWD_V(bf16, 1) else if (nv <= 2) LN_FWD_V(bf16, 2) else if (nv <= 3) LN_FWD_V(bf16, 3)
        else if (nv <= 5) LN_FWD_V(bf16, 5)
      }
    }
#undef LN_FWD_V
  }
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
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  int want = (dgamma && dbeta) ? 1 : 0;
  size_t sm = sizeof(float) * 2 * C;
  {
    const int V = dtype == SIDLSG_F32 ? 4 : 8;
    const int nv = (C % V) ? 0 : (C / V + 31) / 32;
#define LN_BWD_V(T, NV) { ln_bwd_vec_kernel<T, NV><<<blocks, 256, sm, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (T*)dx, dgamma, dbeta, rows, C, want); return check_launch("layernorm_bwd"); }
    if (nv >= 1 && nv <= 10 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(dx) & 15) == 0) {
      if (dtype == SIDLSG_F32) {
        if (nv <= 1) LN_BWD_V(float, 1) else if (nv <= 2) LN_BWD_V(float, 2) else if (nv <= 3) LN_BWD_V(float, 3)
        else if (nv <= 5) LN_BWD_V(float, 5) else LN_BWD_V(float, 10)
      } else {
        if (nv <= 1) LN_BWD_V(bf16, 1) else if (nv <= 2) LN_BWD_V(bf16, 2) else if (nv <= 3) LN_BWD_V(bf16, 3)
        else if (nv <= 5) LN_BWD_V(bf16, 5)
      }
    }
#undef LN_BWD_V
  }
#define LN_BWD(T, NS) ln_bwd_kernel<T, NS><<<blocks, 256, sm, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (T*)dx, dgamma, dbeta, rows, C, want)
  if (dtype == SIDLSG_F32) { if (C <= 320) LN_BWD(float, 10); else if (C <= 640) LN_BWD(float, 20); else LN_BWD(float, 40); }
  else { if (C <= 320) LN_BWD(bf16, 10); else if (C <= 640) LN_BWD(bf16, 20); else LN_BWD(bf16, 40); }
#undef LN_BWD
  return check_launch("layernorm_bwd");
}
