// CUDA-core (FFMA) tiled GEMM / implicit-GEMM conv3x3 with fp32 accumulation.
//
// This is the fp32-exact path of the hot path (the reference's fp32 mode runs with TF32 off, i.e.
// without tensor cores: /root/reference/training/sid_training_loop.py:241-243) and the catch-all for
// GEMM shapes the tcgen05 kernels do not take (tiny M, K=36 conv_in, N=4 conv_out).  One tile core,
// pluggable operand loaders:
//   Dense       A(m,k)/B(k,n) with arbitrary strides and two batch levels (linear layers, Q.K^T, P.V, dgrad, wgrad)
//   ConvGather  NHWC 3x3 window gather with stride / fused nearest-2x upsample / transposed (stride-2 dgrad)
//   ConvWeight  [n][tap][k] weights with optional tap flip (dgrad)
#include "common.cuh"

namespace sidlsg {

struct ConvGeom {
  int B, Hi, Wi, Cin;   // tensor being gathered from ("input" of this pass)
  int Ho, Wo;           // pixel domain being produced
  int stride, up, transposed;
};

template <class T>
struct Dense {
  const T* p;
  long s_row, s_col, s_b1, s_b2;
  int rows, cols, nb2;
  __device__ __forceinline__ void set_batch(int z) { p += (long)(z / nb2) * s_b1 + (long)(z % nb2) * s_b2; }
  __device__ __forceinline__ float get(int r, int c) const {
    return (r < rows && c < cols) ? to_f(p[(long)r * s_row + (long)c * s_col]) : 0.f;
  }
};

// element (pix, kk): pix over [B,Ho,Wo], kk = tap*Cin + ci
template <class T, bool PIX_IS_ROW>
struct ConvGather {
  const T* x;
  ConvGeom g;
  int npix, kk_total;
  __device__ __forceinline__ void set_batch(int) {}
  __device__ __forceinline__ float get(int r, int c) const {
    int pix = PIX_IS_ROW ? r : c;
    int kk = PIX_IS_ROW ? c : r;
    if (pix >= npix || kk >= kk_total) return 0.f;
    int tap = kk / g.Cin, ci = kk - tap * g.Cin;
    int dy = tap / 3, dx = tap - dy * 3;
    int hw = g.Ho * g.Wo;
    int b = pix / hw, rem = pix - b * hw;
    int oy = rem / g.Wo, ox = rem - oy * g.Wo;
    int iy, ix;
    if (!g.transposed) {
      int uy = oy * g.stride + dy - 1, ux = ox * g.stride + dx - 1;
      if (uy < 0 || ux < 0 || uy >= g.Hi * g.up || ux >= g.Wi * g.up) return 0.f;
      iy = uy / g.up; ix = ux / g.up;
    } else {
      int ty = oy + dy - 1, tx = ox + dx - 1;
      if (ty < 0 || tx < 0 || (ty % g.stride) || (tx % g.stride)) return 0.f;
      iy = ty / g.stride; ix = tx / g.stride;
      if (iy >= g.Hi || ix >= g.Wi) return 0.f;
    }
    return to_f(x[(((long)b * g.Hi + iy) * g.Wi + ix) * g.Cin + ci]);
  }
};

// B(k,n) with k = tap*Kc + kc : w[n*s_n + tap'*s_tap + kc*s_k]
template <class T>
struct ConvWeight {
  const T* w;
  long s_n, s_tap, s_k;
  int Kc, N, flip;
  __device__ __forceinline__ void set_batch(int) {}
  __device__ __forceinline__ float get(int k, int n) const {
    if (n >= N || k >= 9 * Kc) return 0.f;
    int tap = k / Kc, kc = k - tap * Kc;
    if (flip) tap = 8 - tap;
    return to_f(w[(long)n * s_n + (long)tap * s_tap + (long)kc * s_k]);
  }
};

template <class TO>
struct Epilogue {
  TO* c;
  long ldc, c_b1, c_b2;
  const float* bias;      // [N] or null
  const TO* res;          // residual, same indexing family as c (own strides) or null
  long ldr, r_b1, r_b2;
  const float* rowvec;    // [M / rows_per_vec][N] broadcast add (timestep-embedding projection) or null
  int rows_per_vec;
  float alpha;
  int accumulate;         // 0 store, 1 c += v, 2 atomicAdd (fp32 only; split-K)
  int M, N, nb2;
  // generic [n][tap][k]-strided output for conv wgrad: n index -> (tap, ci)
  int wg_cin;             // 0 = plain; else column n = tap*wg_cin+ci and row m = co
  long wg_sco, wg_stap, wg_sci;
  __device__ __forceinline__ void store(int z, int m, int n, float v, bool first_split) const {
    if (m >= M || n >= N) return;
    v *= alpha;
    long zo1 = z / nb2, zo2 = z % nb2;
    if (first_split) {
      if (bias) v += bias[n];
      if (rowvec) v += rowvec[(long)(m / rows_per_vec) * N + n];
      if (res) v += to_f(res[zo1 * r_b1 + zo2 * r_b2 + (long)m * ldr + n]);
    }
    TO* dst;
    if (wg_cin) {
      int tap = n / wg_cin, ci = n - tap * wg_cin;
      dst = c + (long)m * wg_sco + (long)tap * wg_stap + (long)ci * wg_sci;
    } else {
      dst = c + zo1 * c_b1 + zo2 * c_b2 + (long)m * ldc + n;
    }
    if (accumulate == 0) *dst = from_f<TO>(v);
    else if (accumulate == 1) *dst = from_f<TO>(to_f(*dst) + v);
    else atomic_add(dst, v);
  }
  __device__ __forceinline__ static void atomic_add(float* p, float v) { atomicAdd(p, v); }
  __device__ __forceinline__ static void atomic_add(bf16* p, float v) { atomicAdd(p, __float2bfloat16_rn(v)); }
};

// C tile BMxBN, K step BK, 256 threads, each thread TMxTN (strided by 16 so smem reads are conflict-free)
template <int BM, int BN, int BK, int TM, int TN, class LA, class LB, class TO>
__global__ void __launch_bounds__(256)
gemm_core(LA la, LB lb, Epilogue<TO> epi, int M, int N, int K, int ksplit, int a_kcontig, int b_kcontig) {
  static_assert(BM == 16 * TM && BN == 16 * TN, "16x16 thread grid");
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int z = blockIdx.z / ksplit, split = blockIdx.z - z * ksplit;
  la.set_batch(z);
  lb.set_batch(z);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ktiles = (K + BK - 1) / BK;
  const int kt_per = (ktiles + ksplit - 1) / ksplit;
  const int kt_begin = split * kt_per;
  const int kt_end = min(ktiles, kt_begin + kt_per);

  constexpr int A_PER = BM * BK / 256, B_PER = BN * BK / 256;
  float ra[A_PER], rb[B_PER];
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto load_tile = [&](int kt) {
    const int k0 = kt * BK;
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int idx = tid + i * 256;
      int mm = a_kcontig ? idx / BK : idx % BM;
      int kk = a_kcontig ? idx % BK : idx / BM;
      ra[i] = la.get(m0 + mm, k0 + kk);
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int idx = tid + i * 256;
      int nn = b_kcontig ? idx / BK : idx % BN;
      int kk = b_kcontig ? idx % BK : idx / BN;
      rb[i] = lb.get(k0 + kk, n0 + nn);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int idx = tid + i * 256;
      int mm = a_kcontig ? idx / BK : idx % BM;
      int kk = a_kcontig ? idx % BK : idx / BM;
      As[buf][kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int idx = tid + i * 256;
      int nn = b_kcontig ? idx / BK : idx % BN;
      int kk = b_kcontig ? idx % BK : idx / BN;
      Bs[buf][kk][nn] = rb[i];
    }
  };

  if (kt_begin < kt_end) {
    load_tile(kt_begin);
    stash(0);
    __syncthreads();
    for (int kt = kt_begin; kt < kt_end; ++kt) {
      const int buf = (kt - kt_begin) & 1;
      if (kt + 1 < kt_end) load_tile(kt + 1);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float av[TM], bv[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) av[i] = As[buf][kk][ty + 16 * i];
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] = Bs[buf][kk][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (kt + 1 < kt_end) {
        stash(buf ^ 1);
        __syncthreads();
      }
    }
  } else if (split != 0) {
    return;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) epi.store(z, m0 + ty + 16 * i, n0 + tx + 16 * j, acc[i][j], split == 0);
}

long g_simt_launches = 0;
extern thread_local int g_last_path;

template <class LA, class LB, class TO>
static int launch_core(LA la, LB lb, Epilogue<TO> epi, int M, int N, int K, int nbatch, int ksplit,
                       int a_kc, int b_kc, cudaStream_t st) {
  if (M <= 0 || N <= 0 || nbatch <= 0) return SIDLSG_OK;
  __atomic_add_fetch(&g_simt_launches, 1, __ATOMIC_RELAXED);
  g_last_path = 0;
  bool big = (long)M * N >= 128L * 128 * 64 && M >= 128 && N >= 96;
  if (big) {
    dim3 grid(cdiv(N, 128), cdiv(M, 128), nbatch * ksplit);
    gemm_core<128, 128, 8, 8, 8, LA, LB, TO><<<grid, 256, 0, st>>>(la, lb, epi, M, N, K, ksplit, a_kc, b_kc);
  } else {
    dim3 grid(cdiv(N, 64), cdiv(M, 64), nbatch * ksplit);
    gemm_core<64, 64, 16, 4, 4, LA, LB, TO><<<grid, 256, 0, st>>>(la, lb, epi, M, N, K, ksplit, a_kc, b_kc);
  }
  return check_launch("gemm_core");
}

static int pick_ksplit(int M, int N, int K, int nbatch) {
  long tiles = (long)cdiv(M, 128) * cdiv(N, 128) * nbatch;
  if (tiles >= 148 || K < 2048) return 1;
  long s = (2 * 148 + tiles - 1) / tiles;
  long maxs = K / 512;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return (int)s;
}

// ---- skinny fp32 GEMMs (timestep-embedding path: M = batch <= 32 rows, or a reduction over <= 32 rows) ----------
// weight-bandwidth bound: the weight matrix is streamed exactly once with coalesced / 128-bit accesses.
constexpr int SK_M = 32;

// C[m][n] = sum_k A[m][k] W[n][k]   (W K-contiguous, K % 4 == 0): one warp per SK_NW output columns, lanes split K in
// float4s, so every A quad fetched (L1-resident, shared by all warps) feeds 4 * SK_NW FMAs; blockIdx.y = K slab (slabs
// meet through fp32 atomics when gridDim.y > 1: the caller zeroes C first unless it accumulates).
// The cross-lane reduction is a butterfly that halves the live values at every step (31 shuffles per column instead
// of 32 x 5) and leaves row m's total in lane m, so the whole kernel is ~1.5k instructions: the fully unrolled
// predicated version it replaces was 38k instructions (600 KB of code) and spent ~100 us per launch missing the
// instruction cache.
constexpr int SK_NW = 4;
__global__ void __launch_bounds__(256)
skinny_nt_kernel(const float* __restrict__ A, long a_sm, const float* __restrict__ W, long w_sn, float* __restrict__ C,
                 long ldc, const float* __restrict__ bias, const float* __restrict__ res, long ldr, int M, int N, int K,
                 int kslab, float alpha, int accumulate) {
  const int n0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * SK_NW;
  const int lane = threadIdx.x & 31;
  if (n0 >= N) return;
  const int k0 = blockIdx.y * kslab, k1 = min(K, k0 + kslab);
  float acc[SK_NW][SK_M];
#pragma unroll
  for (int j = 0; j < SK_NW; ++j)
#pragma unroll
    for (int m = 0; m < SK_M; ++m) acc[j][m] = 0.f;
  const float* wrow[SK_NW];
#pragma unroll
  for (int j = 0; j < SK_NW; ++j) wrow[j] = W + (long)min(n0 + j, N - 1) * w_sn;
  for (int k = k0 + lane * 4; k < k1; k += 128) {
    float4 wv[SK_NW];
#pragma unroll
    for (int j = 0; j < SK_NW; ++j) wv[j] = *reinterpret_cast<const float4*>(wrow[j] + k);
#pragma unroll
    for (int m = 0; m < SK_M; ++m) {
      // rows >= M re-read row M-1 (their sums are never stored): no predicates in the unrolled body
      const float4 av = *reinterpret_cast<const float4*>(A + (long)min(m, M - 1) * a_sm + k);
#pragma unroll
      for (int j = 0; j < SK_NW; ++j)
        acc[j][m] = fmaf(av.x, wv[j].x, fmaf(av.y, wv[j].y, fmaf(av.z, wv[j].z, fmaf(av.w, wv[j].w, acc[j][m]))));
    }
  }
  const bool first = blockIdx.y == 0;
  const bool atomic = gridDim.y > 1;
#pragma unroll
  for (int j = 0; j < SK_NW; ++j) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float send = up ? acc[j][i] : acc[j][i + off];
        const float keep = up ? acc[j][i + off] : acc[j][i];
        acc[j][i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    // acc[j][0] of lane m = sum over lanes of the original acc[j][m]
    const int n = n0 + j;
    if (lane < M && n < N) {
      float v = acc[j][0] * alpha;
      if (first) {
        if (bias) v += bias[n];
        if (res) v += res[(long)lane * ldr + n];
      }
      float* dst = C + (long)lane * ldc + n;
      if (atomic) atomicAdd(dst, v);
      else *dst = accumulate ? *dst + v : v;
    }
  }
}

// C[m][n] (+)= sum_k A[m][k] B[k][n]   (B n-contiguous): one thread per column n, blockIdx.y = a slab of SK_KSLAB
// reduction rows (every k row of B is read once, coalesced); slabs meet in C through fp32 atomics, so the
// whole matrix is spread over N/128 x K/SK_KSLAB CTAs instead of N/128 (the data-gradient of the timestep
// projections has N = 1280: ten CTAs streamed a 6.5 MB weight matrix before).  C must be zeroed by the caller
// when it is not an accumulation (skinny_try does).
constexpr int SK_KSLAB = 32;
__global__ void __launch_bounds__(128)
skinny_nn_kernel(const float* __restrict__ A, long a_sm, long a_sk, const float* __restrict__ Bm, long b_sk,
                 float* __restrict__ C, long ldc, int M, int N, int K, float alpha) {
  __shared__ float As[SK_KSLAB][SK_M];
  const int n = blockIdx.x * 128 + threadIdx.x;
  const int k0 = blockIdx.y * SK_KSLAB, k1 = min(K, k0 + SK_KSLAB);
  for (int i = threadIdx.x; i < SK_KSLAB * SK_M; i += 128) {
    const int kk = i / SK_M, m = i - kk * SK_M;
    As[kk][m] = (m < M && k0 + kk < k1) ? A[(long)m * a_sm + (long)(k0 + kk) * a_sk] : 0.f;
  }
  __syncthreads();
  if (n >= N) return;
  float acc[SK_M];
#pragma unroll
  for (int m = 0; m < SK_M; ++m) acc[m] = 0.f;
  for (int k = k0; k < k1; ++k) {
    const float bv = Bm[(long)k * b_sk + n];
    const float4* ar = reinterpret_cast<const float4*>(As[k - k0]);
#pragma unroll
    for (int m4 = 0; m4 < SK_M / 4; ++m4) {
      const float4 a4 = ar[m4];
      acc[4 * m4] = fmaf(a4.x, bv, acc[4 * m4]);
      acc[4 * m4 + 1] = fmaf(a4.y, bv, acc[4 * m4 + 1]);
      acc[4 * m4 + 2] = fmaf(a4.z, bv, acc[4 * m4 + 2]);
      acc[4 * m4 + 3] = fmaf(a4.w, bv, acc[4 * m4 + 3]);
    }
  }
#pragma unroll
  for (int m = 0; m < SK_M; ++m)
    if (m < M) atomicAdd(C + (long)m * ldc + n, alpha * acc[m]);
}

// C[m][n] (+)= sum_{k < K <= 32} A[k][m] B[k][n]   (both read along their contiguous dimension): weight gradients
__global__ void __launch_bounds__(256)
skinny_tn_kernel(const float* __restrict__ A, long a_sk, const float* __restrict__ Bm, long b_sk, float* __restrict__ C,
                 long ldc, int M, int N, int K, float alpha, int accumulate) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc = fmaf(A[(long)k * a_sk + m], Bm[(long)k * b_sk + n], acc);
  float* dst = C + (long)m * ldc + n;
  *dst = accumulate ? *dst + alpha * acc : alpha * acc;
}

// C[0:M][0:N] = 0 for a row-strided fp32 matrix.  (cudaMemset2DAsync costs a DMA operation per row: ~4 us each,
// 100-160 us for the 16-32 row outputs of the timestep projections - more than the GEMM it prepares.)
__global__ void zero_rows_kernel(float* __restrict__ c, long ldc, int M, int N) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)M * N) return;
  const long m = i / N;
  c[m * ldc + (i - m * N)] = 0.f;
}
static void zero_rows(float* c, long ldc, int M, int N, cudaStream_t st) {
  if (ldc == N) cudaMemsetAsync(c, 0, sizeof(float) * (size_t)M * N, st);
  else zero_rows_kernel<<<cdiv((long)M * N, 256), 256, 0, st>>>(c, ldc, M, N);
}

// returns 1 if handled
static int skinny_try(const float* a, long a_sm, long a_sk, const float* b, long b_sn, long b_sk, float* c, long ldc,
                      const float* bias, const float* res, long ldr, const float* rowvec, float alpha, int accumulate,
                      int M, int N, int K, cudaStream_t st) {
  if (rowvec) return 0;
  if (M <= 4 * SK_M && a_sk == 1 && b_sk == 1 && (a_sm % 4) == 0 && (b_sn % 4) == 0 && (K % 4) == 0 &&
      ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) {
    const int nblk = cdiv(N, 8 * SK_NW);
    int ks = (148 + nblk - 1) / nblk;                 // enough K slabs to cover the SMs ...
    if (ks > K / 256) ks = K / 256;                   // ... of at least 256 reduction elements each
    if (ks < 1) ks = 1;
    const int kslab = (cdiv(K, ks) + 3) / 4 * 4;
    ks = cdiv(K, kslab);
    if (ks > 1 && !accumulate) zero_rows(c, ldc, M, N, st);
    for (int m0 = 0; m0 < M; m0 += SK_M)
      skinny_nt_kernel<<<dim3(nblk, ks), 256, 0, st>>>(a + m0 * a_sm, a_sm, b, b_sn, c + m0 * ldc, ldc, bias,
                                                        res ? res + m0 * ldr : nullptr, ldr, min(SK_M, M - m0), N, K,
                                                        kslab, alpha, accumulate);
    return check_launch("skinny_nt") == SIDLSG_OK ? 1 : SIDLSG_ERR_CUDA;
  }
  if (M <= 4 * SK_M && b_sn == 1 && !bias && !res) {
    if (!accumulate) zero_rows(c, ldc, M, N, st);
    for (int m0 = 0; m0 < M; m0 += SK_M)
      skinny_nn_kernel<<<dim3(cdiv(N, 128), cdiv(K, SK_KSLAB)), 128, 0, st>>>(
          a + m0 * a_sm, a_sm, a_sk, b, b_sk, c + m0 * ldc, ldc, min(SK_M, M - m0), N, K, alpha);
    return check_launch("skinny_nn") == SIDLSG_OK ? 1 : SIDLSG_ERR_CUDA;
  }
  if (K <= 4 * SK_M && a_sm == 1 && b_sn == 1 && !bias && !res && M <= 65535) {
    skinny_tn_kernel<<<dim3(cdiv(N, 256), M), 256, 0, st>>>(a, a_sk, b, b_sk, c, ldc, M, N, K, alpha, accumulate);
    return check_launch("skinny_tn") == SIDLSG_OK ? 1 : SIDLSG_ERR_CUDA;
  }
  return 0;
}

// tensor-core (tcgen05) paths of gemm_tc.cu: return 1 = handled, 0 = shape not eligible, <0 = error
int tc_gemm_try(const void* a, long a_sm, long a_sk, long a_sb1, long a_sb2, const void* b, long b_sn, long b_sk,
                long b_sb1, long b_sb2, void* c, long ldc, long c_sb1, long c_sb2, const float* bias, const void* res,
                long ldr, const float* rowvec, int rows_per_vec, float alpha, int accumulate, int M, int N, int K,
                int nb1, int nb2, int in_dtype, int out_dtype, cudaStream_t st);
int tc_conv3x3_try(const void* x, const void* w, void* y, const float* bias, const void* res, const float* rowvec,
                   int B, int Hi, int Wi, int Kc, int Ho, int Wo, int N, long w_sn, long w_stap, long w_sk, int stride,
                   int up, int transposed, int flip, int accumulate, int in_dtype, int out_dtype, cudaStream_t st);
int tc_conv3x3_wgrad_try(const void* x, const void* dy, float* dw, int B, int Hi, int Wi, int Cin, int Ho, int Wo,
                         int Cout, long dw_sco, long dw_stap, long dw_sci, int stride, int up, int accumulate,
                         int in_dtype, cudaStream_t st);
// direct kernels for the 4-channel convolutions (conv_small.cu)
int conv_small_try(const void* x, const void* w, void* y, const float* bias, const void* res, const float* rowvec, int B,
                   int Hi, int Wi, int Kc, int Ho, int Wo, int N, long w_sn, long w_stap, long w_sk, int stride, int up,
                   int transposed, int flip, int accumulate, int in_dtype, int out_dtype, cudaStream_t st);
int conv_small_wgrad_try(const void* x, const void* dy, float* dw, int B, int Hi, int Wi, int Cin, int Ho, int Wo,
                         int Cout, long dw_sco, long dw_stap, long dw_sci, int stride, int up, int accumulate,
                         int in_dtype, cudaStream_t st);

}  // namespace sidlsg

using namespace sidlsg;

extern "C" int sidlsg_gemm(const void* a, long a_sm, long a_sk, long a_sb1, long a_sb2,
                           const void* b, long b_sn, long b_sk, long b_sb1, long b_sb2,
                           void* c, long ldc, long c_sb1, long c_sb2,
                           const float* bias, const void* res, long ldr, long r_sb1, long r_sb2,
                           const float* rowvec, int rows_per_vec, float alpha, int accumulate,
                           int M, int N, int K, int nb1, int nb2, int in_dtype, int out_dtype, void* stream) {
  if (M < 0 || N < 0 || K < 0 || nb1 < 1 || nb2 < 1) { set_error("sidlsg_gemm: bad shape"); return SIDLSG_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  int nbatch = nb1 * nb2;
  if (M > 0 && N > 0 && K > 0) {
    int r = tc_gemm_try(a, a_sm, a_sk, a_sb1, a_sb2, b, b_sn, b_sk, b_sb1, b_sb2, c, ldc, c_sb1, c_sb2, bias, res, ldr,
                        rowvec, rows_per_vec, alpha, accumulate, M, N, K, nb1, nb2, in_dtype, out_dtype, st);
    if (r != 0) return r < 0 ? r : SIDLSG_OK;
  }
  if (nbatch == 1 && in_dtype == SIDLSG_F32 && out_dtype == SIDLSG_F32 && M > 0 && N > 0 && K > 0) {
    int r = skinny_try((const float*)a, a_sm, a_sk, (const float*)b, b_sn, b_sk, (float*)c, ldc, bias, (const float*)res,
                       ldr, rowvec, alpha, accumulate, M, N, K, st);
    if (r != 0) return r < 0 ? r : SIDLSG_OK;
  }
  int ksplit = 1;
  if (accumulate == 2) {
    if (out_dtype != SIDLSG_F32) { set_error("sidlsg_gemm: atomic accumulate needs fp32 output"); return SIDLSG_ERR_ARG; }
    ksplit = pick_ksplit(M, N, K, nbatch);
  }
#define RUN(TI, TO_)                                                                                      \
  {                                                                                                       \
    Dense<TI> la{(const TI*)a, a_sm, a_sk, a_sb1, a_sb2, M, K, nb2};                                      \
    Dense<TI> lb{(const TI*)b, b_sk, b_sn, b_sb1, b_sb2, K, N, nb2};                                      \
    Epilogue<TO_> epi{(TO_*)c, ldc, c_sb1, c_sb2, bias, (const TO_*)res, ldr, r_sb1, r_sb2, rowvec,       \
                      rows_per_vec > 0 ? rows_per_vec : 1, alpha, accumulate, M, N, nb2, 0, 0, 0, 0};     \
    return launch_core(la, lb, epi, M, N, K, nbatch, ksplit, a_sk == 1, b_sk == 1, st);                   \
  }
  if (in_dtype == SIDLSG_F32 && out_dtype == SIDLSG_F32) RUN(float, float)
  if (in_dtype == SIDLSG_BF16 && out_dtype == SIDLSG_BF16) RUN(bf16, bf16)
  if (in_dtype == SIDLSG_BF16 && out_dtype == SIDLSG_F32) RUN(bf16, float)
#undef RUN
  set_error("sidlsg_gemm: unsupported dtype pair %d -> %d", in_dtype, out_dtype);
  return SIDLSG_ERR_UNSUPPORTED;
}

// y[B,Ho,Wo,N] = conv3x3(x[B,Hi,Wi,Kc]) ; weights addressed w[n*w_sn + tap*w_stap + kc*w_sk].
// forward: n=cout,kc=cin,flip=0.  dgrad: x:=dy, n=cin, kc=cout, flip=1, transposed=(stride>1).
extern "C" int sidlsg_conv3x3(const void* x, const void* w, void* y, const float* bias, const void* res,
                              const float* rowvec, int B, int Hi, int Wi, int Kc, int Ho, int Wo, int N,
                              long w_sn, long w_stap, long w_sk, int stride, int up, int transposed, int flip,
                              int accumulate, int in_dtype, int out_dtype, void* stream) {
  if (B < 0 || Hi <= 0 || Wi <= 0 || Kc <= 0 || N <= 0 || stride < 1 || up < 1) {
    set_error("sidlsg_conv3x3: bad shape"); return SIDLSG_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (B > 0) {
    int r = tc_conv3x3_try(x, w, y, bias, res, rowvec, B, Hi, Wi, Kc, Ho, Wo, N, w_sn, w_stap, w_sk, stride, up,
                           transposed, flip, accumulate, in_dtype, out_dtype, st);
    if (r != 0) return r < 0 ? r : SIDLSG_OK;
    r = conv_small_try(x, w, y, bias, res, rowvec, B, Hi, Wi, Kc, Ho, Wo, N, w_sn, w_stap, w_sk, stride, up, transposed,
                       flip, accumulate, in_dtype, out_dtype, st);
    if (r != 0) return r < 0 ? r : SIDLSG_OK;
  }
  int M = B * Ho * Wo, K = 9 * Kc;
  ConvGeom g{B, Hi, Wi, Kc, Ho, Wo, stride, up, transposed};
#define RUN(TI, TO_)                                                                                      \
  {                                                                                                       \
    ConvGather<TI, true> la{(const TI*)x, g, M, K};                                                       \
    ConvWeight<TI> lb{(const TI*)w, w_sn, w_stap, w_sk, Kc, N, flip};                                     \
    Epilogue<TO_> epi{(TO_*)y, (long)N, 0, 0, bias, (const TO_*)res, (long)N, 0, 0, rowvec, Ho * Wo, 1.f, \
                      accumulate, M, N, 1, 0, 0, 0, 0};                                                   \
    return launch_core(la, lb, epi, M, N, K, 1, 1, 1, w_sk == 1, st);                                     \
  }
  if (in_dtype == SIDLSG_F32 && out_dtype == SIDLSG_F32) RUN(float, float)
  if (in_dtype == SIDLSG_BF16 && out_dtype == SIDLSG_BF16) RUN(bf16, bf16)
  if (in_dtype == SIDLSG_BF16 && out_dtype == SIDLSG_F32) RUN(bf16, float)
#undef RUN
  set_error("sidlsg_conv3x3: unsupported dtype pair");
  return SIDLSG_ERR_UNSUPPORTED;
}

// dw[co*sco + tap*stap + ci*sci] (+)= sum_pix dy[pix,co] * gather(x)[pix, tap, ci]   (fp32 output)
extern "C" int sidlsg_conv3x3_wgrad(const void* x, const void* dy, float* dw, int B, int Hi, int Wi, int Cin,
                                    int Ho, int Wo, int Cout, long dw_sco, long dw_stap, long dw_sci,
                                    int stride, int up, int accumulate, int in_dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B > 0) {
    int r = tc_conv3x3_wgrad_try(x, dy, dw, B, Hi, Wi, Cin, Ho, Wo, Cout, dw_sco, dw_stap, dw_sci, stride, up,
                                 accumulate, in_dtype, st);
    if (r != 0) return r < 0 ? r : SIDLSG_OK;
    r = conv_small_wgrad_try(x, dy, dw, B, Hi, Wi, Cin, Ho, Wo, Cout, dw_sco, dw_stap, dw_sci, stride, up, accumulate,
                             in_dtype, st);
    if (r != 0) return r < 0 ? r : SIDLSG_OK;
  }
  int M = Cout, N = 9 * Cin, K = B * Ho * Wo;
  ConvGeom g{B, Hi, Wi, Cin, Ho, Wo, stride, up, 0};
  int ksplit = pick_ksplit(M, N, K, 1);
  if (!accumulate) {
    // split-K accumulates atomically -> caller must pass a zeroed buffer when accumulate==0 and ksplit>1
    if (ksplit > 1) {
      if (!(dw_sci == 1 && dw_stap == Cin && dw_sco == 9L * Cin)) ksplit = 1;
      else cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * 9 * Cin, st);
    }
  }
  int acc_mode = (ksplit > 1 || accumulate) ? 2 : 0;
#define RUN(TI)                                                                                           \
  {                                                                                                       \
    Dense<TI> la{(const TI*)dy, 1, (long)Cout, 0, 0, M, K, 1};                                            \
    ConvGather<TI, true> lb{(const TI*)x, g, K, N};                                                      \
    Epilogue<float> epi{dw, 0, 0, 0, nullptr, nullptr, 0, 0, 0, nullptr, 1, 1.f, acc_mode, M, N, 1,       \
                        Cin, dw_sco, dw_stap, dw_sci};                                                    \
    return launch_core(la, lb, epi, M, N, K, 1, ksplit, 0, 0, st);                                        \
  }
  if (in_dtype == SIDLSG_F32) RUN(float)
  if (in_dtype == SIDLSG_BF16) RUN(bf16)
#undef RUN
  set_error("sidlsg_conv3x3_wgrad: unsupported dtype");
  return SIDLSG_ERR_UNSUPPORTED;
}
