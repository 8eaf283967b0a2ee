// 3x3 convolutions with a 4-channel side: conv_in (4 -> 320) and conv_out (320 -> 4) of the UNet, forward, data
// gradient and weight gradient (diffusers UNet2DConditionModel.conv_in / conv_out behind
// /root/reference/training/sid_sd_util.py:184,245,263).  K = 36 or N = 4 is far too thin for a 128-wide MMA tile, so
// these are direct CUDA-core kernels bound by the wide tensor's HBM traffic:
//   narrow_in  : y[pix][n]  = sum_{tap,c<Cs} xs[pix+tap][c] w(n,tap,c)   (Cs <= 8; conv_in fwd, conv_out dgrad)
//   narrow_out : y[pix][n<Ns] = sum_{tap,c} xw[pix+tap][c] w(n,tap,c)    (Ns <= 8; conv_out fwd, conv_in dgrad)
//   wgrad      : dw = sum_pix wide[pix(+tap)][c] * narrow[pix(+tap)][j]  (both layers)
// Weights are addressed w[n*w_sn + tap*w_stap + c*w_sk] with optional tap flip, exactly like sidlsg_conv3x3.
#include "common.cuh"

namespace sidlsg {

constexpr int CS_MAX = 8;

// ---- narrow input: one thread = one pixel x 8 output channels -------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
conv_narrow_in_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y, const float* __restrict__ bias,
                      const float* __restrict__ rowvec, int B, int H, int W, int Cs, int N, long w_sn, long w_stap,
                      long w_sk, int flip) {
  extern __shared__ float ws[];   // [9*CS_MAX][N], zero padded for c >= Cs: every loop below has static bounds
  constexpr int KK = 9 * CS_MAX;
  for (int i = threadIdx.x; i < N * KK; i += blockDim.x) {
    const int k = i / N, n = i - k * N, tap = k / CS_MAX, c = k - tap * CS_MAX;
    ws[i] = c < Cs ? to_f(w[(long)n * w_sn + (long)(flip ? 8 - tap : tap) * w_stap + (long)c * w_sk]) : 0.f;
  }
  __syncthreads();
  const int ngrp = N / 8;
  const long total = (long)B * H * W * ngrp;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % ngrp);
    const long pix = idx / ngrp;
    const int ox = (int)(pix % W);
    const long r = pix / W;
    const int oy = (int)(r % H);
    const long b = r / H;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = g * 8 + j;
      acc[j] = (bias ? bias[n] : 0.f) + (rowvec ? rowvec[b * N + n] : 0.f);
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const T* px = x + ((b * H + iy) * W + ix) * Cs;
#pragma unroll
      for (int c = 0; c < CS_MAX; ++c) {
        if (c < Cs) {
          const float xv = to_f(px[c]);
          const float4 w0 = *reinterpret_cast<const float4*>(ws + (tap * CS_MAX + c) * N + g * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + (tap * CS_MAX + c) * N + g * 8 + 4);
          acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
          acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
          acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
          acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
        }
      }
    }
    T* py = y + pix * N + g * 8;
    if (sizeof(T) == 2) {
      uint4 o;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
      *reinterpret_cast<uint4*>(py) = o;
    } else {
      *reinterpret_cast<float4*>(py) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(py + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

// ---- narrow output: one warp = one pixel, lanes split the wide channels -------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
conv_narrow_out_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y, const float* __restrict__ bias,
                       int B, int H, int W, int C, int Ns, long w_sn, long w_stap, long w_sk, int flip) {
  extern __shared__ float ws[];   // [Ns][9][C]
  for (int i = threadIdx.x; i < Ns * 9 * C; i += blockDim.x) {
    const int n = i / (9 * C), r = i - n * 9 * C, tap = r / C, c = r - tap * C;
    ws[i] = to_f(w[(long)n * w_sn + (long)(flip ? 8 - tap : tap) * w_stap + (long)c * w_sk]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long npix = (long)B * H * W;
  for (long pix = (long)blockIdx.x * 8 + (threadIdx.x >> 5); pix < npix; pix += (long)gridDim.x * 8) {
    const int ox = (int)(pix % W);
    const long r = pix / W;
    const int oy = (int)(r % H);
    const long b = r / H;
    float acc[CS_MAX];
#pragma unroll
    for (int j = 0; j < CS_MAX; ++j) acc[j] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const T* px = x + ((b * H + iy) * W + ix) * C;
      for (int c = lane * 2; c < C; c += 64) {
        const float x0 = to_f(px[c]), x1 = to_f(px[c + 1]);
#pragma unroll
        for (int j = 0; j < CS_MAX; ++j)
          if (j < Ns) {
            const float* wj = ws + (j * 9 + tap) * C + c;
            acc[j] = fmaf(x0, wj[0], fmaf(x1, wj[1], acc[j]));
          }
      }
    }
#pragma unroll
    for (int j = 0; j < CS_MAX; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0)
      for (int j = 0; j < Ns; ++j) y[pix * Ns + j] = from_f<T>(acc[j] + (bias ? bias[j] : 0.f));
  }
}

// ---- weight gradients: thread = one wide channel, 9*Cs accumulators, block = a chunk of pixels ------------------------
// wide_is_dy = 1 (conv_in):  dw[c][tap][j] += dy[pix][c] * xs[pix+tap][j]     wide = dy [.,Cw], narrow = x [.,Cs]
// wide_is_dy = 0 (conv_out): dw[j][tap][c] += dy[pix][j] * xw[pix+tap][c]     wide = x  [.,Cw], narrow = dy [.,Cs]
template <class T>
__global__ void __launch_bounds__(320)
conv_small_wgrad_kernel(const T* __restrict__ wide, const T* __restrict__ narrow, float* __restrict__ dw, int B, int H,
                        int W, int Cw, int Cs, int wide_is_dy, int pix_per_block) {
  __shared__ float nb[9 * CS_MAX];
  const long npix = (long)B * H * W;
  const long p0 = (long)blockIdx.x * pix_per_block, p1 = min(npix, p0 + pix_per_block);
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  float acc[9 * CS_MAX];
#pragma unroll
  for (int i = 0; i < 9 * CS_MAX; ++i) acc[i] = 0.f;
  for (long pix = p0; pix < p1; ++pix) {
    const int ox = (int)(pix % W);
    const long r = pix / W;
    const int oy = (int)(r % H);
    const long b = r / H;
    __syncthreads();
    if (wide_is_dy) {
      // narrow window of x around pix
      if (threadIdx.x < 9 * CS_MAX) {
        const int tap = threadIdx.x / CS_MAX, j = threadIdx.x % CS_MAX;
        const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
        const bool ok = j < Cs && iy >= 0 && iy < H && ix >= 0 && ix < W;
        nb[threadIdx.x] = ok ? to_f(narrow[((b * H + iy) * W + ix) * Cs + j]) : 0.f;
      }
      __syncthreads();
      if (c < Cw) {
        const float d = to_f(wide[pix * Cw + c]);
#pragma unroll
        for (int i = 0; i < 9 * CS_MAX; ++i) acc[i] = fmaf(d, nb[i], acc[i]);
      }
    } else {
      if (threadIdx.x < CS_MAX) nb[threadIdx.x] = threadIdx.x < Cs ? to_f(narrow[pix * Cs + threadIdx.x]) : 0.f;
      __syncthreads();
      if (c < Cw) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int iy = oy + tap / 3 - 1, ix = ox + tap % 3 - 1;
          if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
          const float xv = to_f(wide[((b * H + iy) * W + ix) * Cw + c]);
#pragma unroll
          for (int j = 0; j < CS_MAX; ++j) acc[tap * CS_MAX + j] = fmaf(xv, nb[j], acc[tap * CS_MAX + j]);
        }
      }
    }
  }
  if (c >= Cw) return;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int j = 0; j < CS_MAX; ++j) {
      if (j >= Cs) continue;
      // physical weight layout [Cout][3][3][Cin]
      float* dst = wide_is_dy ? dw + ((long)c * 9 + tap) * Cs + j : dw + ((long)j * 9 + tap) * Cw + c;
      atomicAdd(dst, acc[tap * CS_MAX + j]);
    }
}


// =====================================================================================================================
// The same two layers on the TENSOR CORES (bf16 mode): a 4-channel 3x3 convolution is a GEMM with K = 36 (or N = 36)
// once the narrow tensor's 3x3 windows are written out ("im2col": 9 * Cs <= 64 values per pixel, 33 MB for 64 x 64 x 64
// pixels) - the direct kernels above spend ~400 instructions per 16-byte store and ran 10-16x above the HBM time of the
// wide tensor (conv_in forward 538 us, conv_out forward 601 us, their weight gradients 700 us per call at n = 32).
// These helpers do the data movement; the contractions themselves are sidlsg_gemm calls composed in ops.py.
//   off(tap) = (tap / 3 - 1, tap % 3 - 1);  k(tap, c) = tap * Cs + c  (layout 0)  or  c * 9 + tap  (layout 1)
// col[p][k(tap, c)] = x[p + sign * off(tap)][c]   (0 outside the image; columns >= 9 * Cs zero), col is [M, 64] bf16
__global__ void __launch_bounds__(256)
narrow_im2col_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int B, int H, int W, int Cs, int sign, int layout) {
  const long npix = (long)B * H * W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;    // one thread = one pixel x one 8-column piece
  if (idx >= npix * 8) return;
  const long pix = idx >> 3;
  const int piece = (int)(idx & 7);
  const int ox = (int)(pix % W);
  const long r = pix / W;
  const int oy = (int)(r % H);
  const long b = r / H;
  uint4 out;
  bf16* o = reinterpret_cast<bf16*>(&out);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = piece * 8 + j;
    bf16 v = __float2bfloat16_rn(0.f);
    if (k < 9 * Cs) {
      const int tap = layout ? k % 9 : k / Cs, c = layout ? k / 9 : k % Cs;
      const int iy = oy + sign * (tap / 3 - 1), ix = ox + sign * (tap % 3 - 1);
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[((b * H + iy) * W + ix) * Cs + c];
    }
    o[j] = v;
  }
  *reinterpret_cast<uint4*>(col + pix * 64 + piece * 8) = out;
}

// y[p][c] = bias[c] + sum_tap col[p + sign * off(tap)][k(tap, c)]     (col [M, ld] bf16, y [M, Cs] bf16 or fp32 NCHW-free)
__global__ void __launch_bounds__(256)
narrow_col2im_kernel(const bf16* __restrict__ col, int ld, bf16* __restrict__ y, const float* __restrict__ bias, int B, int H,
                     int W, int Cs, int sign, int layout) {
  const long npix = (long)B * H * W;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;    // one thread = one pixel x one channel
  if (idx >= npix * Cs) return;
  const long pix = idx / Cs;
  const int c = (int)(idx - pix * Cs);
  const int ox = (int)(pix % W);
  const long r = pix / W;
  const int oy = (int)(r % H);
  const long b = r / H;
  float acc = bias ? bias[c] : 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int iy = oy + sign * (tap / 3 - 1), ix = ox + sign * (tap % 3 - 1);
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    acc += to_f(col[((b * H + iy) * W + ix) * ld + (layout ? c * 9 + tap : tap * Cs + c)]);
  }
  y[idx] = __float2bfloat16_rn(acc);
}

// dst [Rp, Kp] = zero-padded copy of src [R, K] (row stride lds); bf16
__global__ void __launch_bounds__(256)
pad2d_kernel(const bf16* __restrict__ src, long lds, bf16* __restrict__ dst, int R, int K, int Rp, int Kp) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)Rp * Kp) return;
  const int r = (int)(idx / Kp), k = (int)(idx - (long)r * Kp);
  dst[idx] = (r < R && k < K) ? src[(long)r * lds + k] : __float2bfloat16_rn(0.f);
}

// dst[k][r] += src[r][k]   (src [R, lds] fp32, first K columns; dst [K, R] fp32)
__global__ void __launch_bounds__(256)
add_transposed_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int R, int K) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)R * K) return;
  const int k = (int)(idx / R), r = (int)(idx - (long)k * R);
  dst[idx] += src[(long)r * lds + k];
}

// returns 1 if handled, 0 if not eligible, <0 on error
int conv_small_try(const void* x, const void* w, void* y, const float* bias, const void* res, const float* rowvec, int B,
                   int Hi, int Wi, int Kc, int Ho, int Wo, int N, long w_sn, long w_stap, long w_sk, int stride, int up,
                   int transposed, int flip, int accumulate, int in_dtype, int out_dtype, cudaStream_t st) {
  if (stride != 1 || up != 1 || transposed || accumulate || res || in_dtype != out_dtype) return 0;
  if (Hi != Ho || Wi != Wo) return 0;
  const long npix = (long)B * Hi * Wi;
  if (Kc <= CS_MAX && N % 8 == 0 && N * 9 * CS_MAX * 4 <= 96 * 1024) {
    const size_t sm = sizeof(float) * N * 9 * CS_MAX;
    const long total = npix * (N / 8);
    const int blocks = (int)min((long)148 * 16, (total + 255) / 256);
#define RUN(T)                                                                                                        \
    {                                                                                                                 \
      cudaFuncSetAttribute(conv_narrow_in_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);         \
      conv_narrow_in_kernel<T><<<blocks, 256, sm, st>>>((const T*)x, (const T*)w, (T*)y, bias, rowvec, B, Hi, Wi, Kc, N, \
                                                        w_sn, w_stap, w_sk, flip);                                    \
    }
    if (in_dtype == SIDLSG_F32) RUN(float) else RUN(bf16)
#undef RUN
    return check_launch("conv_narrow_in") == SIDLSG_OK ? 1 : SIDLSG_ERR_CUDA;
  }
  if (N <= CS_MAX && !rowvec && Kc % 2 == 0 && (long)N * 9 * Kc * 4 <= 160 * 1024) {
    const size_t sm = sizeof(float) * N * 9 * Kc;
    const int blocks = (int)min((long)148 * 8, (npix + 7) / 8);
#define RUN(T)                                                                                                        \
    {                                                                                                                 \
      cudaFuncSetAttribute(conv_narrow_out_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);       \
      conv_narrow_out_kernel<T><<<blocks, 256, sm, st>>>((const T*)x, (const T*)w, (T*)y, bias, B, Hi, Wi, Kc, N, w_sn,  \
                                                         w_stap, w_sk, flip);                                         \
    }
    if (in_dtype == SIDLSG_F32) RUN(float) else RUN(bf16)
#undef RUN
    return check_launch("conv_narrow_out") == SIDLSG_OK ? 1 : SIDLSG_ERR_CUDA;
  }
  return 0;
}

int conv_small_wgrad_try(const void* x, const void* dy, float* dw, int B, int Hi, int Wi, int Cin, int Ho, int Wo,
                         int Cout, long dw_sco, long dw_stap, long dw_sci, int stride, int up, int accumulate,
                         int in_dtype, cudaStream_t st) {
  if (stride != 1 || up != 1 || Hi != Ho || Wi != Wo) return 0;
  if (!(dw_sci == 1 && dw_stap == Cin && dw_sco == 9L * Cin)) return 0;
  const bool small_in = Cin <= CS_MAX, small_out = Cout <= CS_MAX;
  if (small_in == small_out) return 0;
  const long npix = (long)B * Hi * Wi;
  if (!accumulate) cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * 9 * Cin, st);
  const int Cw = small_in ? Cout : Cin, Cs = small_in ? Cin : Cout;
  int ppb = 2048;   // few, fat blocks: every block ends with 9*Cs*Cw atomics into the same small dw
  while (ppb > 32 && (npix + ppb - 1) / ppb * ((Cw + 319) / 320) < 148) ppb >>= 1;
  dim3 grid((unsigned)((npix + ppb - 1) / ppb), (Cw + 319) / 320);
  const void* wide = small_in ? dy : x;
  const void* narrow = small_in ? x : dy;
  if (in_dtype == SIDLSG_F32)
    conv_small_wgrad_kernel<float><<<grid, 320, 0, st>>>((const float*)wide, (const float*)narrow, dw, B, Hi, Wi, Cw, Cs,
                                                         small_in ? 1 : 0, ppb);
  else
    conv_small_wgrad_kernel<bf16><<<grid, 320, 0, st>>>((const bf16*)wide, (const bf16*)narrow, dw, B, Hi, Wi, Cw, Cs,
                                                        small_in ? 1 : 0, ppb);
  return check_launch("conv_small_wgrad") == SIDLSG_OK ? 1 : SIDLSG_ERR_CUDA;
}

}  // namespace sidlsg

using namespace sidlsg;

// ---- data-movement helpers of the tensor-core form of the 4-channel convolutions (see narrow_im2col_kernel) ----------
extern "C" int sidlsg_narrow_im2col(const void* x, void* col, int B, int H, int W, int Cs, int sign, int layout,
                                    void* stream) {
  if (Cs < 1 || 9 * Cs > 64) { set_error("narrow_im2col: Cs=%d", Cs); return SIDLSG_ERR_ARG; }
  const long n = (long)B * H * W * 8;
  if (n == 0) return SIDLSG_OK;
  narrow_im2col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)col, B, H, W,
                                                                                       Cs, sign, layout);
  return check_launch("narrow_im2col");
}
extern "C" int sidlsg_narrow_col2im(const void* col, int ld, void* y, const float* bias, int B, int H, int W, int Cs,
                                    int sign, int layout, void* stream) {
  if (Cs < 1 || 9 * Cs > ld) { set_error("narrow_col2im: Cs=%d ld=%d", Cs, ld); return SIDLSG_ERR_ARG; }
  const long n = (long)B * H * W * Cs;
  if (n == 0) return SIDLSG_OK;
  narrow_col2im_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)col, ld, (bf16*)y, bias,
                                                                                       B, H, W, Cs, sign, layout);
  return check_launch("narrow_col2im");
}
extern "C" int sidlsg_pad2d(const void* src, long lds, void* dst, int R, int K, int Rp, int Kp, void* stream) {
  const long n = (long)Rp * Kp;
  if (n <= 0 || R > Rp || K > Kp) { set_error("pad2d: bad shape"); return SIDLSG_ERR_ARG; }
  pad2d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, lds, (bf16*)dst, R, K, Rp, Kp);
  return check_launch("pad2d");
}
extern "C" int sidlsg_add_transposed(const float* src, int lds, float* dst, int R, int K, void* stream) {
  const long n = (long)R * K;
  if (n <= 0) return SIDLSG_OK;
  add_transposed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, lds, dst, R, K);
  return check_launch("add_transposed");
}
