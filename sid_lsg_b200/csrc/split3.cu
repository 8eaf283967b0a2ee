// fp32-accurate contractions ON THE TENSOR CORES: x = hi + lo with hi = bf16(x), lo = bf16(x - hi), and
//   A B  ~=  A_hi B_hi + A_hi B_lo + A_lo B_hi        (error ~2^-16 |a||b| per product, fp32 accumulation in TMEM)
// run as three passes of the SAME tcgen05 kernel the bf16 throughput mode uses (gemm_tc.cu: same TMA maps, MMA
// descriptors, pipeline, epilogue; the 2nd and 3rd pass add into the fp32 output with red.global.add).  This is the
// parity instrument BASELINE.json's north_star asks for ("match the reference fp32 path within 1e-3"): the reference
// runs fp32 with TF32 off (/root/reference/training/sid_training_loop.py:241-243), which a single bf16 pass cannot
// match, and the CUDA-core kernels of gemm_simt.cu - which can - are not the kernels that are benchmarked.
// Operands are packed (and, if needed, transposed) into dense K-major bf16 pairs inside a caller-provided workspace,
// so every strided / batched / MN-major call shape of sidlsg_gemm maps onto the best-tested kk-mode of the kernel.
// Speed is irrelevant here (3 passes + packing).
#include "common.cuh"

namespace sidlsg {

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
bool tc_enabled();

// src element (z1, z2, r, k) at src[z1*sb1 + z2*sb2 + r*sr + k*sk]  ->  hi/lo[((z1*nb2 + z2)*R + r)*Kp + k], k >= K zero
__global__ void __launch_bounds__(256)
pack_split_kernel(const float* __restrict__ src, long sr, long sk, long sb1, long sb2, int R, int K, int Kp, int nb2,
                  long total, bf16* __restrict__ hi, bf16* __restrict__ lo) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    const long t = i / Kp;
    const int r = (int)(t % R);
    const long z = t / R;
    float v = 0.f;
    if (k < K) v = src[(z / nb2) * sb1 + (z % nb2) * sb2 + (long)r * sr + (long)k * sk];
    const bf16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// dense, same layout: n elements
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ src, long n, bf16* __restrict__ hi, bf16* __restrict__ lo) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = src[i];
    const bf16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

static inline long align256(long b) { return (b + 255) & ~255L; }
static inline int blocks_for(long n) { long b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

}  // namespace sidlsg

using namespace sidlsg;

// workspace for the three entry points below: two bf16 copies (hi, lo) of each of the two operands
extern "C" long sidlsg_split3_ws_bytes(long a_elems, long b_elems) {
  return 2 * align256(a_elems * 2) + 2 * align256(b_elems * 2) + 1024;
}

// sidlsg_gemm semantics for fp32 operands and fp32 output (accumulate 0 = store, 2 = atomic +=).  a_elems / b_elems
// for the workspace query: nb1*nb2*M*ceil8(K) and nb1*nb2*N*ceil8(K).
// Returns SIDLSG_ERR_UNSUPPORTED when the tensor-core kernel cannot take the shape (caller uses sidlsg_gemm).
extern "C" int sidlsg_gemm_split3(const float* a, long a_sm, long a_sk, long a_sb1, long a_sb2,
                                  const float* b, long b_sn, long b_sk, long b_sb1, long b_sb2,
                                  float* c, long ldc, long c_sb1, long c_sb2,
                                  const float* bias, const float* res, long ldr, long r_sb1, long r_sb2,
                                  const float* rowvec, int rows_per_vec, float alpha, int accumulate,
                                  int M, int N, int K, int nb1, int nb2, void* ws, long ws_bytes, void* stream) {
  if (!tc_enabled() || M < 64 || N < 16 || K < 64 || accumulate == 1 || (ldc & 3) || (N & 15)) {
    set_error("gemm_split3: shape M=%d N=%d K=%d not eligible", M, N, K);
    return SIDLSG_ERR_UNSUPPORTED;
  }
  if (res && ((ldr & 7) || (reinterpret_cast<uintptr_t>(res) & 15))) { set_error("gemm_split3: residual alignment"); return SIDLSG_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  const int Kp = (K + 7) & ~7;
  const long nb = (long)nb1 * nb2;
  const long ae = nb * M * Kp, be = nb * N * Kp;
  if (!ws || ws_bytes < sidlsg_split3_ws_bytes(ae, be)) { set_error("gemm_split3: workspace too small"); return SIDLSG_ERR_ARG; }
  uint8_t* w8 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  bf16* ah = reinterpret_cast<bf16*>(w8);
  bf16* al = reinterpret_cast<bf16*>(w8 + align256(ae * 2));
  bf16* bh = reinterpret_cast<bf16*>(w8 + 2 * align256(ae * 2));
  bf16* bl = reinterpret_cast<bf16*>(w8 + 2 * align256(ae * 2) + align256(be * 2));
  pack_split_kernel<<<blocks_for(ae), 256, 0, st>>>(a, a_sm, a_sk, a_sb1, a_sb2, M, K, Kp, nb2, ae, ah, al);
  pack_split_kernel<<<blocks_for(be), 256, 0, st>>>(b, b_sn, b_sk, b_sb1, b_sb2, N, K, Kp, nb2, be, bh, bl);
  int r = check_launch("gemm_split3 pack");
  if (r != SIDLSG_OK) return r;
  for (long z = 0; z < nb; ++z) {
    const long z1 = z / nb2, z2 = z % nb2;
    float* cz = c + z1 * c_sb1 + z2 * c_sb2;
    const float* rz = res ? res + z1 * r_sb1 + z2 * r_sb2 : nullptr;
    const bf16 *azh = ah + z * M * Kp, *azl = al + z * M * Kp, *bzh = bh + z * N * Kp, *bzl = bl + z * N * Kp;
    const bf16* pa[3] = {azh, azh, azl};
    const bf16* pb[3] = {bzh, bzl, bzh};
    for (int pass = 0; pass < 3; ++pass) {
      const bool first = pass == 0;
      r = tc_gemm_try(pa[pass], Kp, 1, 0, 0, pb[pass], Kp, 1, 0, 0, cz, ldc, 0, 0, first ? bias : nullptr,
                      first ? rz : nullptr, ldr, first ? rowvec : nullptr, rows_per_vec, alpha,
                      (first && accumulate == 0) ? 0 : 2, M, N, Kp, 1, 1, SIDLSG_BF16, SIDLSG_F32, st);
      if (r < 0) return r;
      if (r == 0) {
        set_error("gemm_split3: tensor-core kernel declined M=%d N=%d K=%d (pass %d)", M, N, K, pass);
        return pass == 0 ? SIDLSG_ERR_UNSUPPORTED : SIDLSG_ERR_CUDA;
      }
    }
  }
  return SIDLSG_OK;
}

// sidlsg_conv3x3 semantics (stride 1 or 2, up = 1, not transposed) for dense fp32 x [B,Hi,Wi,Kc], fp32 weights addressed
// w[n*w_sn + tap*w_stap + kc*w_sk] inside a dense block of w_elems floats, fp32 y/res.  Workspace: (B*Hi*Wi*Kc, w_elems).
extern "C" int sidlsg_conv3x3_split3(const float* x, const float* w, long w_elems, float* y, const float* bias,
                                     const float* res, const float* rowvec, int B, int Hi, int Wi, int Kc, int Ho,
                                     int Wo, int N, long w_sn, long w_stap, long w_sk, int stride, int flip,
                                     void* ws, long ws_bytes, void* stream) {
  if (!tc_enabled()) { set_error("conv3x3_split3: no tcgen05 device"); return SIDLSG_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  const long xe = (long)B * Hi * Wi * Kc;
  if (!ws || ws_bytes < sidlsg_split3_ws_bytes(xe, w_elems)) { set_error("conv3x3_split3: workspace too small"); return SIDLSG_ERR_ARG; }
  uint8_t* w8 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  bf16* xh = reinterpret_cast<bf16*>(w8);
  bf16* xl = reinterpret_cast<bf16*>(w8 + align256(xe * 2));
  bf16* wh = reinterpret_cast<bf16*>(w8 + 2 * align256(xe * 2));
  bf16* wl = reinterpret_cast<bf16*>(w8 + 2 * align256(xe * 2) + align256(w_elems * 2));
  split_kernel<<<blocks_for(xe), 256, 0, st>>>(x, xe, xh, xl);
  split_kernel<<<blocks_for(w_elems), 256, 0, st>>>(w, w_elems, wh, wl);
  int r = check_launch("conv3x3_split3 split");
  if (r != SIDLSG_OK) return r;
  const bf16* px[3] = {xh, xh, xl};
  const bf16* pw[3] = {wh, wl, wh};
  for (int pass = 0; pass < 3; ++pass) {
    const bool first = pass == 0;
    r = tc_conv3x3_try(px[pass], pw[pass], y, first ? bias : nullptr, first ? res : nullptr, first ? rowvec : nullptr, B,
                       Hi, Wi, Kc, Ho, Wo, N, w_sn, w_stap, w_sk, stride, 1, 0, flip, first ? 0 : 2, SIDLSG_BF16,
                       SIDLSG_F32, st);
    if (r < 0) return r;
    if (r == 0) {
      set_error("conv3x3_split3: tensor-core kernel declined B=%d H=%d C=%d N=%d (pass %d)", B, Hi, Kc, N, pass);
      return pass == 0 ? SIDLSG_ERR_UNSUPPORTED : SIDLSG_ERR_CUDA;
    }
  }
  return SIDLSG_OK;
}

// sidlsg_conv3x3_wgrad semantics (accumulating into dw) for dense fp32 x [B,Hi,Wi,Cin] and dy [B,Ho,Wo,Cout].
extern "C" int sidlsg_conv3x3_wgrad_split3(const float* x, const float* dy, float* dw, int B, int Hi, int Wi, int Cin,
                                           int Ho, int Wo, int Cout, long dw_sco, long dw_stap, long dw_sci, int stride,
                                           void* ws, long ws_bytes, void* stream) {
  if (!tc_enabled()) { set_error("conv3x3_wgrad_split3: no tcgen05 device"); return SIDLSG_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  const long xe = (long)B * Hi * Wi * Cin, ye = (long)B * Ho * Wo * Cout;
  if (!ws || ws_bytes < sidlsg_split3_ws_bytes(xe, ye)) { set_error("conv3x3_wgrad_split3: workspace too small"); return SIDLSG_ERR_ARG; }
  uint8_t* w8 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  bf16* xh = reinterpret_cast<bf16*>(w8);
  bf16* xl = reinterpret_cast<bf16*>(w8 + align256(xe * 2));
  bf16* yh = reinterpret_cast<bf16*>(w8 + 2 * align256(xe * 2));
  bf16* yl = reinterpret_cast<bf16*>(w8 + 2 * align256(xe * 2) + align256(ye * 2));
  split_kernel<<<blocks_for(xe), 256, 0, st>>>(x, xe, xh, xl);
  split_kernel<<<blocks_for(ye), 256, 0, st>>>(dy, ye, yh, yl);
  int r = check_launch("conv3x3_wgrad_split3 split");
  if (r != SIDLSG_OK) return r;
  const bf16* px[3] = {xh, xl, xh};
  const bf16* py[3] = {yh, yh, yl};
  for (int pass = 0; pass < 3; ++pass) {
    r = tc_conv3x3_wgrad_try(px[pass], py[pass], dw, B, Hi, Wi, Cin, Ho, Wo, Cout, dw_sco, dw_stap, dw_sci, stride, 1, 1,
                             SIDLSG_BF16, st);
    if (r < 0) return r;
    if (r == 0) {
      set_error("conv3x3_wgrad_split3: tensor-core kernel declined (pass %d)", pass);
      return pass == 0 ? SIDLSG_ERR_UNSUPPORTED : SIDLSG_ERR_CUDA;
    }
  }
  return SIDLSG_OK;
}
