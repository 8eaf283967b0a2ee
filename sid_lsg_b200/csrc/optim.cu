// Fused optimiser pass over the flat parameter bucket of one network: gradient averaging over ranks,
// nan_to_num(nan=0, +-inf -> +-1e5), optional value clip, Adam, optional EMA lerp and the bf16 compute
// shadow - one streaming pass instead of the reference's three per-parameter Python loops + foreach Adam:
//   nan_to_num  /root/reference/training/sid_training_loop.py:458-460, 541-543
//   clip (fp16) /root/reference/training/sid_training_loop.py:546-547
//   Adam        /root/reference/sid_train.py:219-226 (betas (0, 0.999), eps 1e-8) = torch.optim.Adam semantics
//   EMA         /root/reference/training/sid_training_loop.py:553-565  (p_ema <- p + beta (p_ema - p), after the step)
// HBM-bound: reads p, g, v (+m, +ema), writes p, v (+m, +ema, +bf16 shadows of p and ema): 24-36 B per parameter.
#include "common.cuh"

namespace sidlsg {

struct AdamArgs {
  float* p;
  const float* g;
  float* m;       // null when beta1 == 0 (m == g)
  float* v;
  float* ema;     // null: no EMA
  bf16* shadow;   // null: no bf16 shadow
  bf16* ema_shadow;  // null: the EMA network keeps no bf16 shadow
  long n;
  float lr, beta1, beta2, eps, bc1, bc2_sqrt, grad_scale, clip, ema_beta, weight_decay;
  const float* hyper;   // optional device float[4] {lr, bc1, bc2_sqrt, ema_beta} read at run time (CUDA-graph replay: the
                        // per-step scalars must not be baked into the captured launch)
};

__device__ __forceinline__ float sanitize(float g) {
  if (isnan(g)) return 0.f;
  if (isinf(g)) return g > 0.f ? 1e5f : -1e5f;
  return g;
}

__global__ void __launch_bounds__(256) adam_kernel(AdamArgs a) {
  if (a.hyper) {
    const float4 h = *reinterpret_cast<const float4*>(a.hyper);
    a.lr = h.x; a.bc1 = h.y; a.bc2_sqrt = h.z; a.ema_beta = h.w;
  }
  const long nvec = a.n >> 2;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    float4 g = reinterpret_cast<const float4*>(a.g)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    float4 m = a.m ? reinterpret_cast<float4*>(a.m)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 e = a.ema ? reinterpret_cast<float4*>(a.ema)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
#define ADAM_LANE(c)                                                         \
    {                                                                        \
      float gg = sanitize(g.c * a.grad_scale);                               \
      if (a.clip > 0.f) gg = fminf(fmaxf(gg, -a.clip), a.clip);              \
      if (a.weight_decay != 0.f) p.c *= 1.f - a.lr * a.weight_decay;         \
      float mm = a.m ? a.beta1 * m.c + (1.f - a.beta1) * gg : gg;            \
      m.c = mm;                                                              \
      v.c = a.beta2 * v.c + (1.f - a.beta2) * gg * gg;                       \
      float denom = sqrtf(v.c) / a.bc2_sqrt + a.eps;                         \
      p.c -= (a.lr / a.bc1) * (mm / denom);                                  \
      e.c = p.c + a.ema_beta * (e.c - p.c);                                  \
    }
    ADAM_LANE(x) ADAM_LANE(y) ADAM_LANE(z) ADAM_LANE(w)
#undef ADAM_LANE
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.v)[i] = v;
    if (a.m) reinterpret_cast<float4*>(a.m)[i] = m;
    if (a.ema) reinterpret_cast<float4*>(a.ema)[i] = e;
    if (a.shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
      uint2 pk;
      pk.x = *reinterpret_cast<unsigned*>(&lo);
      pk.y = *reinterpret_cast<unsigned*>(&hi);
      reinterpret_cast<uint2*>(a.shadow)[i] = pk;
    }
    if (a.ema && a.ema_shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(e.x, e.y), hi = __floats2bfloat162_rn(e.z, e.w);
      uint2 pk;
      pk.x = *reinterpret_cast<unsigned*>(&lo);
      pk.y = *reinterpret_cast<unsigned*>(&hi);
      reinterpret_cast<uint2*>(a.ema_shadow)[i] = pk;
    }
  }
}

// p_ema <- p + beta (p_ema - p) on its own (EMA of a network whose step is taken elsewhere)
__global__ void __launch_bounds__(256) ema_kernel(const float* __restrict__ p, float* __restrict__ ema, long n, float beta) {
  const long nvec = n >> 2;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 a = reinterpret_cast<const float4*>(p)[i];
    float4 e = reinterpret_cast<float4*>(ema)[i];
    e.x = a.x + beta * (e.x - a.x);
    e.y = a.y + beta * (e.y - a.y);
    e.z = a.z + beta * (e.z - a.z);
    e.w = a.w + beta * (e.w - a.w);
    reinterpret_cast<float4*>(ema)[i] = e;
  }
}

// Per-step scalars of the two optimiser passes computed ON THE DEVICE from device-side counters, so a captured iteration
// (CUDA graph) advances its own step count: counters = {Adam steps of f_psi, Adam steps of G_theta, images seen}.
// hyper[0..3] = f_psi {lr, 1 - b1^t, sqrt(1 - b2^t), 0};  hyper[4..7] = G_theta {glr, ..., ..., ema_beta}
// (EMA beta as training/sid_training_loop.py:553-558: 0.5^(batch / max(min(halflife, nimg * rampup), 1e-8))).
__global__ void hyper_advance_kernel(float* hyper, long long* counters, float lr, float glr, float beta1, float beta2,
                                     double batch, double halflife_nimg, double rampup, int ema_on) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long tf = ++counters[0], tg = ++counters[1], nimg = counters[2];
  hyper[0] = lr;
  hyper[1] = (float)(1.0 - pow((double)beta1, (double)tf));
  hyper[2] = (float)sqrt(1.0 - pow((double)beta2, (double)tf));
  hyper[3] = 0.f;
  hyper[4] = glr;
  hyper[5] = (float)(1.0 - pow((double)beta1, (double)tg));
  hyper[6] = (float)sqrt(1.0 - pow((double)beta2, (double)tg));
  double half = halflife_nimg;
  if (rampup >= 0.0) half = fmin(half, (double)nimg * rampup);
  hyper[7] = ema_on ? (float)pow(0.5, batch / fmax(half, 1e-8)) : 0.f;
  counters[2] = nimg + (long long)batch;
}

}  // namespace sidlsg

using namespace sidlsg;

extern "C" int sidlsg_hyper_advance(float* hyper, long long* counters, float lr, float glr, float beta1, float beta2,
                                    double batch, double halflife_nimg, double rampup, int ema_on, void* stream) {
  hyper_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(hyper, counters, lr, glr, beta1, beta2, batch, halflife_nimg,
                                                           rampup, ema_on);
  return check_launch("hyper_advance");
}

// n must be a multiple of 4 (the flat bucket is padded).  step >= 1 is the Adam step count AFTER this update.
extern "C" int sidlsg_adam_step(float* p, const float* g, float* m, float* v, float* ema, void* shadow_bf16,
                                void* ema_shadow_bf16, long n,
                                float lr, float beta1, float beta2, float eps, int step, float grad_scale, float clip,
                                float ema_beta, float weight_decay, const float* hyper, void* stream) {
  if (n % 4) { set_error("adam_step: n=%ld not a multiple of 4", n); return SIDLSG_ERR_ARG; }
  if (step < 1) { set_error("adam_step: step must be >= 1"); return SIDLSG_ERR_ARG; }
  if (beta1 != 0.f && !m) { set_error("adam_step: beta1 != 0 needs the m buffer"); return SIDLSG_ERR_ARG; }
  if (n == 0) return SIDLSG_OK;
  AdamArgs a;
  a.p = p; a.g = g; a.m = (beta1 != 0.f) ? m : nullptr; a.v = v; a.ema = ema; a.shadow = (bf16*)shadow_bf16;
  a.ema_shadow = (bf16*)ema_shadow_bf16; a.n = n;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.grad_scale = grad_scale; a.clip = clip; a.ema_beta = ema_beta; a.weight_decay = weight_decay;
  a.hyper = hyper;
  long nvec = n / 4;
  int blocks = (int)min((long)148 * 8, (nvec + 255) / 256);
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("adam_step");
}

extern "C" int sidlsg_ema_update(const float* p, float* ema, long n, float beta, void* stream) {
  if (n % 4) { set_error("ema_update: n=%ld not a multiple of 4", n); return SIDLSG_ERR_ARG; }
  if (n == 0) return SIDLSG_OK;
  long nvec = n / 4;
  int blocks = (int)min((long)148 * 8, (nvec + 255) / 256);
  ema_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, ema, n, beta);
  return check_launch("ema_update");
}
