// Flash attention on tcgen05 / TMEM / TMA (bf16 mode): softmax(Q K^T / sqrt(d)) V per (batch, head) without
// materialising the scores.  Replaces diffusers' Attention (xformers memory_efficient_attention / SDPA,
// /root/reference/training/sid_sd_util.py:102-113) for the self- and cross-attention of every
// BasicTransformerBlock: (N, d) = (4096,40) (1024,80) (256,160) (64,160) for SD1.5, d = 64 for SD2.1, 77 text keys.
//
// One CTA = one 128-row query tile of one (batch, head); 192 threads:
//   warps 0-3  softmax: one query row per thread; S is read from TMEM, P = exp2(s*c - m*c) is written as bf16 into
//              a SWIZZLE_128B shared tile that feeds the second MMA; O stays in TMEM and is rescaled lazily (only
//              when the running max grew by more than 2^8), so no per-tile O read-back
//   warp 4     TMA producer: Q once, K/V tiles through a ring; head slices are cut with a 4-D tensor map
//              (d, heads, tokens, batch), so the zero padding of d=40 -> 48/64 comes from TMA out-of-bounds fill
//   warp 5     tcgen05.mma issuer + TMEM allocator: S = Q K^T (K-major x K-major), O += P V (V is the MN-major B)
// Two CTAs fit per SM for d <= 64 so one CTA's exponentials overlap the other's MMAs (the d=40 layers are
// MUFU-bound: 128x128 exps per tile vs ~400 tensor cycles).
#include "tc_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace sidlsg {

constexpr int AT_THREADS = 192;
constexpr int AT_BQ = 128;
constexpr int AT_BKV = 128;
constexpr int AT_CHUNK = 16384;   // [128 rows][64 bf16] SWIZZLE_128B

struct AttnParams {
  int B, H, N, M, d;
  int dchunks, dpad, kv_stages;
  float scale, scale_log2;
  bf16* o;
  float* lse;
  unsigned wait_hint;    // ns, helper-warp mbarrier waits (tc_wait_hint_ns)
  long long* trace;      // attn_fwd3_kernel<true> only: clock64 stamps of CTA (1,0,0), 16 slots per key tile
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
// packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2 issue once for two lanes of data)
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// exp2 of two values on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, max relative error 7.5e-5 - far
// below the bf16 rounding P receives): x = n + f with n = round(x) taken from the low mantissa bits of x + 1.5 * 2^23,
// 2^f ~ ((c3 f + c2) f + c1) f + c0 on [-0.5, 0.5], and n added into the exponent field.  The MUFU pipe (16 ex2 / clk / SM)
// is what bounds the softmax loops (scripts/trace_attn_fwd.py: exponentials 2170 of 3300 clk per tile pair); moving a
// fraction of the exponentials here lets both pipes work at once (FlashAttention-4's software exp2).
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x0 = fmaxf(x0, -125.f);                       // keeps the exponent arithmetic in range; 2^-125 is zero in bf16 anyway
  x1 = fmaxf(x1, -125.f);
  const uint64_t xc = pack2(x0, x1);
  const uint64_t t2 = fadd2(xc, pack2(12582912.f, 12582912.f));
  const uint64_t n2 = fadd2(t2, pack2(-12582912.f, -12582912.f));
  const uint64_t f2 = ffma2(n2, pack2(-1.f, -1.f), xc);
  uint64_t q2 = ffma2(f2, pack2(0.05517167f, 0.05517167f), pack2(0.24261113f, 0.24261113f));
  q2 = ffma2(q2, f2, pack2(0.69326097f, 0.69326097f));
  q2 = ffma2(q2, f2, pack2(0.99992806f, 0.99992806f));
  float q0, q1, t0, t1;
  unpack2(q2, q0, q1);
  unpack2(t2, t0, t1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}
// which of the 16 score pairs of a 32-column chunk take the polynomial: NPOLY of 16, spread evenly
template <int NPOLY>
__device__ __forceinline__ constexpr bool poly_pair(int i2) {
  return ((i2 + 1) * NPOLY) / 16 != (i2 * NPOLY) / 16;
}

template <int TCOLS>
__global__ void __launch_bounds__(AT_THREADS, TCOLS == 256 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const int ST = p.kv_stages;
  const uint32_t q_smem = smem_base;
  const uint32_t k_smem = q_smem + p.dchunks * AT_CHUNK;
  const uint32_t v_smem = k_smem + ST * p.dchunks * AT_CHUNK;
  const uint32_t p_smem = v_smem + ST * p.dchunks * AT_CHUNK;
  const uint32_t bar_base = p_smem + 2 * AT_CHUNK;
  // barriers: q_full, s_full, p_full, o_full, k_full[2], v_full[2], kv_empty[2], then the TMEM slot
  const uint32_t q_full = bar_base, s_full = bar_base + 8, p_full = bar_base + 16, o_full = bar_base + 24;
  auto k_full = [&](int s) { return bar_base + 32 + 8u * s; };
  auto v_full = [&](int s) { return bar_base + 48 + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 64 + 8u * s; };
  const uint32_t tmem_slot = bar_base + 80;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
  const int T = (p.M + AT_BKV - 1) / AT_BKV;

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) { printf("sidlsg attention: dynamic smem not 1024-aligned\n"); __trap(); }
    mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(k_full(s), 1); mbar_init(v_full(s), 1); mbar_init(kv_empty(s), 1); }
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == 5) tmem_alloc<TCOLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t s_tmem = tmem_base, o_tmem = tmem_base + 128;
  const uint32_t tile_bytes = p.dchunks * AT_CHUNK;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (role_leader()) {
      mbar_expect_tx(q_full, tile_bytes);
      for (int c = 0; c < p.dchunks; ++c) tma_load_4d(&tmQ, q_full, q_smem + c * AT_CHUNK, c * 64, h, q0, b);
      for (int j = 0; j < T; ++j) {
        const int st = j % ST;
        const uint32_t ph = (j / ST) & 1;
        mbar_wait_h(p.wait_hint, kv_empty(st), ph ^ 1);
        mbar_expect_tx(k_full(st), tile_bytes);
        for (int c = 0; c < p.dchunks; ++c)
          tma_load_4d(&tmK, k_full(st), k_smem + (st * p.dchunks + c) * AT_CHUNK, c * 64, h, j * AT_BKV, b);
        mbar_expect_tx(v_full(st), tile_bytes);
        for (int c = 0; c < p.dchunks; ++c)
          tma_load_4d(&tmV, v_full(st), v_smem + (st * p.dchunks + c) * AT_CHUNK, c * 64, h, j * AT_BKV, b);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (role_leader()) {
      const int dsteps = p.dpad >> 4;
      auto issue_qk = [&](int j) {
        const int st = j % ST;
        const int n_valid = min(AT_BKV, p.M - j * AT_BKV);
        const uint32_t idesc = make_idesc((n_valid + 15) & ~15, 0, 0);
        const uint32_t kb = k_smem + st * tile_bytes;
        for (int s = 0; s < dsteps; ++s) {
          const uint32_t off = (s >> 2) * AT_CHUNK + (s & 3) * 32;
          tc_mma_bf16(s_tmem, make_desc(q_smem + off, 1024, 0), make_desc(kb + off, 1024, 0), idesc, s > 0);
        }
      };
      mbar_wait_h(p.wait_hint, q_full, 0);
      mbar_wait_h(p.wait_hint, k_full(0), 0);
      tc_fence_after();
      issue_qk(0);
      tc_commit(s_full);
      const uint32_t idesc_pv = make_idesc(p.dpad, 0, 1);
      for (int j = 0; j < T; ++j) {
        const int st = j % ST;
        const uint32_t ph = (j / ST) & 1;
        const int n_valid = min(AT_BKV, p.M - j * AT_BKV);
        mbar_wait_h(p.wait_hint, p_full, j & 1);
        mbar_wait_h(p.wait_hint, v_full(st), ph);
        tc_fence_after();
        const uint32_t vb = v_smem + st * tile_bytes;
        const int ksteps = (n_valid + 15) >> 4;
        for (int s = 0; s < ksteps; ++s) {
          tc_mma_bf16(o_tmem, make_desc(p_smem + (s >> 2) * AT_CHUNK + (s & 3) * 32, 1024, 0),
                      make_desc(vb + s * 2048, 1024, AT_CHUNK), idesc_pv, (j > 0 || s > 0) ? 1u : 0u);
        }
        tc_commit(kv_empty(st));
        tc_commit(o_full);
        if (j + 1 < T) {
          const int st2 = (j + 1) % ST;
          mbar_wait_h(p.wait_hint, k_full(st2), ((j + 1) / ST) & 1);
          tc_fence_after();
          issue_qk(j + 1);
          tc_commit(s_full);
        }
      }
    }
  } else {
    // ===================== softmax / output (warps 0-3) =====================
    // Instruction budget per score: FFMA2 (1/2) + MUFU.EX2 (1) + FADD2 (1/2) + FMNMX3 (1/2) + F2FP (1/2) = 3 issue
    // slots, so the loop is bound by the 16/clk/SM MUFU pipe.  Full tiles after the first run ONE pass over S with
    // the running max (speculative); the tile is redone only if its max exceeds the running max by more than 2^8.
    const int row = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float c = p.scale_log2;
    const uint64_t c2 = pack2(c, c);
    const uint32_t p_row = p_smem + row * 128;
    const int rx = row & 7;
    float m_used = -INFINITY, l_run = 0.f;

    auto tile_max = [&](int n_valid) {
      float mx = -INFINITY;
      if (n_valid == AT_BKV) {
#pragma unroll
        for (int cc = 0; cc < AT_BKV; cc += 32) {
          uint32_t r[32];
          tmem_ld32_nowait(s_tmem + lane_off + cc, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = fmaxf(fmaxf(mx, __uint_as_float(r[i])), __uint_as_float(r[i + 1]));
        }
      } else {
        for (int cc = 0; cc < n_valid; cc += 32) {
          uint32_t r[32];
          tmem_ld32_nowait(s_tmem + lane_off + cc, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) if (cc + i < n_valid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      return mx;
    };
    // P = exp2(s*c - m*c) -> bf16 smem tile; returns the row sum, tracks the raw row max in mx
    auto tile_exp = [&](int n_valid, float m, float& mx) {
      const float nmc = -m * c;
      const uint64_t nmc2 = pack2(nmc, nmc);
      uint64_t ls2 = pack2(0.f, 0.f);
      float ls = 0.f;
      if (n_valid == AT_BKV) {
#pragma unroll
        for (int cc = 0; cc < AT_BKV; cc += 32) {
          uint32_t r[32], pk[16];
          tmem_ld32_nowait(s_tmem + lane_off + cc, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float x0 = __uint_as_float(r[i]), x1 = __uint_as_float(r[i + 1]);
            mx = fmaxf(fmaxf(mx, x0), x1);
            float t0, t1;
            unpack2(ffma2(pack2(x0, x1), c2, nmc2), t0, t1);
            const float p0 = ex2f(t0), p1 = ex2f(t1);
            ls2 = fadd2(ls2, pack2(p0, p1));
            pk[i >> 1] = pack_bf16(p0, p1);
          }
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            const int u = ((cc & 63) >> 3) + qd;
            st_shared_v4(p_row + (cc >> 6) * AT_CHUNK + ((u ^ rx) << 4), pk[4 * qd], pk[4 * qd + 1], pk[4 * qd + 2],
                         pk[4 * qd + 3]);
          }
        }
        float a0, a1;
        unpack2(ls2, a0, a1);
        ls = a0 + a1;
      } else {
        for (int cc = 0; cc < n_valid; cc += 32) {
          uint32_t r[32], pk[16];
          tmem_ld32_nowait(s_tmem + lane_off + cc, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const bool v0 = cc + i < n_valid, v1 = cc + i + 1 < n_valid;
            const float x0 = __uint_as_float(r[i]), x1 = __uint_as_float(r[i + 1]);
            if (v0) mx = fmaxf(mx, x0);
            if (v1) mx = fmaxf(mx, x1);
            const float p0 = v0 ? ex2f(fmaf(x0, c, nmc)) : 0.f;
            const float p1 = v1 ? ex2f(fmaf(x1, c, nmc)) : 0.f;
            ls += p0 + p1;
            pk[i >> 1] = pack_bf16(p0, p1);
          }
#pragma unroll
          for (int qd = 0; qd < 4; ++qd)
            st_shared_v4(p_smem + sw128_offset(row, cc + 8 * qd, AT_CHUNK), pk[4 * qd], pk[4 * qd + 1], pk[4 * qd + 2],
                         pk[4 * qd + 3]);
        }
      }
      return ls;
    };
    // O *= f (per-lane factor) in TMEM; all 32 lanes call it together
    auto rescale_o = [&](float f) {
      for (int cc = 0; cc < p.dpad; cc += 16) {
        uint32_t r[16];
        tmem_ld16_nowait(o_tmem + lane_off + cc, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
        tmem_st16(o_tmem + lane_off + cc, r);
      }
      tmem_wait_st();
    };

    for (int j = 0; j < T; ++j) {
      const int n_valid = min(AT_BKV, p.M - j * AT_BKV);
      const bool spec = j > 0 && n_valid == AT_BKV;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (!spec) {
        const float m_new = fmaxf(m_used, tile_max(n_valid));
        const bool need = (m_new - m_used) * c > 8.f;
        if (j > 0) {
          mbar_wait(o_full, (j - 1) & 1);   // P V of the previous tile retired: O and the P tile are ours again
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
            const float f = need ? ex2f((m_used - m_new) * c) : 1.f;
            rescale_o(f);
            l_run *= f;
          }
        }
        if (need) m_used = m_new;
      } else {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
      }
      float mx = -INFINITY;
      float l_tile = tile_exp(n_valid, m_used, mx);
      if (spec) {
        const float m_new = fmaxf(m_used, mx);
        const bool need = (m_new - m_used) * c > 8.f;
        if (__any_sync(0xffffffffu, need)) {       // rare: this tile raised the max by more than 2^8 -> redo it
          const float f = need ? ex2f((m_used - m_new) * c) : 1.f;
          rescale_o(f);
          l_run *= f;
          if (need) m_used = m_new;
          l_tile = tile_exp(n_valid, m_used, mx);
        }
      }
      l_run += l_tile;
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_full, (T - 1) & 1);
    tc_fence_after();
    const int q = q0 + row;
    const float inv = 1.f / l_run;
    const int C = p.H * p.d;
    bf16* orow = p.o + ((long)b * p.N + q) * C + h * p.d;
    for (int cc = 0; cc < p.dpad; cc += 16) {
      uint32_t r[16];
      tmem_ld16_nowait(o_tmem + lane_off + cc, r);
      tmem_wait_ld();
      if (q < p.N) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (cc + 8 * hh < p.d) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r[8 * hh + 0]) * inv, __uint_as_float(r[8 * hh + 1]) * inv);
            o.y = pack_bf16(__uint_as_float(r[8 * hh + 2]) * inv, __uint_as_float(r[8 * hh + 3]) * inv);
            o.z = pack_bf16(__uint_as_float(r[8 * hh + 4]) * inv, __uint_as_float(r[8 * hh + 5]) * inv);
            o.w = pack_bf16(__uint_as_float(r[8 * hh + 6]) * inv, __uint_as_float(r[8 * hh + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + cc + 8 * hh) = o;
          }
        }
      }
    }
    if (q < p.N && p.lse) p.lse[((long)b * p.H + h) * p.N + q] = m_used * p.scale + logf(l_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward, two query tiles per CTA (d <= 64: the N = 4096, d = 40 layers of SD1.5 that hold 3/4 of all attention time, and
// every SD2.1 layer).  Measured on B200 with elect.sync issue (profiles/r02_ubench_mma_issue_b200.txt): per 128 x 128
// tile the tensor pipe needs 3 x 109 clk for S = Q K^T and 8 x 55 clk for O += P V when P is the TMEM A operand (8 x 81
// from shared memory), the MUFU pipe 1037 clk for the exponentials - the single-tile kernel above runs them strictly one
// after the other (2150 clk per tile).  Here ONE CTA per SM owns TWO 128-row query tiles A, B with their own S, O and P
// regions in TMEM (S 2 x 128 | O 2 x 64 | P 2 x 64 columns = 512) and their own 4 softmax warps; the single MMA warp
// alternates  P_A V -> Q_A K'^T -> P_B V -> Q_B K'^T,  so the exponentials of one tile run under the MMAs of the other
// (FlashAttention-4's ping-pong schedule).  P never touches shared memory (tcgen05.st, bf16 pairs per 32-bit column),
// K and V travel through separate TMA rings (K_{j+1} is needed while V_j is still in use).
//   barriers: q_full[2], s_full[2] (MMA -> softmax), p_full[2] (softmax -> MMA, 4 warp arrivals), o_full[2] (P_X V retired),
//             k_full / k_empty [KST], v_full / v_empty [VST]
constexpr int A3_THREADS = 320;
constexpr int A3_KST = 4, A3_VST = 4;

template <bool TRACE, int NPOLY>
__global__ void __launch_bounds__(A3_THREADS, 1)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t q_smem = smem_base;                          // 2 x 16 KB
  const uint32_t k_smem = q_smem + 2 * AT_CHUNK;              // A3_KST x 16 KB
  const uint32_t v_smem = k_smem + A3_KST * AT_CHUNK;         // A3_VST x 16 KB
  const uint32_t bar_base = v_smem + A3_VST * AT_CHUNK;
  auto q_full = [&](int x) { return bar_base + 8u * x; };
  auto s_full = [&](int x) { return bar_base + 16 + 8u * x; };
  auto p_full = [&](int x) { return bar_base + 32 + 8u * x; };
  auto o_full = [&](int x) { return bar_base + 48 + 8u * x; };
  auto k_full = [&](int s) { return bar_base + 64 + 8u * s; };
  auto k_empty = [&](int s) { return bar_base + 96 + 8u * s; };
  auto v_full = [&](int s) { return bar_base + 128 + 8u * s; };
  auto v_empty = [&](int s) { return bar_base + 160 + 8u * s; };
  const uint32_t tmem_slot = bar_base + 192;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * AT_BQ), h = blockIdx.y, b = blockIdx.z;
  const int T = (p.M + AT_BKV - 1) / AT_BKV;
  const int ntile = (q0 + AT_BQ < p.N) ? 2 : 1;               // the second query tile may lie entirely past N

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) { printf("sidlsg attention: dynamic smem not 1024-aligned\n"); __trap(); }
    for (int x = 0; x < 2; ++x) { mbar_init(q_full(x), 1); mbar_init(s_full(x), 1); mbar_init(p_full(x), 4); mbar_init(o_full(x), 1); }
    for (int s2 = 0; s2 < 4; ++s2) { mbar_init(k_full(s2), 1); mbar_init(k_empty(s2), 1); mbar_init(v_full(s2), 1); mbar_init(v_empty(s2), 1); }
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  auto s_tmem = [&](int x) { return tmem_base + 128u * x; };
  auto o_tmem = [&](int x) { return tmem_base + 256u + 64u * x; };
  auto p_tmem = [&](int x) { return tmem_base + 384u + 64u * x; };

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (role_leader()) {
      for (int x = 0; x < ntile; ++x) {
        mbar_expect_tx(q_full(x), AT_CHUNK);
        tma_load_4d(&tmQ, q_full(x), q_smem + x * AT_CHUNK, 0, h, q0 + x * AT_BQ, b);
      }
#pragma unroll 1
      for (int j = 0; j < T; ++j) {
        const int ks = j % A3_KST, vs = j % A3_VST;
        mbar_wait_h(p.wait_hint, k_empty(ks), ((j / A3_KST) & 1) ^ 1);
        mbar_expect_tx(k_full(ks), AT_CHUNK);
        tma_load_4d(&tmK, k_full(ks), k_smem + ks * AT_CHUNK, 0, h, j * AT_BKV, b);
        mbar_wait_h(p.wait_hint, v_empty(vs), ((j / A3_VST) & 1) ^ 1);
        mbar_expect_tx(v_full(vs), AT_CHUNK);
        tma_load_4d(&tmV, v_full(vs), v_smem + vs * AT_CHUNK, 0, h, j * AT_BKV, b);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    // One thread, one long dependent instruction stream: at ~3,000 clk per key tile it WAS the period of the kernel
    // (phase trace r02: every wait on an already-complete barrier cost 150-250 clk - two constant-bank loads, a uniform
    // compare and the try_wait in series - and every MMA ~55 clk of dependent uniform-datapath descriptor arithmetic,
    // while the tensor pipe needs ~1,700 clk).  So: the suspend hint lives in a register, descriptors are base + immediate
    // (independent adds, loops fully unrolled with predicates), and nothing but the fence sits between seeing P_X(j)
    // and issuing S_X(j + 1).
    if (role_leader()) {
      uint32_t hint;
      asm volatile("mov.u32 %0, %1;" : "=r"(hint) : "r"(p.wait_hint));
      auto wait = [&](uint32_t bar, uint32_t parity) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok) mbar_wait_h(hint, bar, parity);
      };
      const int dsteps = p.dpad >> 4;
      const uint32_t idesc_pv = make_idesc(p.dpad, 0, 1);
      const uint32_t idesc_full = make_idesc(AT_BKV, 0, 0);
      const int n_last = p.M - (T - 1) * AT_BKV;
      const uint32_t idesc_last = make_idesc((n_last + 15) & ~15, 0, 0);
      const int ksteps_last = (n_last + 15) >> 4;
      const uint64_t qd0 = make_desc(q_smem, 1024, 0), qd1 = make_desc(q_smem + AT_CHUNK, 1024, 0);
      const uint64_t kd0 = make_desc(k_smem, 1024, 0), vd0 = make_desc(v_smem, 1024, AT_CHUNK);
      const uint32_t s0 = s_tmem(0), s1 = s_tmem(1), o0 = o_tmem(0), o1 = o_tmem(1), p0 = p_tmem(0), p1 = p_tmem(1);
      const bool two = ntile == 2;
      // S_X = Q_X K^T: <= 4 MMAs of K = 16 (descriptor + 2 = 32 bytes along the 128-byte swizzled row)
      auto issue_qk = [&](uint32_t st, uint64_t qd, uint64_t kd, uint32_t idesc, uint32_t bar) {
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2)
          if (s2 < dsteps) tc_mma_bf16(st, qd + 2 * s2, kd + 2 * s2, idesc, s2 > 0);
        tc_commit(bar);
      };
      // O_X += P_X V: <= 8 MMAs of K = 16 keys (P from TMEM, 8 columns each; V rows 16 s2 .. = + 2,048 bytes)
      auto issue_pv = [&](uint32_t ot, uint32_t pt, uint64_t vd, int ksteps, uint32_t acc0, uint32_t bar) {
#pragma unroll
        for (int s2 = 0; s2 < 8; ++s2)
          if (s2 < ksteps) tc_mma_bf16_ta(ot, pt + 8 * s2, vd + 128 * s2, idesc_pv, s2 > 0 ? 1u : acc0);
        tc_commit(bar);
      };
      wait(q_full(0), 0);
      if (two) wait(q_full(1), 0);
      wait(k_full(0), 0);
      tc_fence_after();
      {
        const uint32_t id0 = T == 1 ? idesc_last : idesc_full;
        issue_qk(s0, qd0, kd0, id0, s_full(0));
        if (two) issue_qk(s1, qd1, kd0, id0, s_full(1));
      }
      tc_commit(k_empty(0));
#pragma unroll 1
      for (int j = 0; j < T; ++j) {
        const int vs = j & (A3_VST - 1), ks = (j + 1) & (A3_KST - 1);
        const bool more = j + 1 < T;
        const int ksteps = more ? AT_BKV / 16 : ksteps_last;
        const uint32_t idq = (j + 2 < T) ? idesc_full : idesc_last;
        const uint64_t vd = vd0 + (uint64_t)(vs * (AT_CHUNK >> 4)), kd = kd0 + (uint64_t)(ks * (AT_CHUNK >> 4));
        const uint32_t acc0 = j > 0 ? 1u : 0u, par = j & 1;
        const bool tr = TRACE && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && j < 32;
        if (tr) p.trace[j * 16 + 14] = clock64();
        if (more) wait(k_full(ks), ((j + 1) >> 2) & 1);
        if (tr) p.trace[j * 16 + 8] = clock64();
        wait(p_full(0), par);
        if (tr) p.trace[j * 16 + 9] = clock64();
        tc_fence_after();
        // S_X(j) has been consumed (P_X(j) exists): the NEXT tile's scores go first, so the softmax warps get them
        // right away instead of after the 8 MMAs of P V (their exponentials of tile j + 1 then run under P_X(j) V_j;
        // they wait on o_full before touching P_X or O_X again)
        if (more) issue_qk(s0, qd0, kd, idq, s_full(0));
        if (tr) p.trace[j * 16 + 7] = clock64();
        wait(v_full(vs), (j >> 2) & 1);
        issue_pv(o0, p0, vd, ksteps, acc0, o_full(0));
        if (tr) p.trace[j * 16 + 10] = clock64();
        if (two) {
          if (tr) p.trace[j * 16 + 11] = clock64();
          wait(p_full(1), par);
          if (tr) p.trace[j * 16 + 12] = clock64();
          tc_fence_after();
          if (more) issue_qk(s1, qd1, kd, idq, s_full(1));
          issue_pv(o1, p1, vd, ksteps, acc0, o_full(1));
          if (tr) p.trace[j * 16 + 13] = clock64();
        }
        tc_commit(v_empty(vs));
        if (more) tc_commit(k_empty(ks));
      }
    }
  } else if ((warp >> 2) < ntile) {
    // ===================== softmax / output: warps 0-3 own tile A, warps 4-7 tile B =====================
    const int x = warp >> 2, wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;       // a warp reaches TMEM lanes 32 (warp % 4) .. + 31
    const uint32_t st = s_tmem(x) + lane_off, ot = o_tmem(x) + lane_off, pt = p_tmem(x) + lane_off;
    const float c = p.scale_log2;
    const uint64_t c2 = pack2(c, c);
    float m_used = -INFINITY, l_run = 0.f;

    auto tile_max = [&](int n_valid) {
      float mx = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < n_valid; cc += 32) {
        uint32_t r[32];
        tmem_ld32_nowait(st + cc, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) if (cc + i < n_valid) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      return mx;
    };
    // P = exp2(s*c - m*c) -> bf16 pairs in TMEM; returns the row sum, tracks the raw row max in mx
    bool pv_pending = false;     // P_X V_{j-1} may still be reading P_X / writing O_X
    auto wait_pv = [&](int j) {
      if (pv_pending) {
        mbar_wait(o_full(x), (j - 1) & 1);
        tc_fence_after();
        pv_pending = false;
      }
    };
    auto tile_exp = [&](int n_valid, float m, float& mx, int j) {
      const float nmc = -m * c;
      const uint64_t nmc2 = pack2(nmc, nmc);
      float ls = 0.f;
      if (n_valid == AT_BKV) {
        uint64_t ls2 = pack2(0.f, 0.f);
        // software-pipelined over the four 32-column chunks: the TMEM load of chunk k + 1 is in flight while chunk k is
        // exponentiated (two register buffers; a warp has only one partner warp on its scheduler to hide that latency)
        uint32_t rbuf[2][32];
        tmem_ld32_nowait(st, rbuf[0]);
#pragma unroll
        for (int cc = 0; cc < AT_BKV; cc += 32) {
          uint32_t* r = rbuf[(cc >> 5) & 1];
          uint32_t pk[16];
          tmem_wait_ld_dep32(r);
          if (cc + 32 < AT_BKV) tmem_ld32_nowait(st + cc + 32, rbuf[((cc >> 5) + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float x0 = __uint_as_float(r[i]), x1 = __uint_as_float(r[i + 1]);
            mx = fmaxf(fmaxf(mx, x0), x1);
            const uint64_t t2 = ffma2(pack2(x0, x1), c2, nmc2);
            float p0, p1;
            if (poly_pair<NPOLY>(i >> 1)) {
              exp2_poly2(t2, p0, p1);
            } else {
              float t0, t1;
              unpack2(t2, t0, t1);
              p0 = ex2f(t0);
              p1 = ex2f(t1);
            }
            ls2 = fadd2(ls2, pack2(p0, p1));
            pk[i >> 1] = pack_bf16(p0, p1);
          }
          wait_pv(j);
          tmem_st16(pt + (cc >> 1), pk);
        }
        float a0, a1;
        unpack2(ls2, a0, a1);
        ls = a0 + a1;
      } else {
#pragma unroll 1
        for (int cc = 0; cc < n_valid; cc += 32) {
          uint32_t r[32], pk[16];
          tmem_ld32_nowait(st + cc, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const bool v0 = cc + i < n_valid, v1 = cc + i + 1 < n_valid;
            const float x0 = __uint_as_float(r[i]), x1 = __uint_as_float(r[i + 1]);
            if (v0) mx = fmaxf(mx, x0);
            if (v1) mx = fmaxf(mx, x1);
            const float p0 = v0 ? ex2f(fmaf(x0, c, nmc)) : 0.f;
            const float p1 = v1 ? ex2f(fmaf(x1, c, nmc)) : 0.f;
            ls += p0 + p1;
            pk[i >> 1] = pack_bf16(p0, p1);
          }
          wait_pv(j);
          tmem_st16(pt + (cc >> 1), pk);
        }
      }
      tmem_wait_st();
      return ls;
    };
    auto rescale_o = [&](float f) {
#pragma unroll 1
      for (int cc = 0; cc < p.dpad; cc += 16) {
        uint32_t r[16];
        tmem_ld16_nowait(ot + cc, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
        tmem_st16(ot + cc, r);
      }
      tmem_wait_st();
    };

    for (int j = 0; j < T; ++j) {
      const int n_valid = min(AT_BKV, p.M - j * AT_BKV);
      const bool spec = j > 0 && n_valid == AT_BKV;
      // S_X(j) is issued BEFORE P_X V_{j-1}: that product may still be running when the scores arrive, so O_X and the
      // P_X region are touched only after o_full (wait_pv below; by then most of the first 32 columns are done)
      const bool tr = TRACE && (warp & 3) == 0 && lane == 0 && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && j < 32;
      if (tr) p.trace[j * 16 + 4 * x] = clock64();
      mbar_wait(s_full(x), j & 1);
      if (tr) p.trace[j * 16 + 4 * x + 1] = clock64();
      tc_fence_after();
      pv_pending = j > 0;
      if (!spec) {
        const float m_new = fmaxf(m_used, tile_max(n_valid));
        const bool need = (m_new - m_used) * c > 8.f;
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          const float f = need ? ex2f((m_used - m_new) * c) : 1.f;
          wait_pv(j);
          rescale_o(f);
          l_run *= f;
        }
        if (need) m_used = m_new;
      }
      float l_tile;
#pragma unroll 1
      for (int pass = 0;; ++pass) {                // (a loop so that the exponential code exists once: I-cache)
        float mx = -INFINITY;
        l_tile = tile_exp(n_valid, m_used, mx, j);
        if (!spec || pass) break;
        const float m_new = fmaxf(m_used, mx);
        const bool need = (m_new - m_used) * c > 8.f;
        if (!__any_sync(0xffffffffu, need)) break;
        // rare: this tile raised the max by more than 2^8 -> rescale and redo it
        const float f = need ? ex2f((m_used - m_new) * c) : 1.f;
        rescale_o(f);
        l_run *= f;
        if (need) m_used = m_new;
      }
      l_run += l_tile;
      if (tr) p.trace[j * 16 + 4 * x + 2] = clock64();
      if (TRACE && lane == 0 && (warp == 1 || warp == 3) && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && j < 32)
        p.trace[j * 16 + (warp == 1 ? 3 : 15)] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(x));
    }
    mbar_wait(o_full(x), (T - 1) & 1);
    tc_fence_after();
    const int q = q0 + x * AT_BQ + row;
    const float inv = 1.f / l_run;
    const int C = p.H * p.d;
    bf16* orow = p.o + ((long)b * p.N + q) * C + h * p.d;
    for (int cc = 0; cc < p.dpad; cc += 16) {
      uint32_t r[16];
      tmem_ld16_nowait(ot + cc, r);
      tmem_wait_ld();
      if (q < p.N) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (cc + 8 * hh < p.d) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r[8 * hh + 0]) * inv, __uint_as_float(r[8 * hh + 1]) * inv);
            o.y = pack_bf16(__uint_as_float(r[8 * hh + 2]) * inv, __uint_as_float(r[8 * hh + 3]) * inv);
            o.z = pack_bf16(__uint_as_float(r[8 * hh + 4]) * inv, __uint_as_float(r[8 * hh + 5]) * inv);
            o.w = pack_bf16(__uint_as_float(r[8 * hh + 6]) * inv, __uint_as_float(r[8 * hh + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + cc + 8 * hh) = o;
          }
        }
      }
    }
    if (q < p.N && p.lse) p.lse[((long)b * p.H + h) * p.N + q] = m_used * p.scale + logf(l_run);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// head-sliced view of a [B, len, >= H*d] bf16 tensor with row stride ld elements (ld = H*d when dense; 3*H*d for
// the q/k/v thirds of a packed projection output): dims (d, H, len, B)
static bool make_head_map(CUtensorMap* m, const void* base, int d, int H, int len, int B, long ld, int rows = 128) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)H, (uint64_t)len, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)len * ld * 2};
  uint32_t box[4] = {64, 1, (uint32_t)rows, 1};
  return make_map(m, base, 4, dims, strides, box);
}

static bool g_attn_attr_done = false;

}  // namespace sidlsg

using namespace sidlsg;

// q [B,N,H*d], k/v [B,M,H*d] bf16 with row strides ldq/ldk/ldv elements (batch stride = rows * ld; multiples of 8);
// o [B,N,H*d] bf16 dense; lse [B,H,N] fp32 (natural log, may be null).
// Returns SIDLSG_ERR_UNSUPPORTED when the tensor-core path cannot take the shape (caller uses the fp32-exact path).
static int attention_fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, int B, int N,
                              int M, int H, int d, long ldq, long ldk, long ldv, long long* trace, void* stream) {
  if (!tc_enabled()) { set_error("attention_fwd: tcgen05 path unavailable on this device"); return SIDLSG_ERR_UNSUPPORTED; }
  if (d % 8 || d < 16 || d > 192 || B <= 0 || N <= 0 || M <= 0 || H <= 0 || H > 65535 || B > 65535) {
    set_error("attention_fwd: unsupported shape B=%d N=%d M=%d H=%d d=%d", B, N, M, H, d);
    return SIDLSG_ERR_UNSUPPORTED;
  }
  if (ldq < (long)H * d || ldk < (long)H * d || ldv < (long)H * d || ((ldq | ldk | ldv) & 7)) {
    set_error("attention_fwd: row strides must be >= H*d and multiples of 8 elements");
    return SIDLSG_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.N = N; p.M = M; p.d = d;
  p.dchunks = (d + 63) / 64;
  p.dpad = (d + 15) & ~15;
  p.kv_stages = p.dchunks <= 2 ? 2 : 1;
  p.scale = 1.f / sqrtf((float)d);
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.o = (bf16*)o; p.lse = lse;
  p.wait_hint = tc_wait_hint_ns();
  p.trace = trace;
  CUtensorMap tq, tk, tv;
  if (!make_head_map(&tq, q, d, H, N, B, ldq) || !make_head_map(&tk, k, d, H, M, B, ldk) ||
      !make_head_map(&tv, v, d, H, M, B, ldv))
    return SIDLSG_ERR_CUDA;
  // d <= 64: two query tiles per CTA (attn_fwd3_kernel); SIDLSG_ATTN_FWD3=0 keeps the single-tile kernel (A/B switch)
  static int fwd3 = -1;
  if (fwd3 < 0) { const char* e = getenv("SIDLSG_ATTN_FWD3"); fwd3 = (e && e[0] == '0') ? 0 : 1; }
  if (fwd3 && p.dpad <= 64) {
    static bool attr3 = false;
    if (!attr3) {
      cudaFuncSetAttribute(attn_fwd3_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(attn_fwd3_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(attn_fwd3_kernel<false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(attn_fwd3_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(attn_fwd3_kernel<true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      attr3 = true;
    }
    const int smem3 = (2 + A3_KST + A3_VST) * AT_CHUNK + 256;
    dim3 grid3((N + 2 * AT_BQ - 1) / (2 * AT_BQ), H, B);
    // SIDLSG_ATTN_POLY = how many of every 16 exponentials run as a polynomial on the FMA pipe (0, 4, 6, 8; default 6)
    static int npoly = -1;
    if (npoly < 0) { const char* e = getenv("SIDLSG_ATTN_POLY"); npoly = e ? atoi(e) : 6; }
    if (trace) attn_fwd3_kernel<true, 6><<<grid3, A3_THREADS, smem3, st>>>(tq, tk, tv, p);
    else if (npoly <= 0) attn_fwd3_kernel<false, 0><<<grid3, A3_THREADS, smem3, st>>>(tq, tk, tv, p);
    else if (npoly <= 4) attn_fwd3_kernel<false, 4><<<grid3, A3_THREADS, smem3, st>>>(tq, tk, tv, p);
    else if (npoly <= 6) attn_fwd3_kernel<false, 6><<<grid3, A3_THREADS, smem3, st>>>(tq, tk, tv, p);
    else attn_fwd3_kernel<false, 8><<<grid3, A3_THREADS, smem3, st>>>(tq, tk, tv, p);
    return check_launch("attention_fwd");
  }
  const int smem = (p.dchunks * (1 + 2 * p.kv_stages) + 2) * AT_CHUNK + 256;
  if (!g_attn_attr_done) {
    cudaFuncSetAttribute(attn_fwd_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(attn_fwd_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    g_attn_attr_done = true;
  }
  dim3 grid((N + AT_BQ - 1) / AT_BQ, H, B);
  if (p.dpad <= 128) attn_fwd_kernel<256><<<grid, AT_THREADS, smem, st>>>(tq, tk, tv, p);
  else attn_fwd_kernel<512><<<grid, AT_THREADS, smem, st>>>(tq, tk, tv, p);
  return check_launch("attention_fwd");
}

extern "C" int sidlsg_attention_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int N,
                                    int M, int H, int d, long ldq, long ldk, long ldv, void* stream) {
  return attention_fwd_impl(q, k, v, o, lse, B, N, M, H, d, ldq, ldk, ldv, nullptr, stream);
}

// Development aid: forward (d <= 64 kernel) with in-kernel clock64 stamps of CTA (1,0,0); not used by the product path.
extern "C" int sidlsg_debug_attention_fwd_trace(const void* q, const void* k, const void* v, void* o, float* lse, int B,
                                                int N, int M, int H, int d, long ldq, long ldk, long ldv, void* trace,
                                                void* stream) {
  return attention_fwd_impl(q, k, v, o, lse, B, N, M, H, d, ldq, ldk, ldv, (long long*)trace, stream);
}

// =====================================================================================================================
// Backward: dQ, dK, dV with the scores recomputed on the tensor cores (FlashAttention-2 schedule).
// One CTA owns a 128-row K/V tile of one (batch, head) and walks the query tiles.  Everything is computed in the
// TRANSPOSED orientation so every operand is a plain TMA tile and no transpose is ever materialised:
//   S^T  = K Q^T            (A = K tile, B = Q tile, both K-major)              -> TMEM
//   dP^T = V dO^T           (A = V tile, B = dO tile, both K-major)             -> TMEM
//   P^T  = exp2(S^T c - lse), dS^T = P^T o (dP^T - delta) * scale               (8 compute warps, one kv row per thread,
//                                                                                 written as bf16 SWIZZLE_128B tiles)
//   dV  += P^T dO           (A = P^T tile K-major, B = the SAME dO tile read MN-major)
//   dK  += dS^T Q           (A = dS^T tile K-major, B = the SAME Q tile read MN-major)
//   dQ_i = dS K             (A = the dS^T tile read MN-major, B = K tile read MN-major) -> fp32 red.add into dq_acc
// dK/dV accumulate in TMEM over the whole query loop; dQ partials are reduced across K/V tiles with red.global.add.
// 320 threads: warps 0-7 compute (two warpgroups split the 128 query columns), warp 8 TMA, warp 9 MMA + TMEM alloc.
namespace sidlsg {

constexpr int AB_THREADS = 320;

struct AttnBwdParams {
  int B, H, N, M, d;
  int dchunks, dpad, q_stages;
  float scale, scale_log2;
  const float* lse;      // [B,H,N]
  const float* delta;    // [B,H,N] = sum_c O dO
  float* dq_acc;         // [B,N,H*d] fp32, zero-initialised by the caller (written through tmDQ)
  int stage_alias;       // dQ staging aliases the P^T/dS^T tiles (d > 64)
  unsigned wait_hint;    // ns, helper-warp mbarrier waits
  bf16* dk;
  bf16* dv;
  long lddk, lddv;       // row strides (elements) of dk / dv
  long long* trace;      // TRACE instantiation only: clock64 stamps of CTA (1,0,0), 16 slots per query tile
};

template <bool TRACE>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const int QST = p.q_stages;
  const uint32_t tile_bytes = p.dchunks * AT_CHUNK;
  const uint32_t k_smem = smem_base;
  const uint32_t v_smem = k_smem + tile_bytes;
  const uint32_t q_smem = v_smem + tile_bytes;
  const uint32_t do_smem = q_smem + QST * tile_bytes;
  const uint32_t p_smem = do_smem + QST * tile_bytes;     // P^T  [128 kv][128 q] bf16 (2 chunks)
  const uint32_t ds_smem = p_smem + 2 * AT_CHUNK;         // dS^T [128 kv][128 q] bf16 (2 chunks)
  const uint32_t stat_smem = ds_smem + 2 * AT_CHUNK;      // float [2 bufs][2 (lse2, delta)][128]
  // dQ staging for the TMA reduce-add: fp32 [chunks of 32 columns][128 rows][128 B] SWIZZLE_128B.  Own region when
  // it fits (d <= 64), otherwise it aliases the P^T/dS^T tiles (dead once the tile's MMAs have retired).
  const int dq_chunks = (p.d + 31) >> 5;
  const uint32_t dq_stage = p.stage_alias ? p_smem : stat_smem + 2 * 2 * 128 * 4;
  const uint32_t bar_base = stat_smem + 2 * 2 * 128 * 4 + (p.stage_alias ? 0 : dq_chunks * AT_CHUNK);
  const uint32_t kv_full = bar_base, s_full = bar_base + 8, dp_full = bar_base + 16, pds_full = bar_base + 24,
                 dq_full = bar_base + 32;
  auto qdo_full = [&](int s) { return bar_base + 40 + 8u * s; };
  auto qdo_empty = [&](int s) { return bar_base + 56 + 8u * s; };
  const uint32_t sread = bar_base + 72;               // compute warps have read S^T_i out of TMEM
  const uint32_t tmem_slot = bar_base + 80;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));
  float* stat = reinterpret_cast<float*>(smem_raw + (stat_smem - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * AT_BKV, h = blockIdx.y, b = blockIdx.z;
  const int TQ = (p.N + AT_BQ - 1) / AT_BQ;

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) { printf("sidlsg attention bwd: dynamic smem not 1024-aligned\n"); __trap(); }
    mbar_init(kv_full, 1); mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(pds_full, 8); mbar_init(dq_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(qdo_full(s), 1); mbar_init(qdo_empty(s), 1); }
    mbar_init(sread, 8);
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t st_tmem = tmem_base, dp_tmem = tmem_base + 128, dv_tmem = tmem_base + 256,
                 dk_tmem = tmem_base + 256 + p.dpad, dq_tmem = tmem_base + 256 + 2 * p.dpad;

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (role_leader()) {
      mbar_expect_tx(kv_full, 2 * tile_bytes);
      for (int c = 0; c < p.dchunks; ++c) {
        tma_load_4d(&tmK, kv_full, k_smem + c * AT_CHUNK, c * 64, h, kv0, b);
        tma_load_4d(&tmV, kv_full, v_smem + c * AT_CHUNK, c * 64, h, kv0, b);
      }
      for (int i = 0; i < TQ; ++i) {
        const int st = i % QST;
        mbar_wait_h(p.wait_hint, qdo_empty(st), ((i / QST) & 1) ^ 1);
        mbar_expect_tx(qdo_full(st), 2 * tile_bytes);
        for (int c = 0; c < p.dchunks; ++c) {
          tma_load_4d(&tmQ, qdo_full(st), q_smem + (st * p.dchunks + c) * AT_CHUNK, c * 64, h, i * AT_BQ, b);
          tma_load_4d(&tmDO, qdo_full(st), do_smem + (st * p.dchunks + c) * AT_CHUNK, c * 64, h, i * AT_BQ, b);
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (role_leader()) {
      const int dsteps = p.dpad >> 4;
      const uint32_t idesc_s = make_idesc(128, 0, 0);
      const uint32_t idesc_kv = make_idesc(p.dpad, 0, 1);
      const uint32_t idesc_dq = make_idesc(p.dpad, 1, 1);
      mbar_wait_h(p.wait_hint, kv_full, 0);
      auto issue_s = [&](int i) {          // S^T_i = K Q_i^T
        const uint32_t qb = q_smem + (i % QST) * tile_bytes;
        for (int s = 0; s < dsteps; ++s) {
          const uint32_t off = (s >> 2) * AT_CHUNK + (s & 3) * 32;
          tc_mma_bf16(st_tmem, make_desc(k_smem + off, 1024, 0), make_desc(qb + off, 1024, 0), idesc_s, s > 0);
        }
        tc_commit(s_full);
      };
      auto issue_dp = [&](int i) {         // dP^T_i = V dO_i^T
        const uint32_t dob = do_smem + (i % QST) * tile_bytes;
        for (int s = 0; s < dsteps; ++s) {
          const uint32_t off = (s >> 2) * AT_CHUNK + (s & 3) * 32;
          tc_mma_bf16(dp_tmem, make_desc(v_smem + off, 1024, 0), make_desc(dob + off, 1024, 0), idesc_s, s > 0);
        }
        tc_commit(dp_full);
      };
      // With a two-deep Q/dO ring the S^T of tile i+1 is issued as soon as the compute warps have READ S^T_i, so
      // their exponentials for tile i+1 overlap the dV/dK/dQ MMAs of tile i.
      const bool early = QST == 2;
      mbar_wait_h(p.wait_hint, qdo_full(0), 0);
      tc_fence_after();
      issue_s(0);
      issue_dp(0);
      for (int i = 0; i < TQ; ++i) {
        const int st = i % QST;
        const uint32_t qb = q_smem + st * tile_bytes, dob = do_smem + st * tile_bytes;
        if (early && i + 1 < TQ) {
          mbar_wait_h(p.wait_hint, sread, i & 1);
          mbar_wait_h(p.wait_hint, qdo_full((i + 1) % QST), ((i + 1) / QST) & 1);
          tc_fence_after();
          issue_s(i + 1);
        }
        const bool tr = TRACE && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && i < 32;
        if (tr) p.trace[i * 16 + 8] = clock64();
        mbar_wait_h(p.wait_hint, pds_full, i & 1);
        if (tr) p.trace[i * 16 + 9] = clock64();
        tc_fence_after();
        for (int s = 0; s < 8; ++s) {      // reduction over the 128 query rows of this tile
          const uint32_t aoff = (s >> 2) * AT_CHUNK + (s & 3) * 32;
          tc_mma_bf16(dv_tmem, make_desc(p_smem + aoff, 1024, 0), make_desc(dob + s * 2048, 1024, AT_CHUNK), idesc_kv,
                      (i > 0 || s > 0) ? 1u : 0u);
        }
        for (int s = 0; s < 8; ++s) {
          const uint32_t aoff = (s >> 2) * AT_CHUNK + (s & 3) * 32;
          tc_mma_bf16(dk_tmem, make_desc(ds_smem + aoff, 1024, 0), make_desc(qb + s * 2048, 1024, AT_CHUNK), idesc_kv,
                      (i > 0 || s > 0) ? 1u : 0u);
        }
        for (int s = 0; s < 8; ++s) {      // reduction over the 128 kv rows: dQ_i = dS K
          tc_mma_bf16(dq_tmem, make_desc(ds_smem + s * 2048, 1024, AT_CHUNK), make_desc(k_smem + s * 2048, 1024, AT_CHUNK),
                      idesc_dq, s > 0);
        }
        tc_commit(qdo_empty(st));
        tc_commit(dq_full);
        if (tr) p.trace[i * 16 + 10] = clock64();
        if (i + 1 < TQ) {
          if (!early) {
            mbar_wait_h(p.wait_hint, qdo_full((i + 1) % QST), ((i + 1) / QST) & 1);
            tc_fence_after();
            issue_s(i + 1);
          }
          issue_dp(i + 1);
        }
      }
    }
  } else {
    // ===================== compute warps 0-7 =====================
    const int wg = warp >> 2;                       // which 64 query columns
    const int row = (warp & 3) * 32 + lane;         // kv row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const bool row_ok = kv0 + row < p.M;
    const float c = p.scale_log2;
    const int C = p.H * p.d;
    const int tid = threadIdx.x;                    // 0..255
    const uint64_t c2 = pack2(c, c), scale2 = pack2(p.scale, p.scale);
    const uint32_t p_row = p_smem + row * 128, ds_row = ds_smem + row * 128;
    const int rx = row & 7;
    // per-query statistics (negated, log2 domain for lse): tile i lives in stat[i & 1]; tile i+1 is fetched while
    // tile i is being processed so the global-load latency never sits on the critical path
    auto drain_dq = [&](int it) {
      // drain dQ_i (TMEM lane = query row): stage the fp32 tile in smem, then ONE thread hands it to the TMA unit
      // as a bulk reduce-add into dq_acc (the partials of the K/V tiles meet in L2, no per-thread atomics)
      mbar_wait(dq_full, it & 1);
      if (TRACE && tid == 0 && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && it + 1 < 32) p.trace[(it + 1) * 16 + 2] = clock64();
      tc_fence_after();
      if (tid == 0) tma_wait_group_read0();        // the previous reduce has finished reading the staging tile
      named_bar_sync(2, 256);
      for (int blk = wg; blk * 16 < p.dpad; blk += 2) {
        uint32_t r[16];
        tmem_ld16_nowait(dq_tmem + lane_off + blk * 16, r);
        tmem_wait_ld();
        const uint32_t base = dq_stage + (blk >> 1) * AT_CHUNK + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(base + ((((blk & 1) * 4 + j) ^ rx) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      }
      tc_fence_before();
      fence_proxy_async();
      named_bar_sync(2, 256);
      if (tid == 0) {
        for (int ch = 0; ch < dq_chunks; ++ch) tma_reduce_add_4d(&tmDQ, dq_stage + ch * AT_CHUNK, ch * 32, h, it * AT_BQ, b);
        tma_commit_group();
        if (p.stage_alias) tma_wait_group_read0();
      }
      if (p.stage_alias) named_bar_sync(2, 256);   // P^T / dS^T tiles may be rewritten only after the TMA read them
    };
    // raw global value of this thread's statistic (lse for tid < 128, delta otherwise); the negation / log2 scaling
    // is applied when the value is published to smem one iteration later, so nothing depends on the load before then
    auto load_stat = [&](int i) -> float {
      const int qi = i * AT_BQ + (tid & 127);
      if (i >= TQ || qi >= p.N) return tid < 128 ? INFINITY : 0.f;
      const long gi = ((long)b * p.H + h) * p.N + qi;
      return tid < 128 ? p.lse[gi] : p.delta[gi];
    };
    auto stat_value = [&](float raw) -> float { return tid < 128 ? raw * -1.4426950408889634f : -raw; };
    stat[tid] = stat_value(load_stat(0));
    float stat_next = load_stat(1);
    named_bar_sync(1, 256);
    for (int i = 0; i < TQ; ++i) {
      float* sb = stat + (i & 1) * 256;
      // sb holds the NEGATED statistics so they feed the packed FFMA2 / FADD2 directly
      const float* nlse2 = sb + wg * 64;
      const float* ndl = sb + 128 + wg * 64;
      uint64_t ps2[32];                      // P * scale for this thread's 64 query columns (fp32 pairs)
      uint32_t pk[32];                       // P as packed bf16 pairs, stored once the previous tile's MMAs retired
      const bool tr = TRACE && tid == 0 && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && i < 32;
      if (tr) p.trace[i * 16 + 0] = clock64();
      mbar_wait(s_full, i & 1);
      if (tr) p.trace[i * 16 + 1] = clock64();
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld32_nowait(st_tmem + lane_off + wg * 64 + half * 32, r);
        float4 nlv[8];                        // statistics fetched while the TMEM load is in flight
#pragma unroll
        for (int j = 0; j < 8; ++j) nlv[j] = *reinterpret_cast<const float4*>(nlse2 + half * 32 + 4 * j);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 nl = nlv[j >> 2];
          float t0, t1, t2, t3;
          unpack2(ffma2(pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), c2, pack2(nl.x, nl.y)), t0, t1);
          unpack2(ffma2(pack2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), c2, pack2(nl.z, nl.w)), t2, t3);
          float p0 = ex2f(t0), p1 = ex2f(t1), p2 = ex2f(t2), p3 = ex2f(t3);
          if (!row_ok) { p0 = 0.f; p1 = 0.f; p2 = 0.f; p3 = 0.f; }
          pk[half * 16 + (j >> 1)] = pack_bf16(p0, p1);
          pk[half * 16 + (j >> 1) + 1] = pack_bf16(p2, p3);
          ps2[half * 16 + (j >> 1)] = fmul2(pack2(p0, p1), scale2);
          ps2[half * 16 + (j >> 1) + 1] = fmul2(pack2(p2, p3), scale2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sread);     // S^T may be overwritten by the next tile's MMA
      if (tr) p.trace[i * 16 + 3] = clock64();   // exponentials done (slot 2 = dq_full of tile i-1 seen, inside drain_dq)
      if (i > 0) drain_dq(i - 1);            // waits for tile i-1's MMAs: P^T / dS^T tiles and dQ are ours again
      if (tr) p.trace[i * 16 + 4] = clock64();
      // publish the next tile's statistics (buffer (i+1)&1 was last read during tile i-1) and fetch tile i+2's
      stat[((i + 1) & 1) * 256 + tid] = stat_value(stat_next);
      stat_next = load_stat(i + 2);
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          st_shared_v4(p_row + wg * AT_CHUNK + (((half * 4 + qd) ^ rx) << 4), pk[half * 16 + 4 * qd],
                       pk[half * 16 + 4 * qd + 1], pk[half * 16 + 4 * qd + 2], pk[half * 16 + 4 * qd + 3]);
      if (tr) p.trace[i * 16 + 5] = clock64();
      mbar_wait(dp_full, i & 1);
      if (tr) p.trace[i * 16 + 6] = clock64();
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32], dsk[16];
        tmem_ld32_nowait(dp_tmem + lane_off + wg * 64 + half * 32, r);
        float4 ndv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ndv[j] = *reinterpret_cast<const float4*>(ndl + half * 32 + 4 * j);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 nd = ndv[j >> 2];
          float d0, d1, d2, d3;
          unpack2(fmul2(fadd2(pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), pack2(nd.x, nd.y)),
                        ps2[half * 16 + (j >> 1)]), d0, d1);
          unpack2(fmul2(fadd2(pack2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), pack2(nd.z, nd.w)),
                        ps2[half * 16 + (j >> 1) + 1]), d2, d3);
          dsk[j >> 1] = pack_bf16(d0, d1);
          dsk[(j >> 1) + 1] = pack_bf16(d2, d3);
        }
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          st_shared_v4(ds_row + wg * AT_CHUNK + (((half * 4 + qd) ^ rx) << 4), dsk[4 * qd], dsk[4 * qd + 1],
                       dsk[4 * qd + 2], dsk[4 * qd + 3]);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (tr) p.trace[i * 16 + 7] = clock64();
      named_bar_sync(1, 256);                 // statistics of tile i+1 are visible to every compute warp
    }
    drain_dq(TQ - 1);
    if (tid == 0) tma_wait_group_read0();
    // the last dq_full commit covered every MMA: dV / dK accumulators are final. wg0 stores dV, wg1 stores dK.
    {
      const uint32_t acc = wg == 0 ? dv_tmem : dk_tmem;
      bf16* out = (wg == 0 ? p.dv : p.dk) + ((long)b * p.M + kv0 + row) * (wg == 0 ? p.lddv : p.lddk) + h * p.d;
      for (int cc = 0; cc < p.dpad; cc += 16) {
        uint32_t r[16];
        tmem_ld16_nowait(acc + lane_off + cc, r);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (cc + 8 * hh < p.d) {
              uint4 o;
              o.x = pack_bf16(__uint_as_float(r[8 * hh + 0]), __uint_as_float(r[8 * hh + 1]));
              o.y = pack_bf16(__uint_as_float(r[8 * hh + 2]), __uint_as_float(r[8 * hh + 3]));
              o.z = pack_bf16(__uint_as_float(r[8 * hh + 4]), __uint_as_float(r[8 * hh + 5]));
              o.w = pack_bf16(__uint_as_float(r[8 * hh + 6]), __uint_as_float(r[8 * hh + 7]));
              *reinterpret_cast<uint4*>(out + cc + 8 * hh) = o;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward for d <= 64, pipelined (the phase trace of attn_bwd_kernel - scripts/trace_attn_bwd.py, profiles/ - showed its
// 8 compute warps busy or waiting on each other for 4000 of the 4600 clk of a tile: exponentials 1530, dQ drain 1150
// (two 256-thread barriers of skew), P^T store 480, dS 850, strictly in series with the 24 MMAs that follow).  Same
// arithmetic and orientation, different choreography:
//   * P^T is written IN PLACE over S^T in TMEM (bf16 pairs, tcgen05.st) and is the TMEM A operand of dV - no shared
//     memory store, no wait for a free P tile;  dV(i) is issued as soon as P^T(i) exists, S^T(i+1) right behind it, so
//     the next tile's scores are ready while the compute warps are still on dS(i);
//   * dS^T is double buffered in shared memory, dQ double buffered in TMEM;
//   * four dedicated drain warps move dQ(i) TMEM -> shared -> TMA reduce-add; the compute warps never touch it;
//   * per-query statistics (-lse log2 e, -delta) arrive with the Q / dO tiles as TMA bulk copies: no per-tile 256-thread
//     barrier is left in the compute loop.
// 448 threads: warps 0-7 compute (two warpgroups split the 128 query columns), 8-11 dQ drain, 12 TMA, 13 MMA + TMEM.
// TMEM: S^T/P^T 128 | dP^T 128 | dV | dK | dQ0 | dQ1 (dpad each) <= 512 columns.  Needs N % 128 == 0.
constexpr int AB2_THREADS = 448;
constexpr int AB2_QST = 3;

template <bool TRACE>
__global__ void __launch_bounds__(AB2_THREADS, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const AttnBwdParams p, const float* __restrict__ nlse2) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t k_smem = smem_base;                       // 16 KB
  const uint32_t v_smem = k_smem + AT_CHUNK;               // 16 KB
  const uint32_t q_smem = v_smem + AT_CHUNK;               // AB2_QST x 16 KB (Q is held from S^T_i to dK_i: 3 stages)
  const uint32_t do_smem = q_smem + AB2_QST * AT_CHUNK;    // 2 x 16 KB (dO is released after dV_i, half a tile earlier)
  const uint32_t ds_smem = do_smem + 2 * AT_CHUNK;         // dS^T: 2 buffers x [128 kv][128 q] bf16 (2 chunks each)
  const uint32_t dq_stage = ds_smem + 4 * AT_CHUNK;        // fp32 [chunks of 32 columns][128 rows][128 B], 2 chunks
  const uint32_t stat_smem = dq_stage + 2 * AT_CHUNK;      // float [2 stages][nlse2 128 | ndelta 128], own ring (released by
  const uint32_t bar_base = stat_smem + 2 * 1024;          // the compute warps after dS_i; 227 KB leaves no room for a third)
  const uint32_t kv_full = bar_base, s_full = bar_base + 8, p_full = bar_base + 16, dp_full = bar_base + 24,
                 ds_full = bar_base + 32, done_bar = bar_base + 40;
  auto q_full = [&](int s) { return bar_base + 48 + 8u * s; };       // [AB2_QST]
  auto q_empty = [&](int s) { return bar_base + 80 + 8u * s; };
  auto do_full = [&](int s) { return bar_base + 112 + 8u * s; };     // [2]
  auto do_empty = [&](int s) { return bar_base + 128 + 8u * s; };
  auto dq_full = [&](int s) { return bar_base + 144 + 8u * s; };
  auto dq_empty = [&](int s) { return bar_base + 160 + 8u * s; };
  auto stat_full = [&](int s) { return bar_base + 176 + 8u * s; };
  auto stat_empty = [&](int s) { return bar_base + 192 + 8u * s; };
  const uint32_t tmem_slot = bar_base + 208;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));
  const float* stat = reinterpret_cast<const float*>(smem_raw + (stat_smem - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * AT_BKV, h = blockIdx.y, b = blockIdx.z;
  const int TQ = p.N / AT_BQ;

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) { printf("sidlsg attention bwd: dynamic smem not 1024-aligned\n"); __trap(); }
    mbar_init(kv_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 8); mbar_init(dp_full, 1); mbar_init(ds_full, 8);
    mbar_init(done_bar, 1);
    for (int s = 0; s < AB2_QST; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(do_full(s), 1); mbar_init(do_empty(s), 1); mbar_init(dq_full(s), 1); mbar_init(dq_empty(s), 4);
      mbar_init(stat_full(s), 1); mbar_init(stat_empty(s), 8);
    }
    fence_barrier_init();
  }
  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
  }
  if (warp == 13) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t st_tmem = tmem_base, dp_tmem = tmem_base + 128, dv_tmem = tmem_base + 256,
                 dk_tmem = tmem_base + 256 + p.dpad;
  auto dq_tmem = [&](int s) { return tmem_base + 256 + (2 + s) * p.dpad; };
  // (dS^T as a second TMEM A operand, for dK, was measured SLOWER - 0.867 -> 1.02 ms at B8 N4096 d40: the extra TMEM
  // reads of the MMA collide with the compute warps' tcgen05.ld / st traffic; dK and dQ read dS^T from shared memory)

  if (warp == 12) {
    // ===================== TMA producer =====================
    if (role_leader()) {
      mbar_expect_tx(kv_full, 2 * AT_CHUNK);
      tma_load_4d(&tmK, kv_full, k_smem, 0, h, kv0, b);
      tma_load_4d(&tmV, kv_full, v_smem, 0, h, kv0, b);
      const long srow = ((long)b * p.H + h) * p.N;
      for (int i = 0; i < TQ; ++i) {
        const int qs = i % AB2_QST, ds = i & 1;
        mbar_wait_h(p.wait_hint, q_empty(qs), ((i / AB2_QST) & 1) ^ 1);
        mbar_expect_tx(q_full(qs), AT_CHUNK);
        tma_load_4d(&tmQ, q_full(qs), q_smem + qs * AT_CHUNK, 0, h, i * AT_BQ, b);
        mbar_wait_h(p.wait_hint, stat_empty(ds), ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(stat_full(ds), 1024);
        bulk_load_1d(stat_smem + ds * 1024, nlse2 + srow + (long)i * AT_BQ, 512, stat_full(ds));
        bulk_load_1d(stat_smem + ds * 1024 + 512, p.delta + srow + (long)i * AT_BQ, 512, stat_full(ds));
        mbar_wait_h(p.wait_hint, do_empty(ds), ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(do_full(ds), AT_CHUNK);
        tma_load_4d(&tmDO, do_full(ds), do_smem + ds * AT_CHUNK, 0, h, i * AT_BQ, b);
      }
    }
  } else if (warp == 13) {
    // ===================== MMA issuer =====================
    // Low-latency form (see attn_fwd3_kernel): this single thread's dependent instruction stream was the period of the
    // kernel (trace r02: ~3,300 clk per query tile against ~2,500 clk of tensor work; ~250 clk per wait on a barrier that
    // had completed long before, ~80-100 clk per MMA).  Suspend hint in a register, descriptors = base + immediate,
    // loops unrolled, stage indices and phases carried incrementally (no division).
    if (role_leader()) {
      uint32_t hint;
      asm volatile("mov.u32 %0, %1;" : "=r"(hint) : "r"(p.wait_hint));
      auto wait = [&](uint32_t bar, uint32_t parity) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok) mbar_wait_h(hint, bar, parity);
      };
      const int dsteps = p.dpad >> 4;
      const uint32_t idesc_s = make_idesc(128, 0, 0);
      const uint32_t idesc_kv = make_idesc(p.dpad, 0, 1);
      const uint32_t idesc_dq = make_idesc(p.dpad, 1, 1);
      constexpr uint64_t MN = (uint64_t)(AT_CHUNK >> 4) << 16;        // leading-dimension byte offset field: MN-major operand
      constexpr uint64_t STG = AT_CHUNK >> 4;                         // one 16 KB stage in descriptor address units
      const uint64_t kd = make_desc(k_smem, 1024, 0), vd = make_desc(v_smem, 1024, 0);
      const uint64_t qd0 = make_desc(q_smem, 1024, 0), dod0 = make_desc(do_smem, 1024, 0), dsd0 = make_desc(ds_smem, 1024, 0);
      const uint32_t dq0 = dq_tmem(0), dq1 = dq_tmem(1);
      auto issue_s = [&](uint64_t qd) {    // S^T_i = K Q_i^T
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (s < dsteps) tc_mma_bf16(st_tmem, kd + 2 * s, qd + 2 * s, idesc_s, s > 0);
        tc_commit(s_full);
      };
      auto issue_dp = [&](uint64_t dod) {  // dP^T_i = V dO_i^T
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (s < dsteps) tc_mma_bf16(dp_tmem, vd + 2 * s, dod + 2 * s, idesc_s, s > 0);
        tc_commit(dp_full);
      };
      wait(kv_full, 0);
      wait(q_full(0), 0);
      tc_fence_after();
      issue_s(qd0);
      wait(do_full(0), 0);
      tc_fence_after();
      issue_dp(dod0);
      int qs = 0, qph = 0;                 // Q ring stage of tile i and its phase
#pragma unroll 1
      for (int i = 0; i < TQ; ++i) {
        const int st = i & 1;
        const int qs1 = qs + 1 == AB2_QST ? 0 : qs + 1, qph1 = qs + 1 == AB2_QST ? qph ^ 1 : qph;
        const uint64_t qd = qd0 + (uint64_t)qs * STG, qd1 = qd0 + (uint64_t)qs1 * STG;
        const uint64_t dod = dod0 + (uint64_t)st * STG, dod1 = dod0 + (uint64_t)(st ^ 1) * STG;
        const uint64_t dsd = dsd0 + (uint64_t)st * (2 * STG);
        const uint32_t acc0 = i > 0 ? 1u : 0u, par = i & 1;
        const bool more = i + 1 < TQ;
        const bool tr = TRACE && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && i < 32;
        if (tr) p.trace[i * 16 + 8] = clock64();
        wait(p_full, par);
        if (tr) p.trace[i * 16 + 9] = clock64();
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < 8; ++s)        // dV += P^T_i dO_i: A = P^T from TMEM (queries 0-63 at columns 0.., 64-127 at 64..)
          tc_mma_bf16_ta(dv_tmem, st_tmem + (s >> 2) * 64 + (s & 3) * 8, (dod | MN) + 128 * s, idesc_kv, s > 0 ? 1u : acc0);
        tc_commit(do_empty(st));           // dO_i has been read for the last time once dV_i retires
        if (more) {                        // the next tile's scores overwrite P^T_i right behind the product that read it
          wait(q_full(qs1), qph1);
          tc_fence_after();
          issue_s(qd1);
        }
        if (tr) p.trace[i * 16 + 10] = clock64();
        wait(ds_full, par);
        if (tr) p.trace[i * 16 + 11] = clock64();
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < 8; ++s)        // dK += dS^T_i Q_i
          tc_mma_bf16(dk_tmem, dsd + (s >> 2) * STG + (s & 3) * 2, (qd | MN) + 128 * s, idesc_kv, s > 0 ? 1u : acc0);
        if (tr) p.trace[i * 16 + 12] = clock64();
        wait(dq_empty(st), ((i >> 1) & 1) ^ 1);   // the drain warps have read dQ_{i-2}
        if (tr) p.trace[i * 16 + 13] = clock64();
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < 8; ++s)        // dQ_i = dS_i K (reduction over the 128 kv rows)
          tc_mma_bf16(st ? dq1 : dq0, (dsd | MN) + 128 * s, (kd | MN) + 128 * s, idesc_dq, s > 0);
        tc_commit(q_empty(qs));
        tc_commit(dq_full(st));
        if (more) {
          wait(do_full(st ^ 1), ((i + 1) >> 1) & 1);
          tc_fence_after();
          issue_dp(dod1);
        }
        if (tr) p.trace[i * 16 + 14] = clock64();
        qs = qs1; qph = qph1;
      }
      tc_commit(done_bar);
    }
  } else if (warp >= 8) {
    // ===================== dQ drain (warps 8-11: TMEM lane quarter = warp & 3 = 32 query rows) =====================
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const int rx = row & 7;
    const int dq_chunks = (p.d + 31) >> 5;
    const bool leader = threadIdx.x == 256;
    for (int i = 0; i < TQ; ++i) {
      const int st = i & 1;
      mbar_wait(dq_full(st), (i >> 1) & 1);
      if (TRACE && leader && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && i < 32) p.trace[i * 16 + 6] = clock64();
      tc_fence_after();
      if (leader) tma_wait_group_read0();          // the previous reduce has finished reading the staging tile
      named_bar_sync(3, 128);
      for (int blk = 0; blk * 16 < p.dpad; ++blk) {
        uint32_t r[16];
        tmem_ld16_nowait(dq_tmem(st) + lane_off + blk * 16, r);
        tmem_wait_ld();
        const uint32_t base = dq_stage + (blk >> 1) * AT_CHUNK + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_shared_v4(base + ((((blk & 1) * 4 + j) ^ rx) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty(st));    // this dQ accumulator may be overwritten (tile i + 2)
      named_bar_sync(3, 128);
      if (leader) {
        for (int ch = 0; ch < dq_chunks; ++ch) tma_reduce_add_4d(&tmDQ, dq_stage + ch * AT_CHUNK, ch * 32, h, i * AT_BQ, b);
        tma_commit_group();
        if (TRACE && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && i < 32) p.trace[i * 16 + 7] = clock64();
      }
    }
    if (leader) tma_wait_group0();
  } else {
    // ===================== compute warps 0-7 =====================
    const int wg = warp >> 2;                       // which 64 query columns
    const int row = (warp & 3) * 32 + lane;         // kv row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const bool row_ok = kv0 + row < p.M;
    const uint64_t c2 = pack2(p.scale_log2, p.scale_log2), scale2 = pack2(p.scale, p.scale);
    const float rowbias = row_ok ? 0.f : -INFINITY;
    const uint64_t rowbias2 = pack2(rowbias, rowbias);
    const int rx = row & 7;
    for (int i = 0; i < TQ; ++i) {
      const int st = i & 1;
      const float* nlse = stat + st * 256 + wg * 64;
      const float* ndl = stat + st * 256 + 128 + wg * 64;
      const uint32_t ds_row = ds_smem + st * 2 * AT_CHUNK + wg * AT_CHUNK + row * 128;
      uint64_t ps2[32];                      // P * scale for this thread's 64 query columns (fp32 pairs)
      const bool tr = TRACE && threadIdx.x == 0 && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && i < 32;
      if (tr) p.trace[i * 16 + 0] = clock64();
      mbar_wait(stat_full(st), (i >> 1) & 1);    // the statistics of this query tile have landed
      mbar_wait(s_full, i & 1);
      if (tr) p.trace[i * 16 + 1] = clock64();
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32], pk[16];
        tmem_ld32_nowait(st_tmem + lane_off + wg * 64 + half * 32, r);
        float4 nlv[8];                        // statistics fetched while the TMEM load is in flight
#pragma unroll
        for (int j = 0; j < 8; ++j) nlv[j] = *reinterpret_cast<const float4*>(nlse + half * 32 + 4 * j);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 nl = nlv[j >> 2];
          // rows past M (last key tile only) get -inf added: P = 0 there
          const uint64_t ta = fadd2(ffma2(pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), c2, pack2(nl.x, nl.y)), rowbias2);
          const uint64_t tb = fadd2(ffma2(pack2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), c2, pack2(nl.z, nl.w)), rowbias2);
          // (the FMA-pipe polynomial that helps the forward was measured slower here - 0.867 -> 0.925 ms: these warps
          // already fill the FMA pipe with the dS arithmetic - so every exponential goes to the MUFU pipe)
          float t0, t1, t2, t3;
          unpack2(ta, t0, t1);
          unpack2(tb, t2, t3);
          const float p0 = ex2f(t0), p1 = ex2f(t1), p2 = ex2f(t2), p3 = ex2f(t3);
          pk[j >> 1] = pack_bf16(p0, p1);
          pk[(j >> 1) + 1] = pack_bf16(p2, p3);
          ps2[half * 16 + (j >> 1)] = fmul2(pack2(p0, p1), scale2);
          ps2[half * 16 + (j >> 1) + 1] = fmul2(pack2(p2, p3), scale2);
        }
        // P^T in place (bf16 pairs): half h lands on columns [wg*64 + 16h, +16) - all inside [wg*64, wg*64 + 32), which
        // this thread finished reading with its first load, and nobody else touches this lane's columns of this warpgroup
        tmem_st16(st_tmem + lane_off + wg * 64 + half * 16, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (tr) p.trace[i * 16 + 2] = clock64();
      mbar_wait(dp_full, i & 1);
      if (tr) p.trace[i * 16 + 3] = clock64();
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32], dsk[16];
        tmem_ld32_nowait(dp_tmem + lane_off + wg * 64 + half * 32, r);
        float4 ndv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ndv[j] = *reinterpret_cast<const float4*>(ndl + half * 32 + 4 * j);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 nd = ndv[j >> 2];
          float d0, d1, d2, d3;
          unpack2(fmul2(fadd2(pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), pack2(nd.x, nd.y)),
                        ps2[half * 16 + (j >> 1)]), d0, d1);
          unpack2(fmul2(fadd2(pack2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), pack2(nd.z, nd.w)),
                        ps2[half * 16 + (j >> 1) + 1]), d2, d3);
          dsk[j >> 1] = pack_bf16(d0, d1);
          dsk[(j >> 1) + 1] = pack_bf16(d2, d3);
        }
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          st_shared_v4(ds_row + (((half * 4 + qd) ^ rx) << 4), dsk[4 * qd], dsk[4 * qd + 1], dsk[4 * qd + 2],
                       dsk[4 * qd + 3]);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(ds_full); mbar_arrive(stat_empty(st)); }
      if (tr) p.trace[i * 16 + 4] = clock64();
    }
    // every MMA has retired: dV / dK accumulators are final.  wg0 stores dV, wg1 stores dK.
    mbar_wait(done_bar, 0);
    tc_fence_after();
    {
      const uint32_t acc = wg == 0 ? dv_tmem : dk_tmem;
      bf16* out = (wg == 0 ? p.dv : p.dk) + ((long)b * p.M + kv0 + row) * (wg == 0 ? p.lddv : p.lddk) + h * p.d;
      for (int cc = 0; cc < p.dpad; cc += 16) {
        uint32_t r[16];
        tmem_ld16_nowait(acc + lane_off + cc, r);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (cc + 8 * hh < p.d) {
              uint4 o;
              o.x = pack_bf16(__uint_as_float(r[8 * hh + 0]), __uint_as_float(r[8 * hh + 1]));
              o.y = pack_bf16(__uint_as_float(r[8 * hh + 2]), __uint_as_float(r[8 * hh + 3]));
              o.z = pack_bf16(__uint_as_float(r[8 * hh + 4]), __uint_as_float(r[8 * hh + 5]));
              o.w = pack_bf16(__uint_as_float(r[8 * hh + 6]), __uint_as_float(r[8 * hh + 7]));
              *reinterpret_cast<uint4*>(out + cc + 8 * hh) = o;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// delta[b,h,q] = sum_c O[b,q,h*d+c] * dO[b,q,h*d+c].  A block takes `rows` consecutive (b,q) rows; thread = one
// 16-byte piece (8 channels, never straddling a head since d % 8 == 0) of one row, so the block reads contiguous
// memory; the d/8 piece sums of a head meet in shared memory.  (The one-warp-per-(b,q,h) version moved 80-byte
// segments per warp and ran at 0.9 TB/s.)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta, long BN, int N,
                  int H, int d, int ppr, int rows, const float* __restrict__ lse, float* __restrict__ nlse2) {
  __shared__ float part[256];
  const int t = threadIdx.x;
  const int rl = t / ppr, piece = t - rl * ppr;
  const long row = (long)blockIdx.x * rows + rl;
  float s = 0.f;
  if (rl < rows && row < BN) {
    const long off = row * (long)ppr * 8 + piece * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(o + off);
    const uint4 g = *reinterpret_cast<const uint4*>(dout + off);
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      s = fmaf(__low2float(a2[i]), __low2float(g2[i]), fmaf(__high2float(a2[i]), __high2float(g2[i]), s));
  }
  part[t] = s;
  __syncthreads();
  if (t < rows * H) {
    const int r2 = t / H, hh = t - r2 * H;
    const long row2 = (long)blockIdx.x * rows + r2;
    if (row2 < BN) {
      const int pph = d >> 3;
      const float* pp = part + r2 * ppr + hh * pph;
      float acc = 0.f;
      for (int i = 0; i < pph; ++i) acc += pp[i];
      const long bb = row2 / N, q = row2 - bb * N;
      const long gi = (bb * H + hh) * N + q;
      // nlse2 given (two-stage-pipelined backward): statistics are stored NEGATED, lse in the log2 domain, so they feed
      // the packed FFMA2 / FADD2 of the consumer directly and travel as plain TMA bulk copies
      delta[gi] = nlse2 ? -acc : acc;
      if (nlse2) nlse2[gi] = lse[gi] * -1.4426950408889634f;
    }
  }
}

// dq (bf16, row stride ld) = dq_acc (fp32 dense [rows, C]); 8 elements per thread
__global__ void attn_dq_cast_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long rows, int C, long ld) {
  const int vpr = C >> 3;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * vpr) return;
  const long r = i / vpr;
  const int c = (int)(i - r * vpr) << 3;
  const float4 a = *reinterpret_cast<const float4*>(acc + r * C + c);
  const float4 b2 = *reinterpret_cast<const float4*>(acc + r * C + c + 4);
  uint4 o;
  o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w); o.z = pack_bf16(b2.x, b2.y); o.w = pack_bf16(b2.z, b2.w);
  *reinterpret_cast<uint4*>(dq + r * ld + c) = o;
}

static bool g_attn_bwd_attr_done = false;

}  // namespace sidlsg

// dq_acc: fp32 [B,N,H*d] scratch (zeroed here), delta: fp32 [2,B,H,N] scratch (delta and -lse log2 e).  dq/dk/dv bf16.
// q/k/v and dq/dk/dv may be column slices of packed tensors: ld* = their row strides in elements (o, dout dense).
// Supports d % 8 == 0, 16 <= d <= 80 (TMEM: 256 + 3*dpad <= 512 columns); returns SIDLSG_ERR_UNSUPPORTED otherwise.
static int attention_bwd_impl(const void* q, const void* k, const void* v, const void* o, const void* dout,
                              const float* lse, float* delta, float* dq_acc, void* dq, void* dk, void* dv,
                              int B, int N, int M, int H, int d, long ldq, long ldk, long ldv, long lddq,
                              long lddk, long lddv, long long* trace, void* stream) {
  if (!tc_enabled()) { set_error("attention_bwd: tcgen05 path unavailable on this device"); return SIDLSG_ERR_UNSUPPORTED; }
  const int dpad = (d + 15) & ~15;
  if (d % 8 || d < 16 || 256 + 3 * dpad > 512 || B <= 0 || N <= 0 || M <= 0 || H <= 0 || H > 65535 || B > 65535) {
    set_error("attention_bwd: unsupported shape B=%d N=%d M=%d H=%d d=%d", B, N, M, H, d);
    return SIDLSG_ERR_UNSUPPORTED;
  }
  {
    const long C0 = (long)H * d;
    if (ldq < C0 || ldk < C0 || ldv < C0 || lddq < C0 || lddk < C0 || lddv < C0 ||
        ((ldq | ldk | ldv | lddq | lddk | lddv) & 7)) {
      set_error("attention_bwd: row strides must be >= H*d and multiples of 8 elements");
      return SIDLSG_ERR_ARG;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(dq_acc, 0, sizeof(float) * (size_t)B * N * H * d, st);
  // d <= 64 and whole query tiles: the pipelined kernel (attn_bwd2_kernel); SIDLSG_ATTN_BWD2=0 keeps the first one (A/B)
  static int bwd2 = -1;
  if (bwd2 < 0) { const char* e = getenv("SIDLSG_ATTN_BWD2"); bwd2 = (e && e[0] == '0') ? 0 : 1; }
  const bool use_bwd2 = bwd2 && dpad <= 64 && (N % AT_BQ) == 0;
  float* nlse2 = delta + (size_t)B * H * N;          // second half of the statistics scratch
  {
    const int ppr = H * d / 8;                       // 16-byte pieces per (b,q) row; H*d <= 2048 on this path
    if (ppr > 256) { set_error("attention_bwd: H*d = %d too wide for the delta kernel", H * d); return SIDLSG_ERR_UNSUPPORTED; }
    const int rows = 256 / ppr;
    const long BN = (long)B * N;
    attn_delta_kernel<<<(unsigned)((BN + rows - 1) / rows), 256, 0, st>>>((const bf16*)o, (const bf16*)dout, delta, BN, N,
                                                                             H, d, ppr, rows, lse, use_bwd2 ? nlse2 : nullptr);
  }
  AttnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.N = N; p.M = M; p.d = d;
  p.dchunks = (d + 63) / 64;
  p.dpad = dpad;
  p.q_stages = p.dchunks == 1 ? 2 : 1;
  p.scale = 1.f / sqrtf((float)d);
  p.scale_log2 = p.scale * 1.4426950408889634f;
  p.lse = lse; p.delta = delta; p.dq_acc = dq_acc; p.dk = (bf16*)dk; p.dv = (bf16*)dv;
  p.lddk = lddk; p.lddv = lddv;
  p.stage_alias = d > 64;
  p.wait_hint = tc_wait_hint_ns();
  p.trace = trace;
  CUtensorMap tq, tk, tv, tdo, tdq;
  if (!make_head_map(&tq, q, d, H, N, B, ldq) || !make_head_map(&tk, k, d, H, M, B, ldk) ||
      !make_head_map(&tv, v, d, H, M, B, ldv) || !make_head_map(&tdo, dout, d, H, N, B, (long)H * d))
    return SIDLSG_ERR_CUDA;
  {
    const long C = (long)H * d;
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)H, (uint64_t)N, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)d * 4, (uint64_t)C * 4, (uint64_t)N * C * 4};
    uint32_t box[4] = {32, 1, 128, 1};
    if (!make_map(&tdq, dq_acc, 4, dims, strides, box, nullptr, 1)) return SIDLSG_ERR_CUDA;
  }
  const int smem = (2 + 2 * p.q_stages) * p.dchunks * AT_CHUNK + 4 * AT_CHUNK + 2048 + 256 +
                   (p.stage_alias ? 0 : ((d + 31) / 32) * AT_CHUNK);
  if (!g_attn_bwd_attr_done) {
    cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    g_attn_bwd_attr_done = true;
  }
  dim3 grid((M + AT_BKV - 1) / AT_BKV, H, B);
  if (use_bwd2) {
    static bool attr2 = false;
    if (!attr2) {
      cudaFuncSetAttribute(attn_bwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(attn_bwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      attr2 = true;
    }
    const int smem2 = (11 + AB2_QST) * AT_CHUNK + 2 * 1024 + 256;
    if (trace) attn_bwd2_kernel<true><<<grid, AB2_THREADS, smem2, st>>>(tq, tk, tv, tdo, tdq, p, nlse2);
    else attn_bwd2_kernel<false><<<grid, AB2_THREADS, smem2, st>>>(tq, tk, tv, tdo, tdq, p, nlse2);
  } else if (trace) attn_bwd_kernel<true><<<grid, AB_THREADS, smem, st>>>(tq, tk, tv, tdo, tdq, p);
  else attn_bwd_kernel<false><<<grid, AB_THREADS, smem, st>>>(tq, tk, tv, tdo, tdq, p);
  int r = check_launch("attention_bwd");
  if (r != SIDLSG_OK) return r;
  // dq (bf16) = dq_acc (fp32)
  {
    const long rows = (long)B * N;
    const int C = H * d;
    attn_dq_cast_kernel<<<(unsigned)((rows * (C >> 3) + 255) / 256), 256, 0, st>>>(dq_acc, (bf16*)dq, rows, C, lddq);
  }
  return check_launch("attention_bwd dq cast");
}

extern "C" int sidlsg_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout,
                                    const float* lse, float* delta, float* dq_acc, void* dq, void* dk, void* dv,
                                    int B, int N, int M, int H, int d, long ldq, long ldk, long ldv, long lddq,
                                    long lddk, long lddv, void* stream) {
  return attention_bwd_impl(q, k, v, o, dout, lse, delta, dq_acc, dq, dk, dv, B, N, M, H, d, ldq, ldk, ldv, lddq, lddk,
                            lddv, nullptr, stream);
}

// Development aid: the same backward with in-kernel clock64 stamps of CTA (1,0,0) written to `trace` (device memory,
// 32 tiles x 16 slots of long long; see attn_bwd_kernel<true>).  Not used by the product path.
extern "C" int sidlsg_debug_attention_bwd_trace(const void* q, const void* k, const void* v, const void* o,
                                                const void* dout, const float* lse, float* delta, float* dq_acc, void* dq,
                                                void* dk, void* dv, int B, int N, int M, int H, int d, long ldq, long ldk,
                                                long ldv, long lddq, long lddk, long lddv, void* trace, void* stream) {
  return attention_bwd_impl(q, k, v, o, dout, lse, delta, dq_acc, dq, dk, dv, B, N, M, H, d, ldq, ldk, ldv, lddq, lddk,
                            lddv, (long long*)trace, stream);
}
