// DRAFT - NOT BUILT, NOT VALIDATED ON A GPU.  Two-query-tile attention forward (FlashAttention-4 schedule) for d <= 64:
// the plan of experiments/NEXT_ROUND.md section 1.  Paste inside `namespace sidlsg` of csrc/attention_tc.cu after
// attn_fwd_kernel (it uses that file's helpers: ex2f, pack2, ffma2, fadd2, pack_bf16, st_shared_v4, AttnParams) and add a
// dispatch `if (p.dchunks == 1 && env) attn_fwd_2q_kernel<<<dim3((N + 255) / 256, H, B), 320, smem, st>>>(tq, tk, tv, p)`
// with smem = (2 + 2 * ST + 4) * AT_CHUNK + 256 and ST = 3 (p.kv_stages).  Compile check: experiments/README.md.
//
// One CTA per SM, 320 threads: warps 0-3 = softmax of query tile A (rows q0..q0+127), warps 4-7 = tile B (q0+128..),
// warp 8 = TMA, warp 9 = MMA + TMEM allocator.  TMEM (512 columns): S_A 0..127 | S_B 128..255 | O_A 256.. | O_B 384..
// The MMA warp alternates between the tiles (PV_A(j), QK_A(j+1), PV_B(j), QK_B(j+1)), so while warpgroup A waits for
// S_A(j+1) warpgroup B runs its exponentials: the MUFU pipe and the tensor pipe overlap inside ONE CTA instead of
// relying on two co-resident CTAs whose MMA phases collide (phase trace: period 4600 clk per tile per CTA today).
__global__ void __launch_bounds__(320, 1)
attn_fwd_2q_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const int ST = p.kv_stages;                       // >= 2
  const uint32_t q_smem0 = smem_base;               // Q_A, Q_B: 16 KB each (d <= 64: one 64-column chunk)
  const uint32_t k_smem = q_smem0 + 2 * AT_CHUNK;
  const uint32_t v_smem = k_smem + ST * AT_CHUNK;
  const uint32_t p_smem0 = v_smem + ST * AT_CHUNK;  // P_A, P_B: 2 chunks (32 KB) each
  const uint32_t bar_base = p_smem0 + 4 * AT_CHUNK;
  const uint32_t q_full = bar_base;
  auto s_full = [&](int x) { return bar_base + 8 + 8u * x; };
  auto p_full = [&](int x) { return bar_base + 24 + 8u * x; };
  auto o_full = [&](int x) { return bar_base + 40 + 8u * x; };
  auto k_full = [&](int s) { return bar_base + 56 + 8u * s; };
  auto v_full = [&](int s) { return bar_base + 88 + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 120 + 8u * s; };
  const uint32_t tmem_slot = bar_base + 152;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * AT_BQ, h = blockIdx.y, b = blockIdx.z;
  const int T = (p.M + AT_BKV - 1) / AT_BKV;

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) { printf("sidlsg attention: dynamic smem not 1024-aligned\n"); __trap(); }
    mbar_init(q_full, 1);
    for (int x = 0; x < 2; ++x) { mbar_init(s_full(x), 1); mbar_init(p_full(x), 4); mbar_init(o_full(x), 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(k_full(s), 1); mbar_init(v_full(s), 1); mbar_init(kv_empty(s), 1); }
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * AT_CHUNK);
      tma_load_4d(&tmQ, q_full, q_smem0, 0, h, q0, b);
      tma_load_4d(&tmQ, q_full, q_smem0 + AT_CHUNK, 0, h, q0 + AT_BQ, b);   // rows >= N: zero fill
      for (int j = 0; j < T; ++j) {
        const int st = j % ST;
        mbar_wait_h(p.wait_hint, kv_empty(st), ((j / ST) & 1) ^ 1);
        mbar_expect_tx(k_full(st), AT_CHUNK);
        tma_load_4d(&tmK, k_full(st), k_smem + st * AT_CHUNK, 0, h, j * AT_BKV, b);
        mbar_expect_tx(v_full(st), AT_CHUNK);
        tma_load_4d(&tmV, v_full(st), v_smem + st * AT_CHUNK, 0, h, j * AT_BKV, b);
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const int dsteps = p.dpad >> 4;
      const uint32_t idesc_pv = make_idesc(p.dpad, 0, 1);
      auto issue_qk = [&](int x, int j) {             // S_x = Q_x K_j^T, then s_full[x]
        const int n_valid = min(AT_BKV, p.M - j * AT_BKV);
        const uint32_t idesc = make_idesc((n_valid + 15) & ~15, 0, 0);
        const uint32_t kb = k_smem + (j % ST) * AT_CHUNK, qb = q_smem0 + x * AT_CHUNK;
        for (int s = 0; s < dsteps; ++s)
          tc_mma_bf16(tmem_base + x * 128, make_desc(qb + s * 32, 1024, 0), make_desc(kb + s * 32, 1024, 0), idesc, s > 0);
        tc_commit(s_full(x));
      };
      auto issue_pv = [&](int x, int j) {             // O_x += P_x V_j
        const int n_valid = min(AT_BKV, p.M - j * AT_BKV);
        const uint32_t vb = v_smem + (j % ST) * AT_CHUNK, pb = p_smem0 + x * 2 * AT_CHUNK;
        const int ksteps = (n_valid + 15) >> 4;
        for (int s = 0; s < ksteps; ++s)
          tc_mma_bf16(tmem_base + 256 + x * 128, make_desc(pb + (s >> 2) * AT_CHUNK + (s & 3) * 32, 1024, 0),
                      make_desc(vb + s * 2048, 1024, AT_CHUNK), idesc_pv, (j > 0 || s > 0) ? 1u : 0u);
      };
      mbar_wait_h(p.wait_hint, q_full, 0);
      mbar_wait_h(p.wait_hint, k_full(0), 0);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      for (int j = 0; j < T; ++j) {
        const int st = j % ST;
        const bool more = j + 1 < T;
        mbar_wait_h(p.wait_hint, p_full(0), j & 1);
        mbar_wait_h(p.wait_hint, v_full(st), (j / ST) & 1);
        tc_fence_after();
        issue_pv(0, j);
        tc_commit(o_full(0));
        if (more) {
          mbar_wait_h(p.wait_hint, k_full((j + 1) % ST), ((j + 1) / ST) & 1);
          tc_fence_after();
          issue_qk(0, j + 1);                         // S_A is free: P_A(j) has been published
        }
        mbar_wait_h(p.wait_hint, p_full(1), j & 1);
        tc_fence_after();
        issue_pv(1, j);
        tc_commit(kv_empty(st));                      // last reader of K_j / V_j
        tc_commit(o_full(1));
        if (more) issue_qk(1, j + 1);
      }
    }
  } else {
    // ===================== softmax / output: warpgroup x = warp / 4 owns query tile x =====================
    const int x = warp >> 2;
    const int wq = warp & 3;                          // TMEM lane quarter of this warp
    const int row = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t s_tmem = tmem_base + x * 128, o_tmem = tmem_base + 256 + x * 128;
    const uint32_t p_smem = p_smem0 + x * 2 * AT_CHUNK;
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
      mbar_wait(s_full(x), j & 1);
      tc_fence_after();
      if (!spec) {
        const float m_new = fmaxf(m_used, tile_max(n_valid));
        const bool need = (m_new - m_used) * c > 8.f;
        if (j > 0) {
          mbar_wait(o_full(x), (j - 1) & 1);          // P_x V of the previous tile retired: O_x and P_x are ours again
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
            const float f = need ? ex2f((m_used - m_new) * c) : 1.f;
            rescale_o(f);
            l_run *= f;
          }
        }
        if (need) m_used = m_new;
      } else {
        mbar_wait(o_full(x), (j - 1) & 1);
        tc_fence_after();
      }
      float mx = -INFINITY;
      float l_tile = tile_exp(n_valid, m_used, mx);
      if (spec) {
        const float m_new = fmaxf(m_used, mx);
        const bool need = (m_new - m_used) * c > 8.f;
        if (__any_sync(0xffffffffu, need)) {
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
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}
