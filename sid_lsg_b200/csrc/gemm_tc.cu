// tcgen05 / TMEM / TMA GEMM and implicit-GEMM conv3x3 for the bf16 mode of the SiD-LSG step (sm_100a only).
//
// One persistent, warp-specialised kernel covers every dense contraction of the UNet forward/backward:
//   linear fwd      C[M,N]   = A[M,K]  W[N,K]^T          A K-major,   B K-major
//   linear dgrad    dX[M,K]  = dY[M,N] W[N,K]            A K-major,   B MN-major (2-D map over W)
//   linear wgrad    dW[N,K] += dY[M,N]^T X[M,K]          A MN-major,  B MN-major, split-K + fp32 red.add
//   conv3x3 fwd     NHWC implicit GEMM: A = 4-D TMA boxes of x shifted per tap (zero padding = TMA OOB fill),
//                   B = weights [Cout][tap][Cin] K-major
//   conv3x3 dgrad   A = 4-D boxes of dY shifted by -tap, B = 3-D map (Cin, tap, Cout) MN-major
//   conv3x3 wgrad   A = dY MN-major, B = 4-D boxes of x shifted per tap (MN-major), one tap per N tile, split-K
// (reference call sites: every Conv2d / Linear of the UNet behind training/sid_sd_util.py:184,245,263 and their
// autograd, training/sid_training_loop.py:450,533).
//
// Roles (384 threads): warp 0 = TMA producer (one elected lane), warp 1 = tcgen05.mma issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4-11 = epilogue (TMEM -> registers -> bias / timestep-row / residual -> global;
// two warpgroups take alternate 32-column chunks of the tile).
// 4-stage smem ring (128x64 A + up to 256x64 B, SWIZZLE_128B), two 256-column fp32 accumulators in TMEM so the
// epilogue of tile i overlaps the main loop of tile i+1.  Tiles are 128 x block_n with a narrower last N tile
// (UMMA N is a run-time field of the instruction descriptor), so N = 320 = 192 + 128 wastes nothing.
#include "tc_common.cuh"
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace sidlsg {

enum { A_K2D = 0, A_MN2D = 1, A_CONV = 2 };
enum { B_K2D = 0, B_MN2D = 1, B_W3D = 2, B_CONV = 3 };

constexpr int TC_STAGES = 4;          // ring depth with 256-column tiles; narrower tiles get more (p.stages, <= TC_MAX_STAGES)
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;        // 16 KB
constexpr int TC_B_BYTES = 256 * TC_BK * 2;          // 32 KB (max block_n)
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_OUT_BYTES = TC_BM * 64 * 2;         // 16 KB: one [128 rows][64 bf16] output chunk (SWIZZLE_128B)
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 2 * TC_OUT_BYTES + 1024 /*align*/ + 256 /*barriers*/ +
                              1024 /*bias*/;
constexpr int TC_THREADS = 384;   // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue

struct TcParams {
  int M, N;                 // output extent (rows, columns of the logical GEMM)
  int block_n;              // N tile (multiple of 16; multiple of 64 for MN-major B)
  int m_tiles, n_tiles, splits;
  int kb_total;             // 64-deep reduction blocks
  int a_mode, b_mode;
  // convolution geometry (A_CONV / B_CONV / B_W3D)
  int H, W;                 // image extent of the tensor the boxes are cut from
  int cchunks;              // 64-channel chunks per tap in the reduction (conv fwd / dgrad)
  int flip;                 // dgrad: sample at -tap offset
  int cstride;              // convolution stride (1 or 2): TMA traversal stride over W and H
  int n_tiles_per_tap;      // conv wgrad: N tiles per tap
  int cin;                  // conv wgrad: channels per tap (column offset = tap*cin + n_in)
  // epilogue
  void* c;
  long ldc;
  int out_f32;
  const float* bias;
  const void* res;          // same dtype as c
  long ldr;
  const float* rowvec;
  int rows_per_vec;
  float alpha;
  int atomic;               // fp32 red.add (split-K / gradient accumulation)
  int bm2;                  // 256-row tiles: two 128-row A tiles share every B tile (two accumulators, single-buffered)
  int stages;               // smem ring depth: floor(192 KB / (16 KB + B tile bytes)), 4..8
  int stage_bytes;          // 16 KB A tile + B tile (block_n rows x 128 B, or 64-column boxes x 8 KB for MN-major B)
  int tma_store;            // bf16 output leaves through smem staging + TMA bulk stores (tmC) instead of per-row stores
  long long* trace;         // phase trace of CTA 1 (sidlsg_debug_gemm_trace; null in production)
  // batched dense GEMMs (attention score / value contractions per (batch, head)): 4-D operand maps
  // (inner, outer, nb2, nb1); tile index = ((batch * m_tiles) + m) * n_tiles + n
  int batched, nb2;
  long c_sb1, c_sb2;
};

// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  tmem_ld16_nowait(taddr, r);
  tmem_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct TileInfo {
  int m0, n_in, col0, n_valid, tap, kb0, kb1, b1, b2;
};

__host__ __device__ __forceinline__ TileInfo tile_from_blocks(const TcParams& p, int m_blk, int n_blk, int split, int per) {
  TileInfo t;
  t.b1 = t.b2 = 0;
  if (p.batched) {
    const int bidx = m_blk / p.m_tiles;
    m_blk -= bidx * p.m_tiles;
    t.b1 = bidx / p.nb2;
    t.b2 = bidx - t.b1 * p.nb2;
  }
  t.m0 = m_blk * (p.bm2 ? 2 * TC_BM : TC_BM);
  if (p.b_mode == B_CONV) {
    t.tap = n_blk / p.n_tiles_per_tap;
    t.n_in = (n_blk - t.tap * p.n_tiles_per_tap) * p.block_n;
    t.col0 = t.tap * p.cin + t.n_in;
    t.n_valid = min(p.block_n, p.cin - t.n_in);
  } else {
    t.tap = 0;
    t.n_in = n_blk * p.block_n;
    t.col0 = t.n_in;
    t.n_valid = min(p.block_n, p.N - t.n_in);
  }
  t.kb0 = split * per;
  t.kb1 = min(p.kb_total, t.kb0 + per);
  return t;
}

// Walks the tiles blockIdx.x, + gridDim.x, ... of the persistent grid.  Tile index -> (split, m block, n block) costs
// three integer divisions (~100 clk of dependent SASS each); the phase trace of the K = 320 GEMMs showed ~1,500 clk
// between the last chunk of one tile and the first of the next in the EPILOGUE warps, which bound those tiles.  The
// cursor divides once and then steps (m, n) by (stride / n_tiles, stride % n_tiles) with a carry.
// Phase-trace stamps (scripts/trace_gemm.py) are compiled in only with -DSIDLSG_GEMM_TRACE (SIDLSG_NVCC_EXTRA of build.py):
// even switched off at run time they cost the conv tiles ~1 % (same-box A/B).
#ifdef SIDLSG_GEMM_TRACE
#define SIDLSG_TRACE_ON 1
#else
#define SIDLSG_TRACE_ON 0
#endif
struct TileCursor {
  int tile, m_blk, n_blk, split, dq, dr, per, stride, rows;
  __host__ __device__ __forceinline__ TileCursor(const TcParams& p, int tile0, int stride_) {
    tile = tile0; stride = stride_;
    rows = p.m_tiles * (p.batched ? p.batched : 1);               // m blocks per split (batches folded into m)
    const int base = rows * p.n_tiles;                            // (m, n) tiles per split
    per = (p.kb_total + p.splits - 1) / p.splits;
    // split-K tiles are ordered split-major: the CTAs running at the same time work on the SAME K slab for different
    // (m, n) output tiles, so the operand slabs are fetched from HBM once and shared through L2 (with the split index
    // innermost every concurrent CTA streamed its own slab: 300 MB of DRAM reads for 168 MB of operands)
    split = tile0 / base;
    const int r = tile0 - split * base;
    m_blk = r / p.n_tiles;
    n_blk = r - m_blk * p.n_tiles;
    dq = stride / p.n_tiles;
    dr = stride - dq * p.n_tiles;
  }
  __host__ __device__ __forceinline__ TileInfo info(const TcParams& p) const { return tile_from_blocks(p, m_blk, n_blk, split, per); }
  __host__ __device__ __forceinline__ void advance(const TcParams& p) {
    tile += stride;
    n_blk += dr;
    m_blk += dq;
    if (n_blk >= p.n_tiles) { n_blk -= p.n_tiles; ++m_blk; }
    while (m_blk >= rows) { m_blk -= rows; ++split; }
  }
};

template <bool BM2>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B atoms need 1024-B alignment
  const uint32_t out_base = smem_base + TC_STAGES * TC_STAGE_BYTES;   // 2 x 16 KB output staging (1024-B aligned)
  const uint32_t bar_base = out_base + 2 * TC_OUT_BYTES;
  // barriers: full[4], empty[4], tmem_full[2], tmem_empty[2]; then the TMEM base address word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * TC_MAX_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * TC_MAX_STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC_MAX_STAGES + 4);
  const int NST = p.stages;
  const uint32_t STB = (uint32_t)p.stage_bytes;
  const uint32_t bias_base = bar_base + 256;          // float[256]: bias of the current tile's columns (staged epilogue)
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits * (p.batched ? p.batched : 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8); }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int a_mn = p.a_mode == A_MN2D;
  const int b_mn = p.b_mode != B_K2D;
  const int b_boxes = (p.block_n + 63) >> 6;
  const uint32_t a_bytes = BM2 ? 2 * TC_A_BYTES : TC_A_BYTES;
  const uint32_t stage_tx = a_bytes + (b_mn ? b_boxes * 8192 : p.block_n * 128);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (role_leader()) {
      int stage = 0;
      uint32_t phase = 0;
      const int hw = p.H * p.W;
      int tix = 0;
      for (TileCursor cur(p, blockIdx.x, gridDim.x); cur.tile < total_tiles; cur.advance(p), ++tix) {
        const TileInfo t = cur.info(p);
        if (SIDLSG_TRACE_ON && p.trace && blockIdx.x == 1 && tix < 32) p.trace[tix * 16 + 15] = clock64();
        int ab0 = 0, ah0 = 0, ab1 = 0, ah1 = 0;
        if (p.a_mode == A_CONV) {
          ab0 = t.m0 / hw; ah0 = (t.m0 - ab0 * hw) / p.W;
          ab1 = (t.m0 + TC_BM) / hw; ah1 = (t.m0 + TC_BM - ab1 * hw) / p.W;   // second 128-pixel box (bm2)
        }
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * STB;
          const uint32_t sb = sa + a_bytes;
          const uint32_t fb = full_bar(stage);
          mbar_expect_tx(fb, stage_tx);
          int tap = 0, c0 = kb * TC_BK;
          if (p.a_mode == A_CONV || p.b_mode == B_W3D) { tap = kb / p.cchunks; c0 = (kb - tap * p.cchunks) * TC_BK; }
          // ---- A ----
          if (p.batched) {
            if (p.a_mode == A_K2D) {
              tma_load_4d(&tmA, fb, sa, kb * TC_BK, t.m0, t.b2, t.b1);
            } else {
              tma_load_4d(&tmA, fb, sa, t.m0, kb * TC_BK, t.b2, t.b1);
              tma_load_4d(&tmA, fb, sa + 8192, t.m0 + 64, kb * TC_BK, t.b2, t.b1);
            }
          } else if (p.a_mode == A_K2D) {
            tma_load_2d(&tmA, fb, sa, kb * TC_BK, t.m0);
            if (BM2) tma_load_2d(&tmA, fb, sa + TC_A_BYTES, kb * TC_BK, t.m0 + TC_BM);
          } else if (p.a_mode == A_MN2D) {
            tma_load_2d(&tmA, fb, sa, t.m0, kb * TC_BK);
            tma_load_2d(&tmA, fb, sa + 8192, t.m0 + 64, kb * TC_BK);
          } else {
            int oy = tap / 3 - 1, ox = tap % 3 - 1;
            if (p.flip) { oy = -oy; ox = -ox; }
            tma_load_4d(&tmA, fb, sa, c0, ox, ah0 * p.cstride + oy, ab0);
            if (BM2) tma_load_4d(&tmA, fb, sa + TC_A_BYTES, c0, ox, ah1 * p.cstride + oy, ab1);
          }
          // ---- B ----
          if (p.batched) {
            if (p.b_mode == B_K2D) {
              tma_load_4d(&tmB, fb, sb, kb * TC_BK, t.n_in, t.b2, t.b1);
            } else {
              for (int j = 0; j < b_boxes; ++j) tma_load_4d(&tmB, fb, sb + j * 8192, t.n_in + 64 * j, kb * TC_BK, t.b2, t.b1);
            }
          } else if (p.b_mode == B_K2D) {
            tma_load_2d(&tmB, fb, sb, kb * TC_BK, t.n_in);
          } else if (p.b_mode == B_MN2D) {
            for (int j = 0; j < b_boxes; ++j) tma_load_2d(&tmB, fb, sb + j * 8192, t.n_in + 64 * j, kb * TC_BK);
          } else if (p.b_mode == B_W3D) {
            for (int j = 0; j < b_boxes; ++j) tma_load_3d(&tmB, fb, sb + j * 8192, t.n_in + 64 * j, tap, c0);
          } else {  // B_CONV: reduction block = 64 pixels starting at kb*64
            const int pix0 = kb * TC_BK;
            const int bb0 = pix0 / hw;
            const int bh0 = (pix0 - bb0 * hw) / p.W;
            const int oy = t.tap / 3 - 1, ox = t.tap % 3 - 1;
            for (int j = 0; j < b_boxes; ++j)
              tma_load_4d(&tmB, fb, sb + j * 8192, t.n_in + 64 * j, ox, bh0 * p.cstride + oy, bb0);
          }
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One thread's dependent instruction stream: descriptors are base + running offset (no per-MMA field assembly), the
    // first poll of every wait is inline (phase traces of the attention kernels, r02: ~250 clk per wait and ~55 clk per MMA
    // of uniform-datapath latency otherwise - more than a 128-wide MMA takes to execute).
    if (role_leader()) {
      auto wait = [&](uint32_t bar, uint32_t parity) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok) mbar_wait(bar, parity);
      };
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase[2] = {0, 0};
      const uint64_t ad_base = a_mn ? make_desc(smem_base, 1024, 8192) : make_desc(smem_base, 1024, 0);
      const uint64_t bd_base = b_mn ? make_desc(smem_base + a_bytes, 1024, 8192) : make_desc(smem_base + a_bytes, 1024, 0);
      const uint64_t a_step = a_mn ? 128 : 2, b_step = b_mn ? 128 : 2;      // one K = 16 step, in 16-byte descriptor units
      const uint64_t a2_base = make_desc(smem_base + TC_A_BYTES, 1024, 0);  // bm2: second A tile (K-major only)
      uint64_t stage_off = 0;
      int tix = 0;
      for (TileCursor cur(p, blockIdx.x, gridDim.x); cur.tile < total_tiles; cur.advance(p), ++tix) {
        const TileInfo t = cur.info(p);
        const int n_mma = b_mn ? ((t.n_valid + 63) & ~63) : ((t.n_valid + 15) & ~15);
        const uint32_t idesc = make_idesc(n_mma, a_mn, b_mn);
        long long* tr = (SIDLSG_TRACE_ON && p.trace && blockIdx.x == 1 && tix < 32) ? p.trace + tix * 16 : nullptr;
        wait(tempty_bar(acc), acc_phase[acc] ^ 1);
        tc_fence_after();
        if (tr) tr[12] = clock64();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t accumulate = 0;
#pragma unroll 1
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          const uint64_t ad = ad_base + stage_off, bd = bd_base + stage_off, a2 = a2_base + stage_off;
          const uint32_t eb = empty_bar(stage);
          wait(full_bar(stage), phase);
          tc_fence_after();
          if (tr && kb == t.kb0) tr[13] = clock64();
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            tc_mma_bf16(d_tmem, ad + k * a_step, bd + k * b_step, idesc, accumulate);
            // bm2: rows 128..255 of the tile (second A tile, K-major only) against the SAME B tile -> columns 256.. of TMEM
            if (BM2) tc_mma_bf16(d_tmem + 256, a2 + 2 * k, bd + k * b_step, idesc, accumulate);
            accumulate = 1;
          }
          tc_commit(eb);                 // frees the smem slot when these MMAs retire
          stage_off += STB >> 4;
          if (++stage == NST) { stage = 0; phase ^= 1; stage_off = 0; }
        }
        tc_commit(tfull_bar(acc));       // accumulator complete -> epilogue
        if (tr) tr[14] = clock64();
        acc_phase[acc] ^= 1;
        if (!BM2) acc ^= 1;            // bm2: both TMEM halves belong to one tile (single-buffered)
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps: two groups split the 32-column chunks of a tile) =====================
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int grp = (warp - 4) >> 2;     // column-chunk parity this warp handles
    int acc = 0;
    uint32_t acc_phase[2] = {0, 0};
    uint32_t ochunk = 0;                 // output chunks staged so far (staging buffer = parity)
    float* bias_s = reinterpret_cast<float*>(smem_raw + (bias_base - smem_u32(smem_raw)));
    // tile-invariant half of the fast-path test (the phase trace showed ~850 clk of scalar code per tile between the
    // top of the epilogue loop and the accumulator wait: this predicate, re-evaluated from the constant bank every tile)
    const bool fast_inv = (p.out_f32 ? ((p.ldc & 3) == 0) : ((p.ldc & 7) == 0)) &&
                          ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0) &&
                          (((p.c_sb1 | p.c_sb2) & (p.out_f32 ? 3 : 7)) == 0) &&
                          (!p.res || (((p.ldr & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0))) &&
                          (!p.bias || ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) &&
                          (!p.rowvec || (((p.N & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.rowvec) & 15) == 0)));
    // Staging-buffer hand-over.  Short reductions (<= 8 k-blocks per tile: the K = 320 / 512 GEMMs, whose tiles are bound by
    // THIS loop) use ONE named barrier per 64-column chunk: before it the leader waits until every bulk store issued so far
    // has read its buffer, so after it the next chunk may overwrite the older buffer at once.  Deep reductions (conv3x3,
    // K >= 640) keep two barriers per chunk with the relaxed wait (all but the latest store): their stores queue behind the
    // main loop's TMA loads and the strict wait stalled all eight warps (conv 0.194 -> 0.202 ms, same-box A/B).
    const bool one_bar = ((p.kb_total + p.splits - 1) / p.splits) <= 8;
    int tix = 0;
    for (TileCursor cur(p, blockIdx.x, gridDim.x); cur.tile < total_tiles; cur.advance(p), ++tix) {
      const TileInfo t = cur.info(p);
      const int nhalf = BM2 ? 2 : 1;
      long long* tr = (SIDLSG_TRACE_ON && p.trace && blockIdx.x == 1 && threadIdx.x == 128 && tix < 32) ? p.trace + tix * 16 : nullptr;
      if (tr) tr[0] = clock64();
      for (int half = 0; half < nhalf; ++half) {     // bm2: rows 0-127 (TMEM columns 0..) then rows 128-255 (columns 256..)
      const int m = t.m0 + half * TC_BM + q * 32 + lane;
      const bool row_ok = m < p.M;
      const bool has_work = t.kb1 > t.kb0;
      const uint32_t taddr = tmem_base + (acc + half) * 256 + ((uint32_t)(q * 32) << 16);
      const long crow = (long)m * p.ldc + t.col0 + t.b1 * p.c_sb1 + t.b2 * p.c_sb2;
      const float* rv = (p.rowvec && row_ok) ? p.rowvec + (long)(m / p.rows_per_vec) * p.N + t.col0 : nullptr;
      // fast path: whole 16-column groups, 16-byte aligned rows (every shape of the UNet)
      const bool fast = fast_inv && ((t.n_valid & 15) == 0) && ((t.col0 & 15) == 0);
      // bf16 residual of this thread's columns, fetched BEFORE the accumulator is complete: the DRAM latency of the
      // residual tile hides behind the main loop instead of serialising the epilogue (K = 320 GEMMs are output-bound)
      uint4 rpre[16];
      const bool pre = fast && p.res && !p.out_f32 && row_ok && has_work;
      if (pre) {
        const bf16* rp0 = reinterpret_cast<const bf16*>(p.res) + (long)m * p.ldr + t.col0;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = grp * 32 + ci * 64;
          if (c < t.n_valid) {
            rpre[4 * ci] = *reinterpret_cast<const uint4*>(rp0 + c);
            rpre[4 * ci + 1] = *reinterpret_cast<const uint4*>(rp0 + c + 8);
            if (c + 32 <= t.n_valid) {
              rpre[4 * ci + 2] = *reinterpret_cast<const uint4*>(rp0 + c + 16);
              rpre[4 * ci + 3] = *reinterpret_cast<const uint4*>(rp0 + c + 24);
            }
          }
        }
      }
      const bool ts = fast && p.tma_store;        // uniform over the 8 epilogue warps (depends on the tile only)
      if (ts && p.bias && half == 0) {
        // bias of this tile's columns -> smem while the main loop is still running (the per-chunk __ldg's were the
        // epilogue's largest stall).  Single buffer: every read of the previous tile's bias precedes that tile's last
        // named barrier, every read of this one follows the barrier here.
        const int j = threadIdx.x - 128;
        if (j < t.n_valid) bias_s[j] = __ldg(p.bias + t.col0 + j);
        if (one_bar) named_bar_sync(1, 256);       // (two-barrier mode: the first chunk's buffer barrier separates them)
      }
      if (half == 0) {
        mbar_wait(tfull_bar(acc), acc_phase[acc]);   // (all lanes poll: lane-0 polling + __syncwarp measured 20 % SLOWER on K = 320 N = 2560)
        tc_fence_after();
        if (tr) tr[1] = clock64();
      }
      if (ts) {
        // ---- staged epilogue: 64-column chunks -> bf16 SWIZZLE_128B smem tile -> ONE TMA bulk store per chunk (128-byte
        // L2 requests instead of one 16-byte request per thread per store; rows >= M / columns >= N are clipped by tmC).
        // Warp (q, grp) owns rows q*32 + lane and the grp-th 32 columns of the chunk.
        const bool leader = threadIdx.x == 128;
        bool tail_direct = false;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c64 = ci * 64;
          if (c64 >= t.n_valid) break;
          // a partial chunk in the MIDDLE of a row (tile widths that are not multiples of 64, e.g. N = 320 as 160 + 160)
          // cannot go through the 64-column box: the box would overwrite the neighbour tile's columns.  Those chunks
          // leave through per-row stores; partial chunks at the end of the row are clipped by the tensor map.
          const bool direct = (c64 + 64 > t.n_valid) && (t.col0 + t.n_valid < p.N);
          tail_direct = direct;
          const uint32_t obuf = out_base + (uint32_t)(ochunk & 1) * TC_OUT_BYTES;
          if (!direct) {
            ++ochunk;
            if (!one_bar) {
              if (leader) tma_wait_group_read1();   // the store that used this buffer two chunks ago has read it
              named_bar_sync(1, 256);
            }
          }
          if (tr && half == 0 && ci < 3) tr[2 + 3 * ci] = clock64();
          const int c = c64 + grp * 32;
          if (c < t.n_valid) {
            const bool two = c + 32 <= t.n_valid;   // 32 columns, or a 16-column tail
            uint32_t r[32];
            if (two) tmem_ld32_nowait(taddr + c, r);
            else tmem_ld16_nowait(taddr + c, r);
            tmem_wait_ld();
            if (row_ok) {
              const uint32_t srow = obuf + (uint32_t)(q * 32 + lane) * 128u;
              const int rx = lane & 7;
#pragma unroll
              for (int hblk = 0; hblk < 2; ++hblk) {
                if (hblk == 1 && !two) break;
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[hblk * 16 + i]) * p.alpha;
                const int cc = c + hblk * 16;
                if (p.bias) {
#pragma unroll
                  for (int i = 0; i < 16; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cc + i);
                    v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                  }
                }
                if (rv) {
#pragma unroll
                  for (int i = 0; i < 16; i += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(rv + cc + i));
                    v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                  }
                }
                if (pre) {
                  const uint4 r0 = rpre[4 * ci + 2 * hblk], r1 = rpre[4 * ci + 2 * hblk + 1];
                  const __nv_bfloat162* e0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
                  const __nv_bfloat162* e1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    v[2 * i] += __low2float(e0[i]); v[2 * i + 1] += __high2float(e0[i]);
                    v[8 + 2 * i] += __low2float(e1[i]); v[8 + 2 * i + 1] += __high2float(e1[i]);
                  }
                }
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                  o[i] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                if (direct) {
                  bf16* cp = reinterpret_cast<bf16*>(p.c) + crow + cc;
                  *reinterpret_cast<uint4*>(cp) = make_uint4(o[0], o[1], o[2], o[3]);
                  *reinterpret_cast<uint4*>(cp + 8) = make_uint4(o[4], o[5], o[6], o[7]);
                } else {
                  const int u = grp * 4 + hblk * 2;   // 16-byte unit of the 128-byte row, XOR-swizzled by the row
                  st_shared_v4(srow + (uint32_t)(((u) ^ rx) << 4), o[0], o[1], o[2], o[3]);
                  st_shared_v4(srow + (uint32_t)(((u + 1) ^ rx) << 4), o[4], o[5], o[6], o[7]);
                }
              }
            }
          }
          if (tr && half == 0 && ci < 3) tr[3 + 3 * ci] = clock64();
          if (!direct) {
            fence_proxy_async();                    // generic-proxy smem writes -> visible to the TMA engine
            if (one_bar && leader) tma_wait_group_read0();   // the previous chunk's store (issued a whole chunk ago) has read its
            named_bar_sync(2, 256);                          // buffer: after this barrier the NEXT chunk may overwrite it
            if (leader) {
              tma_store_2d(&tmC, obuf, t.col0 + c64, t.m0 + half * TC_BM);
              tma_commit_group();
            }
          }
          if (tr && half == 0 && ci < 3) tr[4 + 3 * ci] = clock64();
        }
        // a direct tail chunk has no barrier behind it: without this one a fast warp could stage the NEXT tile's bias
        // while a slow warp still reads this tile's
        if (tail_direct && p.bias) named_bar_sync(1, 256);
      } else if (fast) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = grp * 32 + ci * 64;
          if (c >= t.n_valid) break;
          const bool two = c + 32 <= t.n_valid;     // 32 columns, or a 16-column tail
          uint32_t r[32];
          __syncwarp();
          if (two) tmem_ld32_nowait(taddr + c, r);
          else tmem_ld16_nowait(taddr + c, r);
          tmem_wait_ld();
          if (row_ok && has_work) {
#pragma unroll
            for (int hblk = 0; hblk < 2; ++hblk) {
              if (hblk == 1 && !two) break;
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[hblk * 16 + i]) * p.alpha;
              const int cc = c + hblk * 16;
              if (p.bias) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + t.col0 + cc + i));
                  v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                }
              }
              if (rv) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(rv + cc + i));
                  v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
                }
              }
              if (p.out_f32) {
                float* cp = reinterpret_cast<float*>(p.c) + crow + cc;
                if (p.res) {
                  const float* rp = reinterpret_cast<const float*>(p.res) + (long)m * p.ldr + t.col0 + cc;
#pragma unroll
                  for (int i = 0; i < 16; i += 4) {
                    const float4 r4 = *reinterpret_cast<const float4*>(rp + i);
                    v[i] += r4.x; v[i + 1] += r4.y; v[i + 2] += r4.z; v[i + 3] += r4.w;
                  }
                }
                if (p.atomic) {
#pragma unroll
                  for (int i = 0; i < 16; i += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + i), "f"(v[i]), "f"(v[i + 1]),
                                 "f"(v[i + 2]), "f"(v[i + 3]) : "memory");
                } else {
#pragma unroll
                  for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(cp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
              } else {
                bf16* cp = reinterpret_cast<bf16*>(p.c) + crow + cc;
                if (pre) {
                  const uint4 r0 = rpre[4 * ci + 2 * hblk], r1 = rpre[4 * ci + 2 * hblk + 1];
                  const __nv_bfloat162* e0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
                  const __nv_bfloat162* e1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    v[2 * i] += __low2float(e0[i]); v[2 * i + 1] += __high2float(e0[i]);
                    v[8 + 2 * i] += __low2float(e1[i]); v[8 + 2 * i + 1] += __high2float(e1[i]);
                  }
                }
                uint4 o0, o1;
                __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
                __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  h0[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                  h1[i] = __floats2bfloat162_rn(v[8 + 2 * i], v[8 + 2 * i + 1]);
                }
                *reinterpret_cast<uint4*>(cp) = o0;
                *reinterpret_cast<uint4*>(cp + 8) = o1;
              }
            }
          }
        }
      } else if (grp == 0) {
        // generic path (odd widths / unaligned rows): scalar tails, one warpgroup
        for (int c = 0; c < t.n_valid; c += 16) {
          float v[16];
          __syncwarp();
          tmem_ld16(taddr + c, v);
          if (row_ok && has_work) {
            const int nc = min(16, t.n_valid - c);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (i < nc) {
                float x = v[i] * p.alpha;
                if (p.bias) x += __ldg(p.bias + t.col0 + c + i);
                if (rv) x += __ldg(rv + c + i);
                if (p.out_f32) {
                  float* cp = reinterpret_cast<float*>(p.c) + crow + c + i;
                  if (p.res) x += reinterpret_cast<const float*>(p.res)[(long)m * p.ldr + t.col0 + c + i];
                  if (p.atomic) atomicAdd(cp, x);
                  else *cp = x;
                } else {
                  if (p.res) x += __bfloat162float(reinterpret_cast<const bf16*>(p.res)[(long)m * p.ldr + t.col0 + c + i]);
                  reinterpret_cast<bf16*>(p.c)[crow + c + i] = __float2bfloat16_rn(x);
                }
              }
            }
          }
        }
      }
      }   // half
      if (tr) tr[11] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc_phase[acc] ^= 1;
      if (!BM2) acc ^= 1;
    }
  }

  if (threadIdx.x == 128) tma_wait_group_read0();   // bulk stores must have read their staging tiles before smem goes away
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}

// bf16 tensor map, SWIZZLE_128B, inner box 64 elements.  dims/strides innermost first; strides[i] (bytes) for dim i+1.
bool make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, const uint32_t* elem_strides, int f32) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return false; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  // with a traversal stride s, TMA loads ceil(box/s) elements: box counts are given in LOADED elements here
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; es[i] = elem_strides ? elem_strides[i] : 1; bx[i] = box[i] * es[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT && r != CUDA_ERROR_NOT_INITIALIZED) break;
    // this thread (e.g. an autograd engine thread) has no current context for THIS runtime instance yet:
    // a runtime call binds the device's primary context, then retry once
    cudaFree(0);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rank %d dims %llu,%llu box %u,%u stride0 %llu base %p", (int)r, rank,
              (unsigned long long)gd[0], (unsigned long long)gd[1], bx[0], bx[1], (unsigned long long)gs[0], base);
    return false;
  }
  return true;
}

static int g_tc_state = -1;   // -1 unknown, 0 off, 1 on
static int g_num_sms = 148;
int tc_num_sms() { return g_num_sms; }
unsigned tc_wait_hint_ns() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SIDLSG_WAIT_HINT_NS"); v = e ? atoi(e) : 100; if (v < 0) v = 0; }
  return (unsigned)v;
}

bool tc_enabled() {
  if (g_tc_state < 0) {
    const char* e = getenv("SIDLSG_DISABLE_TC");
    int dev = 0;
    cudaDeviceProp prop;
    bool ok = !(e && e[0] == '1') && cudaGetDevice(&dev) == cudaSuccess &&
              cudaGetDeviceProperties(&prop, dev) == cudaSuccess && prop.major == 10 && get_encode() != nullptr;
    if (ok) {
      g_num_sms = prop.multiProcessorCount;
      ok = cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) == cudaSuccess &&
           cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES) == cudaSuccess;
    }
    cudaGetLastError();
    g_tc_state = ok ? 1 : 0;
  }
  return g_tc_state == 1;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int pick_block_n(int n, bool mn_major) {
  // largest tile <= 256 that leaves the least padded work; MN-major tiles come in 64-column TMA boxes
  int step = mn_major ? 64 : 16;
  int best = step, best_cost = 1 << 30;
  for (int bn = 256; bn >= 64; bn -= step) {
    int tiles = (n + bn - 1) / bn;
    int last = n - (tiles - 1) * bn;
    int last_pad = (last + step - 1) / step * step;
    int cost = (tiles - 1) * bn + last_pad;          // columns actually multiplied
    cost += tiles * 8;                               // mild preference for fewer, wider tiles
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// Cost model shared by the tiling choosers ("cycles" of one CTA).  Measured with ncu on the 3x3 convolutions: every
// SM receives ~42 B/clk from L2 when all of them stream (the chip-wide L2 -> SM cap), so a 64-deep k-block of a 128 x w
// tile costs max(MMA = 2w, fill = (16 KB + 128 w) / 42) - fill-bound at every width, wide tiles cheaper per column.
// A 256-row tile (bm2: two A tiles against one B tile) moves (32 KB + 128 w) per 4w MMA cycles: 1.4-1.5x fewer bytes
// per FLOP, but its two accumulators fill TMEM, so its epilogue is not hidden behind the next main loop.
static inline double tile_cost(int w, int kb, bool bm2 = false) {
  const double mma = (bm2 ? 4.0 : 2.0) * w;
  const double fill = ((bm2 ? 32768.0 : 16384.0) + 128.0 * w) / 42.0;
  const double epi = bm2 ? 2.0 * (8.0 * w + 600.0) : 6.0 * w;
  return kb * (mma > fill ? mma : fill) + epi + 1500.0;
}

// Tile width for a [m_tiles x n] output on a persistent grid of g_num_sms CTAs: simulates the kernel's round-robin
// tile -> CTA assignment (tile t = m_blk * n_tiles + n_blk goes to CTA t % grid) and keeps the width with the
// smallest busiest-CTA load.  Two things the padding-only rule of pick_block_n misses: (1) N = 320 as 256 + 64 puts
// every wide tile on the even CTAs and every narrow one on the odd CTAs (the grid size is even) - a 4:1 imbalance on
// the most common GEMM of the 64x64 level - where 160 + 160 is balanced; (2) the 8x8 / 16x16 layers have 16-64 row
// tiles, e.g. M=2048, N=1280: 80 tiles of 256 columns leave 68 SMs idle, 144 tiles of 144 columns do not.
struct GridTiling { int bn; double cost; };
static GridTiling pick_block_n_grid(int n, bool mn_major, long m_tiles, int kb_total, bool bm2 = false) {
  if (m_tiles <= 0 || m_tiles > 2048) {
    const int bn = pick_block_n(n, mn_major);
    return GridTiling{bn, (double)m_tiles * ((n + bn - 1) / bn) * tile_cost(bn, kb_total, bm2) / g_num_sms};
  }
  static std::mutex mu;
  static std::unordered_map<uint64_t, GridTiling> memo;
  const uint64_t key = ((uint64_t)n << 40) ^ ((uint64_t)kb_total << 16) ^ ((uint64_t)m_tiles << 2) ^ (bm2 ? 2u : 0u) ^
                       (mn_major ? 1u : 0u);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = memo.find(key);
    if (it != memo.end()) return it->second;
  }
  const int step = mn_major ? 64 : 16;
  double best_cost = -1;
  int best = 64;
  std::vector<double> load(g_num_sms);
  for (int bn = 256; bn >= 64; bn -= step) {
    const int nt = (n + bn - 1) / bn;
    const int last = n - (nt - 1) * bn;
    const int last_pad = (last + step - 1) / step * step;
    const long tiles = m_tiles * nt;
    const int G = (int)(tiles < g_num_sms ? tiles : g_num_sms);
    const double c_full = tile_cost(bn, kb_total, bm2), c_last = tile_cost(last_pad, kb_total, bm2);
    std::fill(load.begin(), load.end(), 0.0);
    for (long t = 0; t < tiles; ++t) load[t % G] += (t % nt) == nt - 1 ? c_last : c_full;
    double cost = 0;
    for (int i = 0; i < G; ++i) cost = load[i] > cost ? load[i] : cost;
    if (best_cost < 0 || cost < best_cost * 0.98) { best_cost = cost; best = bn; }
  }
  std::lock_guard<std::mutex> lk(mu);
  memo[key] = GridTiling{best, best_cost};
  return memo[key];
}

// SIDLSG_BM2=0 disables the 256-row tiles, =2 forces them wherever the kernel can run them (tests), default = cost model
static int bm2_mode() {
  static int state = -1;
  if (state < 0) {
    const char* e = getenv("SIDLSG_BM2");
    state = (e && e[0] == '0') ? 0 : ((e && e[0] == '2') ? 2 : 1);
  }
  return state;
}

// Split-K tilings (weight gradients: fp32 red.add epilogue).  Chooses the tile width AND the number of K splits by
// simulating the persistent grid: tile t = split * (m_tiles * n_tiles) + m_blk * n_tiles + n_blk runs on CTA t % grid and costs
// tile_cost(w, kb).
// `groups` x ceil(n / bn) N tiles per M tile (groups = 9 taps for the conv weight gradient).  The old rule aimed at
// 2 x SMs tiles and often landed just above it (300 tiles = 3 rounds with 4 CTAs busy in the last one).
struct SplitTiling { int bn, splits; };
static SplitTiling pick_split_tiling(int n, int groups, long m_tiles, int kb_total) {
  static std::mutex mu;
  static std::unordered_map<uint64_t, SplitTiling> memo;
  const uint64_t key = ((uint64_t)n << 40) ^ ((uint64_t)groups << 36) ^ ((uint64_t)m_tiles << 24) ^ (uint64_t)kb_total;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = memo.find(key);
    if (it != memo.end()) return it->second;
  }
  SplitTiling best{pick_block_n(n, true), 1};
  double best_cost = -1;
  std::vector<double> load(g_num_sms);
  int maxs = kb_total / 8;
  if (maxs > 128) maxs = 128;
  if (maxs < 1) maxs = 1;
  for (int bn = 256; bn >= 64; bn -= 64) {
    const int ntg = (n + bn - 1) / bn;
    const int last_pad = (n - (ntg - 1) * bn + 63) / 64 * 64;
    const long n_tiles = (long)groups * ntg;
    const long base = m_tiles * n_tiles;
    if (base > 4096) continue;
    for (int sp = 1; sp <= maxs; ++sp) {
      const long tiles = base * sp;
      if (tiles > 16 * (long)g_num_sms && sp > 1) break;
      const int per = (kb_total + sp - 1) / sp;
      const int G = (int)(tiles < g_num_sms ? tiles : g_num_sms);
      std::fill(load.begin(), load.end(), 0.0);
      for (long t = 0; t < tiles; ++t) {
        const int split = (int)(t / base);
        const long r = t - split * base;
        const int nb = (int)(r % n_tiles);
        const int w = (nb % ntg) == ntg - 1 ? last_pad : bn;
        int kb = kb_total - split * per;
        if (kb > per) kb = per;
        if (kb <= 0) continue;
        load[t % G] += tile_cost(w, kb);
      }
      double cost = 0;
      for (int i = 0; i < G; ++i) cost = load[i] > cost ? load[i] : cost;
      if (best_cost < 0 || cost < best_cost * 0.98) { best_cost = cost; best = SplitTiling{bn, sp}; }
    }
  }
  std::lock_guard<std::mutex> lk(mu);
  memo[key] = best;
  return best;
}

long g_tc_launches = 0;
thread_local int g_last_path = 0;   // 1 = the last GEMM/conv entry point of this thread ran on tcgen05
extern long g_simt_launches;

// SIDLSG_TMA_STORE=0 keeps the per-row global stores of the epilogue (A/B switch for the staged TMA-store epilogue)
static bool tma_store_enabled() {
  static int state = -1;
  if (state < 0) {
    const char* e = getenv("SIDLSG_TMA_STORE");
    state = (e && e[0] == '0') ? 0 : 1;
  }
  return state == 1;
}

// output map of the staged epilogue: bf16 [M rows][N columns], row stride ldc elements, box = 64 columns x 128 rows
static bool make_out_map(CUtensorMap* m, TcParams& p, void* c, long M, long N, long ldc) {
  p.tma_store = 0;
  if (!tma_store_enabled() || p.out_f32 || p.atomic || p.batched || p.splits != 1) return true;
  if ((ldc % 8) || (N % 8) || !aligned16(c)) return true;
  uint64_t d[2] = {(uint64_t)N, (uint64_t)M}, s[1] = {(uint64_t)ldc * 2};
  uint32_t bx[2] = {64, 128};
  if (!make_map(m, c, 2, d, s, bx)) return false;
  p.tma_store = 1;
  return true;
}

// narrow tiles leave room for a deeper ring: K = 320 GEMMs (5 k-blocks per tile) then prefetch more than one whole tile
// ahead while the epilogue of the previous tile drains
static void plan_stages(TcParams& p) {
  const bool b_mn = p.b_mode != B_K2D;
  const int b_bytes = b_mn ? ((p.block_n + 63) >> 6) * 8192 : p.block_n * 128;
  p.stage_bytes = (p.bm2 ? 2 : 1) * TC_A_BYTES + b_bytes;
  int ns = TC_STAGES * TC_STAGE_BYTES / p.stage_bytes;     // what fits in the 192 KB ring
  p.stages = ns > TC_MAX_STAGES ? TC_MAX_STAGES : (ns < 2 ? 2 : ns);
}

static long long* g_gemm_trace = nullptr;   // sidlsg_debug_gemm_trace: device buffer [32 tiles][16 slots] stamped by CTA 1

static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, TcParams& p, cudaStream_t st) {
  long tiles = (long)p.m_tiles * p.n_tiles * p.splits * (p.batched ? p.batched : 1);
  if (tiles <= 0) return SIDLSG_OK;
  p.trace = g_gemm_trace;
  __atomic_add_fetch(&g_tc_launches, 1, __ATOMIC_RELAXED);
  g_last_path = 1;
  int grid = (int)(tiles < g_num_sms ? tiles : g_num_sms);
  plan_stages(p);
  if (p.bm2) gemm_tc_kernel<true><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(ta, tb, tc, p);
  else gemm_tc_kernel<false><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(ta, tb, tc, p);
  return check_launch("gemm_tc");
}

static int pick_splits(long tiles, int kb_total) {
  if (tiles >= g_num_sms) return 1;
  long s = (2L * g_num_sms + tiles - 1) / tiles;
  long maxs = kb_total / 8;   // at least 8 k-blocks (512 reduction elements) per split
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  if (s > 128) s = 128;
  return (int)s;
}

// tile shape / split count of a dense GEMM (everything but pointers): shared by tc_gemm_try and sidlsg_debug_tiling
static void plan_dense(TcParams& p, int M, int N, int K, bool a_k, bool b_mn, int accumulate, int nbatch, int out_dtype) {
  p.m_tiles = (M + TC_BM - 1) / TC_BM;
  p.block_n = pick_block_n(N, b_mn);
  if (accumulate == 0 && nbatch == 1) {
    const int kbt = (K + TC_BK - 1) / TC_BK;
    const GridTiling t1 = pick_block_n_grid(N, b_mn, p.m_tiles, kbt);
    p.block_n = t1.bn;
    // 256-row tiles for the L2-bound (deep-K) GEMMs with enough rows to keep every SM busy
    const int mode = bm2_mode();
    static int min_kb = -1;   // SIDLSG_BM2_MINKB: shallowest reduction (in 64-element blocks) that gets 256-row tiles
    if (min_kb < 0) { const char* e = getenv("SIDLSG_BM2_MINKB"); min_kb = e ? atoi(e) : 16; if (min_kb < 1) min_kb = 1; }
    if (mode && a_k && out_dtype == SIDLSG_BF16 && ((kbt >= min_kb && M >= 512) || (mode == 2 && M >= 256))) {
      const long mt2 = (M + 2 * TC_BM - 1) / (2 * TC_BM);
      const GridTiling t2 = pick_block_n_grid(N, b_mn, mt2, kbt, true);
      if (mode == 2 || t2.cost < t1.cost * 0.95) { p.bm2 = 1; p.block_n = t2.bn; p.m_tiles = (int)mt2; }
    }
  }
  p.kb_total = (K + TC_BK - 1) / TC_BK;
  p.splits = 1;
  if (accumulate == 2 && nbatch == 1 && b_mn) {
    const SplitTiling tl = pick_split_tiling(N, 1, p.m_tiles, p.kb_total);
    p.block_n = tl.bn;
    p.splits = tl.splits;
  }
  p.n_tiles = (N + p.block_n - 1) / p.block_n;
  if (accumulate == 2 && !(nbatch == 1 && b_mn)) p.splits = pick_splits((long)p.m_tiles * p.n_tiles, p.kb_total);
}

// Returns 1 if handled on the tensor cores, 0 if the shape is not eligible (caller falls back to the CUDA-core
// kernel of gemm_simt.cu, which is the same arithmetic), <0 on error.
// nb1 x nb2 independent problems (batch strides in elements; 1 x 1 = plain GEMM).
int tc_gemm_try(const void* a, long a_sm, long a_sk, long a_sb1, long a_sb2, const void* b, long b_sn, long b_sk,
                long b_sb1, long b_sb2, void* c, long ldc, long c_sb1, long c_sb2, const float* bias, const void* res,
                long ldr, const float* rowvec, int rows_per_vec, float alpha, int accumulate, int M, int N, int K,
                int nb1, int nb2, int in_dtype, int out_dtype, cudaStream_t st) {
  if (!tc_enabled() || in_dtype != SIDLSG_BF16) return 0;
  const int nbatch = nb1 * nb2;
  if (M < 64 || N < 16 || K < 64) return 0;
  if (accumulate == 1) return 0;
  if (accumulate == 2 && out_dtype != SIDLSG_F32) return 0;
  const bool a_k = a_sk == 1, a_mn = a_sm == 1 && !a_k;
  const bool b_k = b_sk == 1, b_mn = b_sn == 1 && !b_k;
  if (!(a_k || a_mn) || !(b_k || b_mn)) return 0;
  const long lda = a_k ? a_sm : a_sk, ldb = b_k ? b_sn : b_sk;
  if ((lda % 8) || (ldb % 8) || !aligned16(a) || !aligned16(b)) return 0;
  if (out_dtype == SIDLSG_BF16 && ((ldc % 8) || (N % 8))) return 0;
  if (nbatch > 1) {
    if (accumulate || bias || res || rowvec) return 0;
    if ((a_sb1 % 8) || (a_sb2 % 8) || (b_sb1 % 8) || (b_sb2 % 8)) return 0;
    if ((long)nbatch * ((M + TC_BM - 1) / TC_BM) * N > (1L << 30)) return 0;
  }

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.cstride = 1;
  plan_dense(p, M, N, K, a_k, b_mn, accumulate, nbatch, out_dtype);
  p.a_mode = a_k ? A_K2D : A_MN2D;
  p.b_mode = b_k ? B_K2D : B_MN2D;
  p.c = c; p.ldc = ldc; p.out_f32 = out_dtype == SIDLSG_F32;
  p.bias = bias; p.res = res; p.ldr = ldr; p.rowvec = rowvec; p.rows_per_vec = rows_per_vec > 0 ? rows_per_vec : 1;
  p.alpha = alpha; p.atomic = accumulate == 2;
  p.batched = nbatch > 1 ? nbatch : 0; p.nb2 = nb2; p.c_sb1 = c_sb1; p.c_sb2 = c_sb2;

  CUtensorMap ta, tb;
  // operand maps: (inner, outer[, nb2, nb1]); a batch dimension of extent 1 gets a harmless 16-byte-multiple stride
  auto dense_map = [&](CUtensorMap* m, const void* base, bool kmajor, int rows, long ld, long sb1, long sb2,
                       uint32_t box_rows) -> bool {
    uint64_t d[4] = {kmajor ? (uint64_t)K : (uint64_t)rows, kmajor ? (uint64_t)rows : (uint64_t)K, (uint64_t)nb2, (uint64_t)nb1};
    uint64_t s[3] = {(uint64_t)ld * 2, (uint64_t)(nb2 > 1 ? sb2 : ld) * 2, (uint64_t)(nb1 > 1 ? sb1 : ld) * 2};
    uint32_t bx[4] = {64, kmajor ? box_rows : 64u, 1, 1};
    return make_map(m, base, nbatch > 1 ? 4 : 2, d, s, bx);
  };
  if (!dense_map(&ta, a, a_k, M, lda, a_sb1, a_sb2, 128)) return SIDLSG_ERR_CUDA;
  if (!dense_map(&tb, b, b_k, N, ldb, b_sb1, b_sb2, (uint32_t)p.block_n)) return SIDLSG_ERR_CUDA;
  CUtensorMap tcm = ta;   // placeholder (a valid map) when the staged epilogue is not used
  if (!make_out_map(&tcm, p, c, M, N, ldc)) return SIDLSG_ERR_CUDA;
  int r = launch_tc(ta, tb, tcm, p, st);
  return r == SIDLSG_OK ? 1 : r;
}

// tile shape of the implicit-GEMM convolution (forward / data gradient)
static void plan_conv(TcParams& p, long M, int N, int Kc, bool dgrad, bool allow_bm2 = true) {
  p.m_tiles = (int)((M + TC_BM - 1) / TC_BM);
  const int kbt = 9 * (Kc / 64);
  const GridTiling t1 = pick_block_n_grid(N, dgrad, p.m_tiles, kbt);
  p.block_n = t1.bn;
  const int mode = bm2_mode();
  if (mode && allow_bm2 && M >= 256 && (M % (2 * TC_BM)) == 0) {
    const long mt2 = M / (2 * TC_BM);
    const GridTiling t2 = pick_block_n_grid(N, dgrad, mt2, kbt, true);
    if (mode == 2 || (M >= 512 && t2.cost < t1.cost * 0.95)) { p.bm2 = 1; p.block_n = t2.bn; p.m_tiles = (int)mt2; }
  }
}

// tile geometry for cutting 128 (or 64) consecutive pixels out of [B, H, W, C]: whole image rows only
static bool conv_box(int H, int W, int pixels, uint32_t* wt, uint32_t* ht, uint32_t* nt) {
  if (W > pixels || (pixels % W)) return false;
  int rows = pixels / W;
  if (rows <= H) {
    if (H % rows) return false;
    *wt = W; *ht = rows; *nt = 1;
  } else {
    if (rows % H) return false;
    *wt = W; *ht = H; *nt = rows / H;
  }
  return *wt <= 256 && *ht <= 256 && *nt <= 256;
}

// conv3x3 stride 1: forward (flip=0, weights [N][tap][Kc]) or data gradient (flip=1, x := dy, weights [Kc][tap][N])
int tc_conv3x3_try(const void* x, const void* w, void* y, const float* bias, const void* res, const float* rowvec,
                   int B, int Hi, int Wi, int Kc, int Ho, int Wo, int N, long w_sn, long w_stap, long w_sk, int stride,
                   int up, int transposed, int flip, int accumulate, int in_dtype, int out_dtype, cudaStream_t st) {
  if (!tc_enabled() || in_dtype != SIDLSG_BF16) return 0;
  // fp32 output (plain or atomic-add) serves the fp32-accurate 3 x bf16 mode (split3.cu); the bf16 path never accumulates
  if (out_dtype != SIDLSG_BF16 && out_dtype != SIDLSG_F32) return 0;
  if (accumulate == 1 || (accumulate == 2 && out_dtype != SIDLSG_F32)) return 0;
  if ((stride != 1 && stride != 2) || up != 1 || transposed) return 0;
  if ((Kc % 64) || (N % 8) || !aligned16(x) || !aligned16(w)) return 0;
  const bool fwd = !flip && w_sk == 1 && w_stap == Kc && w_sn == 9L * Kc;
  const bool dgrad = flip && w_sn == 1 && w_stap == N && w_sk == 9L * N;
  if (!fwd && !dgrad) return 0;
  if (dgrad && ((N % 64) || stride != 1)) return 0;
  if (stride == 1 && (Hi != Ho || Wi != Wo)) return 0;
  if (stride == 2 && (Hi != 2 * Ho || Wi != 2 * Wo)) return 0;
  uint32_t wt, ht, nt;
  if (!conv_box(Ho, Wo, 128, &wt, &ht, &nt)) return 0;
  const long M = (long)B * Ho * Wo;
  if (M < 128) return 0;

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = N;
  plan_conv(p, M, N, Kc, dgrad, out_dtype == SIDLSG_BF16);   // 256-row tiles: bf16 outputs only (as plan_dense)
  p.n_tiles = (N + p.block_n - 1) / p.block_n;
  p.splits = 1;
  p.cchunks = Kc / 64;
  p.kb_total = 9 * p.cchunks;
  p.a_mode = A_CONV;
  p.b_mode = fwd ? B_K2D : B_W3D;
  p.H = Ho; p.W = Wo; p.flip = flip; p.cstride = stride;
  p.c = y; p.ldc = N; p.out_f32 = out_dtype == SIDLSG_F32; p.atomic = accumulate == 2;
  p.bias = bias; p.res = res; p.ldr = N; p.rowvec = rowvec; p.rows_per_vec = Ho * Wo; p.alpha = 1.f;

  CUtensorMap ta, tb;
  {
    uint64_t d[4] = {(uint64_t)Kc, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)B};
    uint64_t s[3] = {(uint64_t)Kc * 2, (uint64_t)Wi * Kc * 2, (uint64_t)Hi * Wi * Kc * 2};
    uint32_t bx[4] = {64, wt, ht, nt};
    uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    if (!make_map(&ta, x, 4, d, s, bx, es)) return SIDLSG_ERR_CUDA;
  }
  if (fwd) {
    uint64_t d[2] = {(uint64_t)9 * Kc, (uint64_t)N}, s[1] = {(uint64_t)9 * Kc * 2};
    uint32_t bx[2] = {64, (uint32_t)p.block_n};
    if (!make_map(&tb, w, 2, d, s, bx)) return SIDLSG_ERR_CUDA;
  } else {
    // weights physically [Kc = cout][tap][N = cin]: dims (cin, tap, cout)
    uint64_t d[3] = {(uint64_t)N, 9, (uint64_t)Kc}, s[2] = {(uint64_t)N * 2, (uint64_t)9 * N * 2};
    uint32_t bx[3] = {64, 1, 64};
    if (!make_map(&tb, w, 3, d, s, bx)) return SIDLSG_ERR_CUDA;
  }
  CUtensorMap tcm = ta;
  if (!make_out_map(&tcm, p, y, M, N, N)) return SIDLSG_ERR_CUDA;
  int r = launch_tc(ta, tb, tcm, p, st);
  return r == SIDLSG_OK ? 1 : r;
}

// dw[co][tap][ci] += sum_pix dy[pix][co] * window(x)[pix][tap][ci]   (stride 1; dw fp32 dense [Cout][9][Cin])
int tc_conv3x3_wgrad_try(const void* x, const void* dy, float* dw, int B, int Hi, int Wi, int Cin, int Ho, int Wo,
                         int Cout, long dw_sco, long dw_stap, long dw_sci, int stride, int up, int accumulate,
                         int in_dtype, cudaStream_t st) {
  if (!tc_enabled() || in_dtype != SIDLSG_BF16) return 0;
  if ((stride != 1 && stride != 2) || up != 1) return 0;
  if (stride == 1 && (Hi != Ho || Wi != Wo)) return 0;
  if (stride == 2 && (Hi != 2 * Ho || Wi != 2 * Wo)) return 0;
  if ((Cin % 64) || (Cout % 8) || Cout < 64 || !aligned16(x) || !aligned16(dy) || !aligned16(dw)) return 0;
  if (!(dw_sci == 1 && dw_stap == Cin && dw_sco == 9L * Cin)) return 0;
  uint32_t wt, ht, nt;
  if (!conv_box(Ho, Wo, 64, &wt, &ht, &nt)) return 0;
  const long npix = (long)B * Ho * Wo;
  if (npix < 64 || (npix % 64)) return 0;
  if (!accumulate) cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * 9 * Cin, st);

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = Cout; p.N = 9 * Cin;
  p.m_tiles = (Cout + TC_BM - 1) / TC_BM;
  p.kb_total = (int)(npix / 64);
  const SplitTiling tl = pick_split_tiling(Cin, 9, p.m_tiles, p.kb_total);
  p.block_n = tl.bn;
  p.splits = tl.splits;
  p.n_tiles_per_tap = (Cin + p.block_n - 1) / p.block_n;
  p.n_tiles = 9 * p.n_tiles_per_tap;
  p.cin = Cin;
  p.a_mode = A_MN2D;
  p.b_mode = B_CONV;
  p.H = Ho; p.W = Wo; p.cstride = stride;
  p.c = dw; p.ldc = 9L * Cin; p.out_f32 = 1; p.alpha = 1.f; p.atomic = 1; p.rows_per_vec = 1;

  CUtensorMap ta, tb;
  {
    uint64_t d[2] = {(uint64_t)Cout, (uint64_t)npix}, s[1] = {(uint64_t)Cout * 2};
    uint32_t bx[2] = {64, 64};
    if (!make_map(&ta, dy, 2, d, s, bx)) return SIDLSG_ERR_CUDA;
  }
  {
    uint64_t d[4] = {(uint64_t)Cin, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)B};
    uint64_t s[3] = {(uint64_t)Cin * 2, (uint64_t)Wi * Cin * 2, (uint64_t)Hi * Wi * Cin * 2};
    uint32_t bx[4] = {64, wt, ht, nt};
    uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    if (!make_map(&tb, x, 4, d, s, bx, es)) return SIDLSG_ERR_CUDA;
  }
  int r = launch_tc(ta, tb, ta, p, st);
  return r == SIDLSG_OK ? 1 : r;
}

}  // namespace sidlsg

// Host-only diagnostic (no GPU needed): the tiling the tensor-core path would choose.
//   kind 0 linear fwd (A, B K-major), 1 linear dgrad (B MN-major), 2 linear wgrad (A, B MN-major, split-K),
//   3 conv3x3 fwd, 4 conv3x3 dgrad (M = pixels, N = output channels, K = input channels), 5 conv3x3 wgrad (M = Cout,
//   N = Cin, K = pixels).  out[8] = {bm2, block_n, m_tiles, n_tiles, splits, stages, stage_bytes, kb_total}.
extern "C" int sidlsg_debug_tiling(int kind, long M, int N, long K, int* out) {
  using namespace sidlsg;
  if (kind < 0 || kind > 5 || M <= 0 || N <= 0 || K <= 0) { set_error("sidlsg_debug_tiling: bad arguments"); return SIDLSG_ERR_ARG; }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = N;
  if (kind <= 2) {
    const bool a_k = kind != 2, b_mn = kind != 0;
    plan_dense(p, (int)M, N, (int)K, a_k, b_mn, kind == 2 ? 2 : 0, 1, kind == 2 ? SIDLSG_F32 : SIDLSG_BF16);
    p.a_mode = a_k ? A_K2D : A_MN2D;
    p.b_mode = b_mn ? B_MN2D : B_K2D;
  } else if (kind <= 4) {
    plan_conv(p, M, N, (int)K, kind == 4);
    p.n_tiles = (N + p.block_n - 1) / p.block_n;
    p.splits = 1;
    p.kb_total = 9 * ((int)K / 64);
    p.a_mode = A_CONV;
    p.b_mode = kind == 4 ? B_W3D : B_K2D;
  } else {
    p.m_tiles = (int)((M + TC_BM - 1) / TC_BM);
    p.kb_total = (int)(K / 64);
    const SplitTiling tl = pick_split_tiling(N, 9, p.m_tiles, p.kb_total);
    p.block_n = tl.bn;
    p.splits = tl.splits;
    p.n_tiles = 9 * ((N + p.block_n - 1) / p.block_n);
    p.a_mode = A_MN2D;
    p.b_mode = B_CONV;
  }
  plan_stages(p);
  out[0] = p.bm2; out[1] = p.block_n; out[2] = p.m_tiles; out[3] = p.n_tiles; out[4] = p.splits; out[5] = p.stages;
  out[6] = p.stage_bytes; out[7] = p.kb_total;
  return SIDLSG_OK;
}

// diagnostics: out[0] = tcgen05 GEMM/conv launches, out[1] = CUDA-core GEMM/conv launches (host memory)
extern "C" int sidlsg_last_path() { return sidlsg::g_last_path; }
extern "C" int sidlsg_counters(long* out) {
  out[0] = sidlsg::g_tc_launches;
  out[1] = sidlsg::g_simt_launches;
  return SIDLSG_OK;
}

// Debug: GEMM / conv launches after this call stamp clock64 at the phase boundaries of CTA 1's first 32 tiles into
// `trace` (device, 32 x 16 int64; scripts/trace_gemm.py); null switches the stamps off again.
extern "C" int sidlsg_debug_gemm_trace(void* trace) {
  sidlsg::g_gemm_trace = (long long*)trace;
  return SIDLSG_OK;
}

// Debug / test: the tiles CTA `cta` of a `grid`-CTA persistent launch visits, walked on the HOST with the kernel's own
// TileCursor.  geom[9] = {m_tiles, n_tiles, splits, batched, nb2, kb_total, block_n, N, bm2}; out receives up to
// max_tiles records of 8 ints {tile, m0, col0, n_valid, kb0, kb1, b1, b2}; returns the number of tiles (tests compare it
// with the division-per-tile decode the cursor replaced).
extern "C" int sidlsg_debug_tile_walk(const int* geom, int cta, int grid, int max_tiles, int* out) {
  using namespace sidlsg;
  if (!geom || !out || grid <= 0 || cta < 0 || cta >= grid || geom[0] <= 0 || geom[1] <= 0 || geom[2] <= 0 || geom[6] <= 0) {
    set_error("sidlsg_debug_tile_walk: bad arguments");
    return SIDLSG_ERR_ARG;
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.m_tiles = geom[0]; p.n_tiles = geom[1]; p.splits = geom[2]; p.batched = geom[3]; p.nb2 = geom[4] > 0 ? geom[4] : 1;
  p.kb_total = geom[5]; p.block_n = geom[6]; p.N = geom[7]; p.bm2 = geom[8];
  p.b_mode = B_K2D;
  const int total = p.m_tiles * p.n_tiles * p.splits * (p.batched ? p.batched : 1);
  int n = 0;
  for (TileCursor cur(p, cta, grid); cur.tile < total && n < max_tiles; cur.advance(p), ++n) {
    const TileInfo t = cur.info(p);
    int* o = out + 8 * n;
    o[0] = cur.tile; o[1] = t.m0; o[2] = t.col0; o[3] = t.n_valid; o[4] = t.kb0; o[5] = t.kb1; o[6] = t.b1; o[7] = t.b2;
  }
  return n;
}
