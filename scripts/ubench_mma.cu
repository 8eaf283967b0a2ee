// tcgen05.mma issue / completion cost on sm_100a as a function of the tile width N, the number of independent
// accumulators, the A operand source (shared memory vs TMEM) and the number of CTAs sharing the SM's tensor pipe.
// Written after the attention phase trace (profiles/r01_attn_fwd_phase_trace.txt) showed the MMA-issuing thread spending
// ~100 cycles per 128x48x16 step and ~220 per 128x128x16 step: this isolates that cost from everything else.
// Round-1 result for the lane-0 form: a constant 184 cycles per MMA (profiles/r01_ubench_mma_issue_b200.txt) - the
// issuing thread's ELECT..BRA.U.ANY loop, not the pipe.  The elect.sync form (second table) has NOT been run yet.
// Operands are uninitialised shared memory - only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Isid_lsg_b200/csrc -Iinclude -o ubench_mma scripts/ubench_mma.cu
#include "tc_common.cuh"

using namespace sidlsg;

struct MmaCfg { int N, nacc, a_tmem, reps, tcols, a_mn, b_mn; };

template <bool ELECT_UNUSED>
__global__ void __launch_bounds__(128, 2) mma_issue_kernel(long long* out, MmaCfg c) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 16384, bar = base + 16384 + 32768, slot = bar + 8;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) {
    if (c.tcols == 512) tmem_alloc<512>(slot); else tmem_alloc<256>(slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (warp == 0) {
   if (role_leader()) {
    const uint32_t idesc = make_idesc(c.N, c.a_mn, c.b_mn);
    const uint32_t a_t = tmem + c.nacc * c.N;            // A-in-TMEM region behind the accumulators (32 columns = K 64)
    const long long t0 = clock64();
    int i = 0;
    for (int r = 0; r < c.reps; ++r) {
#pragma unroll
      for (int k = 0; k < 4; ++k, ++i) {
        const uint32_t d = tmem + (i % c.nacc) * c.N;
        // K-major: +32 B per 16-element K step; MN-major: 64-wide MN atoms 16 KB apart, +2048 B per 16-row K step
        const uint64_t bd = c.b_mn ? make_desc(b_smem + k * 2048, 1024, 16384) : make_desc(b_smem + k * 32, 1024, 0);
        const uint64_t ad = c.a_mn ? make_desc(a_smem + k * 2048, 1024, 16384) : make_desc(a_smem + k * 32, 1024, 0);
        if (c.a_tmem) tc_mma_bf16_ta(d, a_t + k * 8, bd, idesc, 1);
        else tc_mma_bf16(d, ad, bd, idesc, 1);
      }
    }
    const long long t1 = clock64();
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
   }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (c.tcols == 512) tmem_dealloc<512>(tmem); else tmem_dealloc<256>(tmem);
  }
}

template <bool ELECT>
static int sweep(long long* out, int smem) {
  cudaFuncSetAttribute(mma_issue_kernel<ELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("---- issuing thread = %s\n", ELECT ? "if (elect_one())  [elect.sync]" : "if (lane == 0)");
  printf("%4s %5s %6s %8s | %10s %12s   (cycles per 128xNx16 MMA; nominal math = N/2)\n", "N", "nacc", "A", "CTAs/SM",
         "issue", "issue+drain");
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int a_tmem = 0; a_tmem <= 1; ++a_tmem)
      for (int N : {48, 64, 128, 256})
        for (int nacc : {1, 2, 4}) {
          const int tcols = ctas == 2 ? 256 : 512;
          if (nacc * N + 32 > tcols) continue;
          MmaCfg c{N, nacc, a_tmem, 256, tcols, 0, 0};
          long long h[2 * 296];
          for (int rep = 0; rep < 2; ++rep) {
            mma_issue_kernel<ELECT><<<148 * ctas, 128, smem>>>(out, c);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(h, out, sizeof(long long) * 2 * 148 * ctas, cudaMemcpyDeviceToHost);
          double iss = 0, tot = 0;
          for (int i = 0; i < 148 * ctas; ++i) { iss += (double)h[2 * i]; tot += (double)h[2 * i + 1]; }
          const double n = 148.0 * ctas * c.reps * 4;
          printf("%4d %5d %6s %8d | %10.1f %12.1f\n", N, nacc, a_tmem ? "TMEM" : "smem", ctas, iss / n, tot / n);
        }
  return 0;
}

// operand majorness (the attention backward feeds dV / dK / dQ with MN-major operands: the same Q / dO / K / dS^T tiles
// re-read transposed): cycles per MMA, one CTA per SM, elect.sync issue
static int sweep_major(long long* out, int smem) {
  cudaFuncSetAttribute(mma_issue_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("---- operand majorness (A smem unless TMEM), 1 CTA/SM, elect.sync\n");
  printf("%4s %6s %6s %6s | %10s %12s\n", "N", "A", "a_maj", "b_maj", "issue", "issue+drain");
  for (int N : {48, 64, 128})
    for (int a_tmem = 0; a_tmem <= 1; ++a_tmem)
      for (int a_mn = 0; a_mn <= (a_tmem ? 0 : 1); ++a_mn)
        for (int b_mn = 0; b_mn <= 1; ++b_mn) {
          MmaCfg c{N, 1, a_tmem, 256, 512, a_mn, b_mn};
          long long h[2 * 148];
          for (int rep = 0; rep < 2; ++rep) {
            mma_issue_kernel<true><<<148, 128, smem>>>(out, c);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          cudaMemcpy(h, out, sizeof(long long) * 2 * 148, cudaMemcpyDeviceToHost);
          double iss = 0, tot = 0;
          for (int i = 0; i < 148; ++i) { iss += (double)h[2 * i]; tot += (double)h[2 * i + 1]; }
          const double n = 148.0 * c.reps * 4;
          printf("%4d %6s %6s %6s | %10.1f %12.1f\n", N, a_tmem ? "TMEM" : "smem", a_mn ? "MN" : "K", b_mn ? "MN" : "K", iss / n, tot / n);
        }
  return 0;
}

int main(int argc, char** argv) {
  long long* out;
  cudaMalloc(&out, 2 * 296 * sizeof(long long));
  const int smem = 16384 + 32768 + 1024 + 64;
  if (argc > 1 && argv[1][0] == 'm') return sweep_major(out, smem);
  if (sweep<true>(out, smem)) return 1;
  if (sweep_major(out, smem)) return 1;
  return 0;
}
