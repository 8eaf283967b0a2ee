// PTX wrappers shared by the tcgen05 kernels (gemm_tc.cu, attention_tc.cu): mbarrier, TMA, tcgen05.mma / ld / st,
// shared-memory matrix descriptors (SWIZZLE_128B) and the bf16 instruction descriptor.  sm_100a only.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace sidlsg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  // (not unrolled: the compiler otherwise emits ~30 copies of the poll per wait and the kernels outgrow the 32 KB L1.5 I-cache)
#pragma unroll 1
  for (long it = 0; it < (1L << 26); ++it) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  printf("sidlsg: mbarrier wait timed out (block %d,%d thread %d bar %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar);
  __trap();
}
// same, with a suspend-time hint (ns): the polling thread is parked by the hardware for up to `hint_ns` per try instead
// of re-issuing try_wait + branch every ~40 cycles (half of all warp samples of the attention kernels were such polls,
// competing for issue slots with the softmax warps of the same scheduler); hint_ns = 0 = plain polling
__device__ __forceinline__ void mbar_wait_h(uint32_t hint_ns, uint32_t bar, uint32_t parity) {
  if (hint_ns == 0) { mbar_wait(bar, parity); return; }
  uint32_t ok = 0;
#pragma unroll 1
  for (long it = 0; it < (1L << 24); ++it) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(hint_ns) : "memory");
    if (ok) return;
  }
  printf("sidlsg: mbarrier wait timed out (block %d,%d thread %d bar %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar);
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// 1-D bulk copy global -> shared (bytes % 16 == 0, both addresses 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// smem tile -> global with an fp32 add performed by the TMA unit / L2 (bulk reduction; out-of-bounds elements are clipped)
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// smem tile -> global (plain store; out-of-bounds rows / columns of the box are clipped by the tensor map)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_wait_group_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// One lane of a fully converged warp (elect.sync).  As the CONDITION of the role branch it tells nvcc that exactly one
// thread is active inside: with `if (lane == 0)` every tcgen05.mma / TMA instruction of the branch is wrapped in an
// "ELECT ... BRA.U.ANY" loop over the possibly-many active lanes (6 extra instructions and a branch per MMA; the MMA
// issuer then spends 100-220 cycles per MMA - scripts/ubench_mma.cu measures 184 - and caps the pipe at one 128 x 256
// x 16 MMA per 184 cycles), with `if (elect_one())` the UTCHMMAs are issued back to back (checked in SASS).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\nselp.u32 %0, 1, 0, e;\n}" : "=r"(pred));
  return pred != 0;
}
// role leader of a single-thread role (TMA producer, MMA issuer).  Validated on B200 in round 2
// (profiles/r02_ubench_mma_issue_b200.txt: 81-171 cycles per MMA against a flat 184 for the `lane == 0` form, which is gone).
__device__ __forceinline__ bool role_leader() { return elect_one(); }

// named barrier among `n` threads (ids 1.. ; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// same with the A operand read from TMEM (lane = row, two bf16 K elements per 32-bit column): no shared-memory read
// for A, which is what bounds MMAs with a narrow N (the P V product of attention: 4 KB of A per 1.5 KB of B)
__device__ __forceinline__ void tc_mma_bf16_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem_addr) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem_addr), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.wait::ld that also "defines" the 32 registers of an earlier tmem_ld32_nowait for the compiler: code reading r[]
// cannot be scheduled above this wait (needed when another load is issued between the load of r and its wait, i.e. for
// software-pipelined TMEM reads; with the plain wait directly behind the load the ordering is implicit)
__device__ __forceinline__ void tmem_wait_ld_dep32(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16 consecutive fp32 columns of this thread's TMEM lane (no wait: pair with tmem_wait_ld)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// shared-memory matrix descriptor, SWIZZLE_128B, descriptor version 1 (sm_100).
//   K-major operand  : rows of 64 bf16 (128 B); sbo = 1024 (8 rows), lbo unused (0); +32 B per 16-element K step
//   MN-major operand : 64 MN elements (128 B) per K row; sbo = 1024 (8 K rows), lbo = bytes between 64-wide MN
//                      atoms; +2048 B per 16-row K step
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// instruction descriptor: bf16 x bf16 -> fp32, M = 128, N = n (multiple of 16)
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                      // C format fp32
  d |= 1u << 7;                      // A format bf16
  d |= 1u << 10;                     // B format bf16
  d |= (uint32_t)(a_mn ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn ? 1 : 0) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(128 >> 4) << 24;
  return d;
}

// byte offset of element (row, col) inside a K-major SWIZZLE_128B tile made of [rows x 64]-element chunks of
// `chunk_bytes` each (what TMA writes and what make_desc(.., 1024, 0) describes)
__device__ __forceinline__ uint32_t sw128_offset(int row, int col, uint32_t chunk_bytes) {
  int chunk = col >> 6, within = col & 63;
  int unit = (within >> 3) ^ (row & 7);
  return chunk * chunk_bytes + row * 128 + unit * 16 + (within & 7) * 2;
}

// host: bf16 tensor map, SWIZZLE_128B; dims/box innermost first; strides_bytes[i] is the stride of dim i+1
// box[] counts LOADED elements per dim; elem_strides (optional) = TMA traversal stride per dim
// f32 = 1 builds an fp32 map (inner box <= 32 elements for the 128-byte swizzle)
bool make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, const uint32_t* elem_strides = nullptr, int f32 = 0);
bool tc_enabled();
int tc_num_sms();
// SIDLSG_WAIT_HINT_NS (default 100; 0 = plain polling): suspend-time hint of the mbarrier waits of the TMA / MMA helper warps
unsigned tc_wait_hint_ns();

}  // namespace sidlsg
