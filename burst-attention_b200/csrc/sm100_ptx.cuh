// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (TMEM alloc / mma / ld / st / commit) and the UMMA descriptors.
// Hand-written for this project; bit layouts follow the PTX ISA 8.7 tables for
// tcgen05 shared-memory / instruction descriptors.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ba {

#define BA_DEVICE __device__ __forceinline__

// ------------------------------------------------------------------ misc
BA_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
BA_DEVICE uint32_t lane_id() { return threadIdx.x & 31; }
BA_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
BA_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
BA_DEVICE float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------ mbarrier
BA_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
BA_DEVICE void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
BA_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
BA_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
BA_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps after ~30 s (-> CUDA error surfaced to the
// host) instead of hanging the GPU box.
BA_DEVICE uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
BA_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
#ifdef BA_NO_WATCHDOG
  while (!mbar_try_wait(bar, parity)) {
  }
#else
  const uint64_t t0 = global_ns();
#pragma unroll 1
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 1024; ++i)
      if (mbar_try_wait(bar, parity)) return;
    if (global_ns() - t0 > 30000000000ull) break;
  }
  printf("ba: mbarrier watchdog block(%d,%d,%d) thread %d bar@%u parity %u\n", blockIdx.x, blockIdx.y,
         blockIdx.z, threadIdx.x, smem_u32(bar), parity);
  __trap();
#endif
}

// ------------------------------------------------------------------ fences
BA_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
BA_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
BA_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
BA_DEVICE void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ TMA
BA_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 4-D tiled load: coordinates innermost-first (d, h, s, b).
BA_DEVICE void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                           int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
      : "memory");
}
// 4-D tiled store smem -> global (bulk async group completion).
BA_DEVICE void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      :
      : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// 4-D tiled reduce-add smem -> global (fp32 add performed by the TMA unit / L2).
BA_DEVICE void tma_reduce_add_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      :
      : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
BA_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
BA_DEVICE void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
BA_DEVICE void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------ TMEM
// Whole-warp collective.  ncols: power of two in [32, 512].
BA_DEVICE void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
}
BA_DEVICE void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
BA_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// MMA completion -> mbarrier arrive (implies tcgen05.fence::before_thread_sync).
BA_DEVICE void umma_commit(uint64_t* bar) {
  if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
BA_DEVICE void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
BA_DEVICE void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TMEM address: bits [31:16] lane, [15:0] column.  A warp may only touch lanes
// 32*(warp_id%4) .. +31; with the 32x32b shape thread t of the warp accesses
// lane (base_lane + t) and N consecutive 32-bit columns.
BA_DEVICE uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) { return base + (lane << 16) + col; }

#define BA_R4(a, i) "=r"(a[i]), "=r"(a[i + 1]), "=r"(a[i + 2]), "=r"(a[i + 3])
#define BA_W4(a, i) "r"(a[i]), "r"(a[i + 1]), "r"(a[i + 2]), "r"(a[i + 3])

BA_DEVICE void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : BA_R4(r, 0), BA_R4(r, 4)
               : "r"(taddr));
}
BA_DEVICE void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : BA_R4(r, 0), BA_R4(r, 4), BA_R4(r, 8), BA_R4(r, 12)
      : "r"(taddr));
}
BA_DEVICE void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : BA_R4(r, 0), BA_R4(r, 4), BA_R4(r, 8), BA_R4(r, 12), BA_R4(r, 16), BA_R4(r, 20), BA_R4(r, 24),
        BA_R4(r, 28)
      : "r"(taddr));
}
BA_DEVICE void tmem_st_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               BA_W4(r, 0), BA_W4(r, 4)
               : "memory");
}
BA_DEVICE void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      BA_W4(r, 0), BA_W4(r, 4), BA_W4(r, 8), BA_W4(r, 12)
      : "memory");
}
BA_DEVICE void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      BA_W4(r, 0), BA_W4(r, 4), BA_W4(r, 8), BA_W4(r, 12), BA_W4(r, 16), BA_W4(r, 20), BA_W4(r, 24),
      BA_W4(r, 28)
      : "memory");
}

// ------------------------------------------------------------------ UMMA descriptors
// Shared-memory matrix descriptor (64 bit):
//  [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//  [32,46) stride-dim byte offset >> 4   [46,48) version = 1 on sm_100
//  [49,52) base offset (0: atoms are 1024-B aligned)   [61,64) swizzle: 2 = 128B
//
// K-major, 128B swizzle (rows of 64 16-bit elements = 128 B, 8-row atom = 1 KiB):
//   SBO = byte distance between 8-row groups (1024 when rows are packed), LBO unused (1).
// MN-major, 128B swizzle (a 128-B line = 64 consecutive MN elements at one K):
//   LBO = byte distance between 64-element MN blocks, SBO = distance between 8-K groups.
BA_DEVICE uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}
// Advance a descriptor's start address by a byte offset (must keep 16-B granularity).
BA_DEVICE uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (bytes >> 4); }

// Instruction descriptor for kind::f16 (fp16/bf16 inputs, fp32 accumulate):
//  [4,6) D format (1 = f32)  [7,10) A format  [10,13) B format (0 = f16, 1 = bf16)
//  [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc(bool bf16, int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((a_mn_major ? 1u : 0u) << 15) |
         ((b_mn_major ? 1u : 0u) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// The MMA / commit wrappers must be called by a CONVERGED warp with warp-uniform
// arguments; one elected lane issues.  (Issuing from an `if (lane == 0)` region
// makes ptxas wrap every UTCHMMA in an ELECT/BRA.U.ANY uniformisation loop.)
// D[tmem] (+)= A[smem] * B[smem]
BA_DEVICE void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]     (A: 128 lanes x K/2 32-bit columns, 2 x 16-bit per column)
BA_DEVICE void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- lean issue path: descriptors as (lo, hi) 32-bit halves --------------------------------
// The MMA issuer is ONE warp; every instruction it spends building 64-bit descriptors is time the
// tensor pipe idles.  hi is a compile-time constant (SBO, version, swizzle); lo = start address >> 4
// | LBO << 16 advances by plain 32-bit adds of compile-time constants.
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (2u << 29);
}
__host__ __device__ constexpr uint32_t desc_lo_lbo(uint32_t lbo_bytes) { return ((lbo_bytes >> 4) & 0x3FFF) << 16; }

BA_DEVICE void umma_ss_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                          uint32_t idesc, uint32_t accumulate) {
  if (elect_one())
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
        :
        : "r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
BA_DEVICE void umma_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                          uint32_t accumulate) {
  if (elect_one())
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n"
        :
        : "r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ------------------------------------------------------------------ CTA pairs (cta_group::2)
// Building blocks for the 2-CTA kernels planned next: a cluster of two CTAs on one
// TPC issues ONE tcgen05.mma with M = 256; each CTA supplies its own 128 rows of A and HALF of B, so
// the shared-memory operand fetch per SM drops by the B half.  Validated by ba_selftest modes 4/5.
BA_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
BA_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
BA_DEVICE void tmem_alloc_2cta(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
}
BA_DEVICE void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
BA_DEVICE void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load issued by either CTA of the pair; the transaction bytes complete on the LEADER's barrier
// (shared::cluster address with the CTA-rank bit 24 cleared).
BA_DEVICE void tma_load_4d_2cta(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
BA_DEVICE void umma_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
BA_DEVICE void umma_ts_2cta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (elect_one()) asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster (release at cluster scope)
BA_DEVICE void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n"
      :
      : "r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// 16-byte store into the shared memory of CTA `cta` of the cluster at the same offset as local `p` (DSMEM)
BA_DEVICE void st_shared_remote_v4(void* p, uint32_t cta, uint4 v) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.v4.b32 [ra], {%2, %3, %4, %5};\n\t}\n"
      :
      : "r"(smem_u32(p)), "r"(cta), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
      : "memory");
}
// generic-proxy writes (any state space, incl. remote shared memory) -> visible to the async proxy (TMA, tcgen05)
BA_DEVICE void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// lean (lo, hi) forms of the pair MMAs
BA_DEVICE void umma_ts_2cta_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                               uint32_t accumulate) {
  if (elect_one())
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}\n"
        :
        : "r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// MMA completion -> arrive on the barrier at this smem offset in BOTH CTAs (mask 0b11)
BA_DEVICE void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
  if (elect_one()) asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ------------------------------------------------------------------ no-swizzle K-major operands, fp32 -> 3 x 16 bit
// K-major, NO swizzle: core matrix = 8 rows x 16 B stored contiguously (128 B); LBO = byte distance between the
// two core matrices of one K = 16 step, SBO = byte distance between 8-row groups.  Used for the rank-1 "fold"
// K steps that add a per-row / per-column fp32 vector to an accumulator on the tensor core: the vector is split
// into three 16-bit parts (hi + lo + lolo, exact to fp32 round-off) that multiply 1.0.
BA_DEVICE uint64_t make_smem_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell); swizzle field = 0: none
  return d;
}
template <bool kBF16>
BA_DEVICE uint32_t to16(float x) {
  if constexpr (kBF16) return __bfloat16_as_ushort(__float2bfloat16_rn(x));
  else return __half_as_ushort(__float2half_rn(x));
}
template <bool kBF16>
BA_DEVICE float from16(uint32_t h) {
  if constexpr (kBF16) return __bfloat162float(__ushort_as_bfloat16(static_cast<unsigned short>(h)));
  else return __half2float(__ushort_as_half(static_cast<unsigned short>(h)));
}
template <bool kBF16>
BA_DEVICE void split3(float x, uint32_t& h0, uint32_t& h1, uint32_t& h2) {
  h0 = to16<kBF16>(x);
  float r = x - from16<kBF16>(h0);
  h1 = to16<kBF16>(r);
  r -= from16<kBF16>(h1);
  h2 = to16<kBF16>(r);
}

__host__ __device__ constexpr uint32_t desc_hi_noswz(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14); }

// ------------------------------------------------------------------ packing
template <bool kBF16>
BA_DEVICE uint32_t pack2(float lo, float hi) {
  if constexpr (kBF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  } else {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
}

}  // namespace ba
