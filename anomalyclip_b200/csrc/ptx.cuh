// Thin inline-PTX wrappers for the sm_100a primitives the kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and fences.
// Everything here is a one-instruction wrapper; no policy lives in this file.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace aclip {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every hot-path kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization
// (common.h: launch_pdl): it lets its successor start as soon as all of its own CTAs are running
// (pdl_launch_dependents, first statement) and does its own set-up -- barrier init, TMEM allocation,
// tensor-map prefetch -- before pdl_wait(), which returns once every predecessor grid has completed
// and flushed.  Nothing produced or still read by a predecessor is touched before pdl_wait().
// Without the launch attribute both instructions are no-ops.
// The wait makes the predecessors' (generic-proxy) stores visible to this grid's generic proxy; the
// TMA engine reads through the async proxy, so a cross-proxy fence follows before any
// cp.async.bulk.tensor may read what a predecessor wrote.  (Required by the memory model; it did NOT
// remove the rare fault seen with programmatic launch, which is why that launch mode is opt-in.)
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;\n\tfence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// ---------------------------------------------------------------- fences
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
      "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05
// TMEM allocation: executed by ONE full warp; the base address lands in *smem_slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Predicated forms for a CONVERGED issuing warp (single-CTA kernels): every lane runs the
// warp-uniform control flow and descriptor arithmetic (which then lives in uniform registers), only
// the lane with `issue` set executes the instruction.  An issuer that branches to one lane first
// spends ~150 cycles of scalar descriptor arithmetic per MMA, more than an N = 64 MMA takes.
__device__ __forceinline__ void mma_f16_ss_if(bool issue, uint32_t d_tmem, uint64_t a_desc,
                                              uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(static_cast<uint32_t>(issue))
      : "memory");
}
__device__ __forceinline__ void mma_commit_if(bool issue, uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(static_cast<uint32_t>(issue))
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has retired.
// (Implies tcgen05.fence::before_thread_sync.)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane i of the warp's
// TMEM quarter, columns [col, col+32)).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31},"
      "[%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- clusters / CTA pairs
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the transaction bytes are
// credited to the barrier at the same offset in the pair's leader (even) CTA.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m,
                                                 uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// Multicast form: the box lands at the same shared-memory offset in every CTA of `cta_mask`, and the
// transaction bytes are credited, per destination CTA, to the barrier at this offset in the leader
// (even) CTA of that destination's pair.
__device__ __forceinline__ void tma_load_3d_pair_mc(void* smem_dst, const CUtensorMap* m,
                                                    uint64_t* bar, int c0, int c1, int c2,
                                                    uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".multicast::cluster [%0], [%1, {%4, %5, %6}], [%2], %3;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(bar) & kPeerBitMask), "h"(cta_mask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(void* smem_dst, const CUtensorMap* m,
                                                 uint64_t* bar, int c0, int c1, int c2, int c3,
                                                 int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T over the CTA pair (M = 256: 128 rows from each CTA's A tile;
// N: each CTA supplies half of the B rows); issued by ONE thread of the leader CTA.
__device__ __forceinline__ void mma_bf16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Predicated forms for a converged issuing warp (all lanes run the warp-uniform control flow, the
// lane with `issue` set executes the instruction; descriptors then live in uniform registers).
__device__ __forceinline__ void mma_bf16_ss_pair_if(bool issue, uint32_t d_tmem, uint64_t a_desc,
                                                    uint64_t b_desc, uint32_t idesc,
                                                    uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(static_cast<uint32_t>(issue))
      : "memory");
}
// Same for 8-bit float operands (kind::f8f6f4, K = 32 per instruction, twice the kind::f16 rate).
__device__ __forceinline__ void mma_f8_ss_pair_if(bool issue, uint32_t d_tmem, uint64_t a_desc,
                                                  uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(static_cast<uint32_t>(issue))
      : "memory");
}
// Block-scaled MXFP4 (kind::mxf4: packed e2m1 operands, one UE8M0 scale per 32 values along K,
// K = 64 per instruction, four times the kind::f16 rate); the scale factors of A and B sit in TMEM
// at tsfa / tsfb (csrc/mx.cuh), the byte pair they occupy there is selected in the descriptor.
__device__ __forceinline__ void mma_mxf4_ss_pair_if(bool issue, uint32_t d_tmem, uint64_t a_desc,
                                                    uint64_t b_desc, uint32_t idesc, uint32_t tsfa,
                                                    uint32_t tsfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %7, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb),
      "r"(static_cast<uint32_t>(issue))
      : "memory");
}
// Shared memory -> TMEM copy of one 32 x 16-byte chunk (scale factors of 128 rows), replicated into
// the four lane quarters, in BOTH CTAs of the pair from their own shared memory at `sdesc`.
__device__ __forceinline__ void tmem_cp_32x128b_pair_if(bool issue, uint32_t taddr, uint64_t sdesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;\n\t"
      "}\n" ::"r"(taddr),
      "l"(sdesc), "r"(static_cast<uint32_t>(issue))
      : "memory");
}
__device__ __forceinline__ void mma_commit_pair_if(bool issue, uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "h"(cta_mask), "r"(static_cast<uint32_t>(issue))
      : "memory");
}
// Arrive (once every previously issued MMA of this thread has retired) on the barrier at this
// offset in every CTA of `cta_mask`.
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;" ::"r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// Shared-memory matrix descriptor for a K-major tile whose rows are 128 B (64 bf16) wide and
// stored with the 128-byte swizzle (what TMA SWIZZLE_128B writes). 8 rows form one 1024-B
// swizzle atom, so the stride between 8-row groups (SBO) is 1024 B; LBO is unused for this
// layout (set to 1 like CuTe does). Layout type 2 = SWIZZLE_128B, version 1 = sm_100.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);  // [0,14)  start address >> 4
  d |= static_cast<uint64_t>(1) << 16;                     // [16,30) leading byte offset >> 4
  d |= static_cast<uint64_t>(1024 >> 4) << 32;             // [32,46) stride byte offset >> 4
  d |= static_cast<uint64_t>(1) << 46;                     // [46,48) descriptor version
  d |= static_cast<uint64_t>(2) << 61;                     // [61,64) SWIZZLE_128B
  return d;
}

// Same for rows of 64 B (64 e4m3 values) stored with the 64-byte swizzle (TMA SWIZZLE_64B): the
// swizzle atom is 8 rows x 64 B = 512 B, so SBO = 512 B.  Layout type 4 = SWIZZLE_64B.
__device__ __forceinline__ uint64_t make_kmajor_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}

// Rows of 32 B (64 packed e2m1 values) stored with the 32-byte swizzle (TMA SWIZZLE_32B): the swizzle
// atom is 8 rows x 32 B = 256 B, so SBO = 256 B.  Layout type 6 = SWIZZLE_32B.
__device__ __forceinline__ uint64_t make_kmajor_sw32_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;
  return d;
}
// Unswizzled 32 rows x 16 B chunk (source of tcgen05.cp 32x128b): rows 16 B apart, 8-row groups
// 128 B apart (SBO), a single 16-byte column (LBO unused).
__device__ __forceinline__ uint64_t make_chunk16_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(128 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// Instruction descriptor of the block-scaled kind::mxf4 MMA: e2m1 x e2m1 (format code 1), UE8M0
// scales, both operands K-major, K = 64; a_sf / b_sf = first byte (0 or 2) of the scale pair in the
// operand's TMEM scale columns.
__host__ __device__ constexpr uint32_t make_idesc_mxf4(int m, int n, int a_sf, int b_sf) {
  return (static_cast<uint32_t>(b_sf) << 4) | (1u << 7) | (1u << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (1u << 23) | (static_cast<uint32_t>(m >> 4) << 24) |
         (static_cast<uint32_t>(a_sf) << 29);
}

// Instruction descriptor with operand format code 0 for A and B: fp16 x fp16 under kind::f16,
// e4m3 x e4m3 under kind::f8f6f4; fp32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_fmt0_f32(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(int m, int n) {
  return (1u << 4)                               // D format  = F32
         | (1u << 7)                             // A format  = BF16
         | (1u << 10)                            // B format  = BF16
         | (static_cast<uint32_t>(n >> 3) << 17) // N >> 3
         | (static_cast<uint32_t>(m >> 4) << 24);// M >> 4
}

}  // namespace ptx
}  // namespace aclip
