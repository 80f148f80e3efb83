// tc05.cuh -- thin inline-PTX layer over the Blackwell (sm_100a) tensor-core path used by libcrfgpu:
// tcgen05.mma with shared-memory descriptors and TMEM accumulators, mbarriers, TMEM alloc/ld, and the
// bf16 hi/lo split that carries fp32 operands through bf16 tensor cores at ~16 mantissa bits
// (x = hi + lo, x*y ~= hi*hi' + lo*hi' + hi*lo', accumulated in fp32 TMEM).
//
// Shared-memory operand tiles use the NO-SWIZZLE canonical layouts of the UMMA descriptor
// (core matrix = 8 rows x 16 bytes = 128 contiguous bytes):
//   K-major  tile [R x KC]: elem (r,k) at (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2     16-byte row = 8 consecutive k
//   MN-major tile [R x KC]: elem (r,k) at (r/8)*SBO + (k/8)*LBO + (k%8)*16 + (r%8)*2     16-byte row = 8 consecutive r
// LBO = byte distance between core matrices adjacent along K, SBO = along M/N.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok)
	    : "r"(smem_u32(bar)), "r"(parity)
	    : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	while (!mbar_try_wait(bar, parity)) {}
}
// acquire at cluster scope: pairs with remote arrivals / async stores of other CTAs of the cluster
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
	uint32_t ok;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok)
	    : "r"(smem_u32(bar)), "r"(parity)
	    : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
	while (!mbar_try_wait_cluster(bar, parity)) {}
}

// generic-proxy writes to shared memory -> visible to the async proxy (tcgen05.mma operand fetch, bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM --------------------------------------------------------------------------------------
// one full warp; ncols power of two in [32,512]; the TMEM base address is written to *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// warp-collective: thread i of warp w reads TMEM lane 32*(w%4)+i, 8 / 16 consecutive columns from `taddr`
// (taddr = base | lane<<16 | column; the lane field must be the warp's own 32-lane quadrant)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
	uint32_t r[8];
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
	             : "r"(taddr)
	             : "memory");
#pragma unroll
	for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
	uint32_t r[4];
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
#pragma unroll
	for (int i = 0; i < 4; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
	uint32_t r[16];
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
	      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
	    : "r"(taddr)
	    : "memory");
#pragma unroll
	for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// one lane of a converged warp; the canonical way to issue tcgen05.mma / commit from warp-uniform code (operands that are
// computed by the whole warp stay in uniform registers; wrapping the issue loop in `if (lane == 0)` instead makes the compiler
// broadcast every operand through an ELECT/R2UR loop, ~85 cycles per MMA)
__device__ __forceinline__ bool elect_one() {
	uint32_t pred;
	asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
	return pred != 0;
}

// ---- descriptors -------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle (layout_type 0), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
	uint64_t d = 0;
	d |= (uint64_t)((saddr >> 4) & 0x3fffu);
	d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
	d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
	d |= (uint64_t)1 << 46;
	return d;
}
// instruction descriptor: bf16 x bf16 -> f32, dense; a_mn / b_mn select MN-major operands
__host__ __device__ constexpr uint32_t idesc_bf16_f32(uint32_t M, uint32_t N, bool a_mn, bool b_mn) {
	return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
	    :
	    : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
	    : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 lanes = rows, 8 columns of packed bf16x2 per 16-wide k-step) lives in TMEM
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
	    :
	    : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
	    : "memory");
}
// warp-collective store: thread i of warp w writes TMEM lane 32*(w%4)+i, 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
	             :
	             : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
	             : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrive on `bar` once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- bf16 hi/lo split --------------------------------------------------------------------------
// packs (hi(a), hi(b)) and (lo(a), lo(b)) as bf16x2 words, element `a` in the low half
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
	const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
	const float2 hf = __bfloat1622float2(h);
	const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
	hi = *reinterpret_cast<const uint32_t*>(&h);
	lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 8 consecutive fp32 -> one 16-byte row of the hi tile and one of the lo tile
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
	split2(x[0], x[1], hi.x, lo.x);
	split2(x[2], x[3], hi.y, lo.y);
	split2(x[4], x[5], hi.z, lo.z);
	split2(x[6], x[7], hi.w, lo.w);
}

}  // namespace tc05
