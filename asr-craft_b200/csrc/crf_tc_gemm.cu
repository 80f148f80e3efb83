// Tensor-core GEMMs of libcrfgpu (sm_100a): tcgen05.mma with fp32 accumulators in TMEM, operands carried as
// bf16 hi/lo pairs (three MMAs per k-step: hi*hi + lo*hi + hi*lo, ~16 mantissa bits, measured worst relative
// error 3e-6 on the probe in tools/tc_probe.cu).
//
//   score_gemm_tc   S[n][j]      = sum_k X[n][k] * W[j][k] + bias[j]                (GEMM-1, both operands K-major)
//                   replaces L calls per frame of CRF_StdFeatureMap::computeStateArrayValue
//                   (CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:65-81)
//   reduce_gemm_tc  out[map(i,j)] += scale * sum_n A[n-shift][i] * B[n][j]          (GEMM-3, both operands MN-major)
//   xi_gemm_tc      the same product for the transition-bias counts, ALL durations of a group in one pass
//                   replace the per-frame scatter of computeStateExpF / computeTransExpF (:130-223) and
//                   `grad -= ExpF` (CRF/src/trainers/gradbuilders/CRF_NewGradBuilder.cpp:374-376)
//
// Structure (one CTA = one 128 x 64 output tile; xi: 128 x up to 5*64):
//   producer warps  fp32 loads from HBM/L2 -> split into bf16 hi/lo in registers -> stores into the no-swizzle canonical UMMA
//                   layout of a shared-memory ring.  FOUR LANES SHARE ONE 32-BYTE VECTOR (8 consecutive floats = one
//                   16-byte row of a bf16 core matrix): every load instruction of a warp reads 8 whole 32-byte sectors and
//                   every store instruction writes 128 contiguous bytes (no L1 re-fetch of partially used sectors, no
//                   bank conflicts).  The next chunk's loads are in flight while the current chunk is converted.
//   MMA warp        one elected lane issues tcgen05.mma, releases ring stages with tcgen05.commit
//   warps 0-3       run the epilogue afterwards (warp w owns TMEM lanes 32w..32w+31)
// Register staging (instead of TMA) lets the loaders apply what these operands need on the way: the row shift of the
// Xi product, the constant-1 bias feature, ragged bounds and the 8-byte alignment of window rows.
#include <algorithm>

#include "crf_kernels.cuh"
#include "tc05.cuh"

namespace crfgpu {

using namespace tc05;

namespace {

constexpr int BM = 128, BN = 64, KC = 32, STAGES = 4;
constexpr uint32_t A_TILE = BM * KC * 2, B_TILE = BN * KC * 2;        // bytes of one bf16 tile
constexpr uint32_t STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;             // hi + lo of both operands = 24576
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024;          // + barriers / alignment slack
constexpr int PW = 8, N_PRODUCERS = PW * 32, N_THR = (PW + 1) * 32;   // 8 producer warps + the MMA warp, 2 CTAs per SM

struct Ring {
	uint64_t full[STAGES], empty[STAGES], done;
	uint32_t tmem;
};

// two consecutive floats of a row; `valid` of them (0..2) exist, `vec2` = the address is 8-byte aligned
__device__ __forceinline__ float2 load2(const float* p, uint32_t valid, bool vec2) {
	if (valid == 2 && vec2) return __ldg(reinterpret_cast<const float2*>(p));
	float2 v = make_float2(0.0f, 0.0f);
	if (valid > 0) v.x = __ldg(p);
	if (valid > 1) v.y = __ldg(p + 1);
	return v;
}
__device__ __forceinline__ uint32_t clamp2(uint32_t pos, uint32_t ext) { return pos >= ext ? 0u : min(2u, ext - pos); }
// store one float2 as the bf16x2 words of the hi and the lo tile
__device__ __forceinline__ void store_split(unsigned char* hi_tile, uint32_t lo_off, uint32_t o, float2 v) {
	uint32_t h, l;
	split2(v.x, v.y, h, l);
	*reinterpret_cast<uint32_t*>(hi_tile + o) = h; *reinterpret_cast<uint32_t*>(hi_tile + lo_off + o) = l;
}

// the MMA warp's walk over the ring: 3 MMAs per 16-wide k-step (hi*hi + lo*hi + hi*lo)
template <bool MN_MAJOR>
__device__ __forceinline__ void mma_ring(unsigned char* smem, Ring* ring, uint32_t tmem, uint32_t n_chunks) {
	constexpr uint32_t idesc = idesc_bf16_f32(BM, BN, MN_MAJOR, MN_MAJOR);
	for (uint32_t c = 0; c < n_chunks; c++) {
		const uint32_t s = c % STAGES;
		mbar_wait(&ring->full[s], (c / STAGES) & 1);
		tc_fence_after();
		const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
		if (elect_one()) {
#pragma unroll
			for (int ks = 0; ks < KC / 16; ks++) {
				const uint64_t ah = smem_desc(base + ks * 256, 128, 512), al = smem_desc(base + A_TILE + ks * 256, 128, 512);
				const uint64_t bh = smem_desc(base + 2 * A_TILE + ks * 256, 128, 512), bl = smem_desc(base + 2 * A_TILE + B_TILE + ks * 256, 128, 512);
				mma_ss(tmem, ah, bh, idesc, (c | ks) != 0);
				mma_ss(tmem, al, bh, idesc, true);
				mma_ss(tmem, ah, bl, idesc, true);
			}
			mma_commit(&ring->empty[s]);
		}
		__syncwarp();
	}
	if (elect_one()) mma_commit(&ring->done);
	__syncwarp();
}

// ------------------------------------------------------------------------------------------------
// GEMM-1.  A = X rows (frames) x K features, B = W rows (labels) x K; both K-major.
// smem: element (row r, k) of a tile at (r/8)*512 + (k/8)*128 + (r%8)*16 + (k%8)*2   (LBO 128, SBO 512)
// producer thread = (row sub = lane/4 of a row block, float2 piece = lane%4 of a k-group); warp w owns row blocks w, w+8 of A
// and row block w of B, all four k-groups of the chunk: 12 float2 per chunk.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(N_THR, 2) score_gemm_tc_kernel(ScoreGemmParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	Ring* ring = reinterpret_cast<Ring*>(smem + STAGES * STAGE_BYTES);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t m0 = (p.n_fast ? blockIdx.y : blockIdx.x) * BM, n0 = (p.n_fast ? blockIdx.x : blockIdx.y) * BN;
	const uint32_t n_chunks = (p.K + KC - 1) / KC;
	if (tid == 0) {
		for (int s = 0; s < STAGES; s++) { mbar_init(&ring->full[s], N_PRODUCERS); mbar_init(&ring->empty[s], 1); }
		mbar_init(&ring->done, 1);
		fence_mbar_init();
	}
	if (warp == PW) tmem_alloc(&ring->tmem, BN);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = ring->tmem;

	if (warp < PW) {
		const bool a_vec2 = ((reinterpret_cast<uintptr_t>(p.A) & 7) == 0) && (p.lda % 2 == 0);
		const bool b_vec2 = ((reinterpret_cast<uintptr_t>(p.B) & 7) == 0) && (p.ldb % 2 == 0);
		const uint32_t sub = lane >> 2, piece = lane & 3;
		const uint32_t gm0 = m0 + warp * 8 + sub, gm1 = gm0 + 64, gn = n0 + warp * 8 + sub;
		const bool ok0 = gm0 < p.M, ok1 = gm1 < p.M, okb = gn < p.Ncols;
		const float* a0p = p.A + (uint64_t)(ok0 ? gm0 : 0) * p.lda + piece * 2;
		const float* a1p = p.A + (uint64_t)(ok1 ? gm1 : 0) * p.lda + piece * 2;
		const float* bp = p.B + (uint64_t)(okb ? gn : 0) * p.ldb + piece * 2;
		float2 x[12];
		auto fetch = [&](uint32_t c) {
#pragma unroll
			for (uint32_t g = 0; g < 4; g++) {
				const uint32_t k = c * KC + g * 8, v = clamp2(k + piece * 2, p.K);
				x[g] = load2(a0p + k, ok0 ? v : 0u, a_vec2);
				x[4 + g] = load2(a1p + k, ok1 ? v : 0u, a_vec2);
				x[8 + g] = load2(bp + k, okb ? v : 0u, b_vec2);
			}
		};
		fetch(0);
		const uint32_t o = warp * 512 + sub * 16 + piece * 4;
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t s = c % STAGES;
			uint32_t h[12], l[12];
#pragma unroll
			for (int i = 0; i < 12; i++) split2(x[i].x, x[i].y, h[i], l[i]);
			if (c + 1 < n_chunks) fetch(c + 1);                  // in flight while this chunk is stored and the ring slot awaited
			if (c >= STAGES) mbar_wait(&ring->empty[s], ((c / STAGES) - 1) & 1);
			unsigned char* st = smem + s * STAGE_BYTES;
#pragma unroll
			for (uint32_t g = 0; g < 4; g++) {
				*reinterpret_cast<uint32_t*>(st + o + g * 128) = h[g]; *reinterpret_cast<uint32_t*>(st + A_TILE + o + g * 128) = l[g];
				*reinterpret_cast<uint32_t*>(st + o + 8 * 512 + g * 128) = h[4 + g]; *reinterpret_cast<uint32_t*>(st + A_TILE + o + 8 * 512 + g * 128) = l[4 + g];
				*reinterpret_cast<uint32_t*>(st + 2 * A_TILE + o + g * 128) = h[8 + g]; *reinterpret_cast<uint32_t*>(st + 2 * A_TILE + B_TILE + o + g * 128) = l[8 + g];
			}
			fence_proxy_async_smem();
			mbar_arrive(&ring->full[s]);
		}
	} else {
		mma_ring<false>(smem, ring, tmem, n_chunks);
	}
	// ---- epilogue: TMEM -> registers -> shared transpose -> (+bias) -> coalesced rows of S ----
	if (warp < 4) {
		mbar_wait(&ring->done, 0);
		tc_fence_after();
		float* Cs = reinterpret_cast<float*>(smem);       // [128][65], the ring is idle now
		const uint32_t row = warp * 32 + lane;
#pragma unroll
		for (int c0 = 0; c0 < BN; c0 += 16) {
			float v[16];
			tmem_ld16(tmem + ((warp * 32u) << 16) + c0, v);
			tmem_ld_wait();
#pragma unroll
			for (int j = 0; j < 16; j++) Cs[row * 65 + c0 + j] = v[j];
		}
		tc_fence_before();
		asm volatile("bar.sync 1, 128;" ::: "memory");
		const uint32_t ncol = min((uint32_t)BN, p.Ncols - n0);
		for (uint32_t i = tid; i < BM * BN; i += 128) {
			const uint32_t r = i / BN, j = i % BN, gm = m0 + r;
			if (gm < p.M && j < ncol) p.C[(uint64_t)gm * p.ldc + n0 + j] = Cs[r * 65 + j] + (p.bias ? __ldg(p.bias + n0 + j) : 0.0f);
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == PW) tmem_dealloc(tmem, BN);
}

// ------------------------------------------------------------------------------------------------
// GEMM-3.  Both operands are [frames][columns] row-major, i.e. MN-major with the reduction index slow.
// SWAP = false: MMA M side = p.A columns (I), N side = p.B columns (J)
// SWAP = true : MMA M side = p.B columns (J), N side = p.A columns (I)     (state weights: J = features, I = P)
// smem vector (frame k, column group rg) at rg*512 + k*16   (LBO 128 = 8 frames, SBO 512)
// 4 producer warps, one lane per 32-byte vector (for this operand shape the 4-lanes-per-vector mapping of the other two
// kernels measured slower: 5.2 vs 3.2 ms per step)
// ------------------------------------------------------------------------------------------------
constexpr int V_PRODUCERS = 128;
__device__ __forceinline__ void load8(const float* p, bool ok, bool vec2, float (&x)[8]) {
	if (!ok) {
#pragma unroll
		for (int j = 0; j < 8; j++) x[j] = 0.0f;
	} else if (vec2) {
#pragma unroll
		for (int j = 0; j < 4; j++) { const float2 v = __ldg(reinterpret_cast<const float2*>(p) + j); x[2 * j] = v.x; x[2 * j + 1] = v.y; }
	} else {
#pragma unroll
		for (int j = 0; j < 8; j++) x[j] = __ldg(p + j);
	}
}

template <bool SWAP>
__global__ void __launch_bounds__(160, 2) reduce_gemm_tc_kernel(ReduceGemmParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	Ring* ring = reinterpret_cast<Ring*>(smem + STAGES * STAGE_BYTES);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	// M-side / N-side views of the two operands
	const float* Mp = SWAP ? p.B : p.A; const uint64_t ldm = SWAP ? p.ldb : p.lda; const uint32_t Mext = SWAP ? p.J : p.I;
	const float* Np = SWAP ? p.A : p.B; const uint64_t ldn = SWAP ? p.lda : p.ldb; const uint32_t Next = SWAP ? p.I : p.J;
	const uint32_t m_shift = SWAP ? 0 : p.a_row_shift, n_shift = SWAP ? p.a_row_shift : 0;
	const uint32_t m_ones = SWAP ? p.ones_col : 0xffffffffu, n_ones = SWAP ? 0xffffffffu : p.ones_col;
	const uint32_t m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
	const uint32_t ns = p.n0 + blockIdx.z * p.k_slab, ne = min(ns + p.k_slab, p.n1);
	const uint32_t n_chunks = (ne - ns + KC - 1) / KC;
	if (tid == 0) {
		for (int s = 0; s < STAGES; s++) { mbar_init(&ring->full[s], V_PRODUCERS); mbar_init(&ring->empty[s], 1); }
		mbar_init(&ring->done, 1);
		fence_mbar_init();
	}
	if (warp == 4) tmem_alloc(&ring->tmem, BN);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = ring->tmem;

	if (warp < 4) {
		const bool m_vec2 = ((reinterpret_cast<uintptr_t>(Mp) & 7) == 0) && (ldm % 2 == 0);
		const bool n_vec2 = ((reinterpret_cast<uintptr_t>(Np) & 7) == 0) && (ldn % 2 == 0);
		const uint32_t k8 = lane & 7, rgq = lane >> 3;
		const uint32_t kk = warp * 8 + k8;                 // frame of this thread inside the chunk
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t s = c % STAGES, n = ns + c * KC + kk;
			const bool n_ok = n < ne;
			float xm[4][8], xn[2][8];
#pragma unroll
			for (int it = 0; it < 4; it++) {
				const uint32_t col = m0 + (it * 4 + rgq) * 8;
				const float* src = Mp + (uint64_t)(n - m_shift) * ldm + col;
				if (n_ok && col + 8 <= Mext && m_ones >= col + 8) load8(src, true, m_vec2, xm[it]);
				else {
#pragma unroll
					for (int j = 0; j < 8; j++) xm[it][j] = (!n_ok || col + j >= Mext) ? 0.0f : (col + j == m_ones ? 1.0f : __ldg(src + j));
				}
			}
#pragma unroll
			for (int it = 0; it < 2; it++) {
				const uint32_t col = n0 + (it * 4 + rgq) * 8;
				const float* src = Np + (uint64_t)(n - n_shift) * ldn + col;
				if (n_ok && col + 8 <= Next && n_ones >= col + 8) load8(src, true, n_vec2, xn[it]);
				else {
#pragma unroll
					for (int j = 0; j < 8; j++) xn[it][j] = (!n_ok || col + j >= Next) ? 0.0f : (col + j == n_ones ? 1.0f : __ldg(src + j));
				}
			}
			if (c >= STAGES) mbar_wait(&ring->empty[s], ((c / STAGES) - 1) & 1);
			unsigned char* st = smem + s * STAGE_BYTES;
#pragma unroll
			for (int it = 0; it < 4; it++) {
				uint4 h, l; split8(xm[it], h, l);
				const uint32_t o = (it * 4 + rgq) * 512 + kk * 16;
				*reinterpret_cast<uint4*>(st + o) = h; *reinterpret_cast<uint4*>(st + A_TILE + o) = l;
			}
#pragma unroll
			for (int it = 0; it < 2; it++) {
				uint4 h, l; split8(xn[it], h, l);
				const uint32_t o = (it * 4 + rgq) * 512 + kk * 16;
				*reinterpret_cast<uint4*>(st + 2 * A_TILE + o) = h; *reinterpret_cast<uint4*>(st + 2 * A_TILE + B_TILE + o) = l;
			}
			fence_proxy_async_smem();
			mbar_arrive(&ring->full[s]);
		}
	} else {
		// warp 4: the whole warp walks the ring (uniform descriptors), one elected lane issues
		constexpr uint32_t idesc = idesc_bf16_f32(BM, BN, true, true);
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t s = c % STAGES;
			mbar_wait(&ring->full[s], (c / STAGES) & 1);
			tc_fence_after();
			const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
			if (elect_one()) {
#pragma unroll
				for (int ks = 0; ks < KC / 16; ks++) {
					const uint64_t ah = smem_desc(base + ks * 256, 128, 512), al = smem_desc(base + A_TILE + ks * 256, 128, 512);
					const uint64_t bh = smem_desc(base + 2 * A_TILE + ks * 256, 128, 512), bl = smem_desc(base + 2 * A_TILE + B_TILE + ks * 256, 128, 512);
					mma_ss(tmem, ah, bh, idesc, (c | ks) != 0);
					mma_ss(tmem, al, bh, idesc, true);
					mma_ss(tmem, ah, bl, idesc, true);
				}
				mma_commit(&ring->empty[s]);
			}
			__syncwarp();
		}
		if (elect_one()) mma_commit(&ring->done);
		__syncwarp();
	}
	// ---- epilogue: lane = M-side column, 64 N-side columns; fp64 atomics into the gradient ----
	if (warp < 4 && n_chunks) {
		mbar_wait(&ring->done, 0);
		tc_fence_after();
		const uint32_t gm = m0 + warp * 32 + lane;
#pragma unroll
		for (int c0 = 0; c0 < BN; c0 += 16) {
			float v[16];
			tmem_ld16(tmem + ((warp * 32u) << 16) + c0, v);
			tmem_ld_wait();
			if (gm < Mext) {
#pragma unroll
				for (int j = 0; j < 16; j++) {
					const uint32_t gn = n0 + c0 + j;
					if (gn >= Next || v[j] == 0.0f) continue;
					const uint32_t gi = SWAP ? gn : gm, gj = SWAP ? gm : gn;
					if (p.mode == 0) {
						const double sc = (gj == p.ones_col) ? p.ones_scale : p.scale;
						const uint32_t ri = __ldg(p.row_idx + gi);
						if (ri != 0xffffffffu) atomicAdd(&p.out[(uint64_t)ri + gj], sc * (double)v[j]);
					} else {
						const uint32_t idx = __ldg(p.pair_idx + (uint64_t)gi * p.pair_ld + gj);
						if (idx != 0xffffffffu) atomicAdd(&p.out[idx], p.scale * (double)__ldg(p.Ew + (uint64_t)gi * p.e_ld + gj) * (double)v[j]);
					}
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 4) tmem_dealloc(tmem, BN);
}

// ------------------------------------------------------------------------------------------------
// Xi, all durations of a group in ONE pass over the forward vectors:
//   Xi_d[q][y] = sum_n A[n-d][q] * R[n][(d,y)]        (CRF_StdFeatureMap::computeTransExpF summed over frames, :197-223)
// One CTA = 128 source labels q x up to XI_G column tiles of R (each with its own duration d = row shift of A) x a frame slab.
// The A rows of a chunk are staged once, with a halo of `halo` earlier frames, in the MN-major layout whose k rows are a
// uniform 16 bytes apart, so the operand of duration d is the same tile viewed (halo - d) rows further down:
// only the descriptor's start address changes.  XI_G accumulator tiles live side by side in TMEM.
// 16 producer warps: the A window has 8 frame blocks x 16 column groups, the R tiles 4 frame blocks x 8 column groups each.
// ------------------------------------------------------------------------------------------------
constexpr int XI_G = 5, XI_ROWS = 64, XI_STAGES = 3, XI_PW = 16, XI_THR = (XI_PW + 1) * 32;
constexpr uint32_t XI_A_TILE = 16 * XI_ROWS * 16;                       // [16 column groups][64 frames][16 B] per hi / lo
constexpr uint32_t XI_R_TILE = 8 * KC * 16;                             // [8 column groups][32 frames][16 B] per hi / lo
constexpr uint32_t XI_STAGE = 2 * XI_A_TILE + XI_G * 2 * XI_R_TILE;     // 32 KB + 40 KB
constexpr uint32_t XI_SMEM = XI_STAGES * XI_STAGE + 1024;

struct XiRing {
	uint64_t full[XI_STAGES], empty[XI_STAGES], done;
	uint32_t tmem;
};

__global__ void __launch_bounds__(XI_THR, 1) xi_gemm_tc_kernel(XiGemmParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	XiRing* ring = reinterpret_cast<XiRing*>(smem + XI_STAGES * XI_STAGE);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t ytiles = (p.P + BN - 1) / BN, n_tiles = p.D * ytiles;
	const uint32_t m0 = blockIdx.x * BM, t0 = blockIdx.y * XI_G, nt = min((uint32_t)XI_G, n_tiles - t0);
	const uint32_t ns = p.n0 + blockIdx.z * p.k_slab, ne = min(ns + p.k_slab, p.n1);
	const uint32_t n_chunks = (ne - ns + KC - 1) / KC;
	// column tile j of this CTA: duration, first column in R, width
	auto tile_d = [&](uint32_t j) { return (t0 + j) / ytiles + 1; };
	auto tile_c0 = [&](uint32_t j) { return ((t0 + j) / ytiles) * p.P + ((t0 + j) % ytiles) * BN; };
	auto tile_w = [&](uint32_t j) { return min((uint32_t)BN, p.P - ((t0 + j) % ytiles) * BN); };
	const uint32_t halo = (tile_d(nt - 1) + 7) / 8 * 8;                   // largest shift of the group, rounded to a row group
	if (tid == 0) {
		for (int s = 0; s < XI_STAGES; s++) { mbar_init(&ring->full[s], XI_PW * 32); mbar_init(&ring->empty[s], 1); }
		mbar_init(&ring->done, 1);
		fence_mbar_init();
	}
	if (warp == XI_PW) tmem_alloc(&ring->tmem, 512);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = ring->tmem;

	if (warp < XI_PW) {
		const bool a_vec2 = ((reinterpret_cast<uintptr_t>(p.A) & 7) == 0) && (p.lda % 2 == 0);
		const bool r_al = ((reinterpret_cast<uintptr_t>(p.R) & 7) == 0) && (p.ldb % 2 == 0);
		const uint32_t sub = lane >> 2, piece = lane & 3;
		const uint32_t a_rows = halo + KC;
		// A: warp w owns frame block w%8 of the staged window and the column-group half w/8 (8 float2)
		const uint32_t ar = (warp & 7) * 8 + sub, acg0 = (warp >> 3) * 8;
		// R: 32 (frame block, column group) cells per column tile, nt*32 in all; warp w owns cells w, w+16, ... (up to 10 float2).
		// Everything that does not depend on the chunk is worked out once: column offset and how many of my 2 floats exist.
		uint32_t roff[10], rinfo[10];                          // rinfo: bits 0-1 valid floats, bit 2 8-byte aligned, bit 3 cell exists
#pragma unroll
		for (uint32_t i = 0; i < 10; i++) {
			const uint32_t id = warp + 16 * i, j = id >> 5, cg = id & 7;
			roff[i] = 0; rinfo[i] = 0;
			if (j < nt) {
				const uint32_t c0 = tile_c0(j), cc = cg * 8 + piece * 2;
				roff[i] = c0 + cc;
				rinfo[i] = clamp2(cc, tile_w(j)) | ((r_al && (c0 % 2 == 0)) ? 4u : 0u) | 8u;
			}
		}
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t s = c % XI_STAGES, nc = ns + c * KC;
			float2 xa[8], xr[10];
			{
				const int64_t n = (int64_t)nc - halo + ar;
				const bool ok = ar < a_rows && n >= 0 && n < (int64_t)p.n_frames;
				const float* arow = p.A + (uint64_t)(ok ? n : 0) * p.lda;
#pragma unroll
				for (uint32_t g = 0; g < 8; g++) {
					const uint32_t col = m0 + (acg0 + g) * 8 + piece * 2;
					xa[g] = load2(arow + col, ok ? clamp2(col, p.L) : 0u, a_vec2);
				}
			}
#pragma unroll
			for (uint32_t i = 0; i < 10; i++) {
				const uint32_t fb = ((warp + 16 * i) & 31) >> 3;
				xr[i] = make_float2(0.0f, 0.0f);
				if (rinfo[i] & 8u) {
					const uint32_t n = nc + fb * 8 + sub;
					const bool ok = n < ne;
					xr[i] = load2(p.R + (uint64_t)(ok ? n : 0) * p.ldb + roff[i], ok ? (rinfo[i] & 3u) : 0u, (rinfo[i] & 4u) != 0);
				}
			}
			if (c >= XI_STAGES) mbar_wait(&ring->empty[s], ((c / XI_STAGES) - 1) & 1);
			unsigned char* st = smem + s * XI_STAGE;
			if (ar < a_rows) {
#pragma unroll
				for (uint32_t g = 0; g < 8; g++) store_split(st, XI_A_TILE, (acg0 + g) * (XI_ROWS * 16) + ar * 16 + piece * 4, xa[g]);
			}
#pragma unroll
			for (uint32_t i = 0; i < 10; i++) {
				const uint32_t id = warp + 16 * i, j = id >> 5, fb = (id & 31) >> 3, cg = id & 7;
				if (j < nt) store_split(st + 2 * XI_A_TILE + j * 2 * XI_R_TILE, XI_R_TILE, cg * 512 + (fb * 8 + sub) * 16 + piece * 4, xr[i]);
			}
			fence_proxy_async_smem();
			mbar_arrive(&ring->full[s]);
		}
	} else {
		constexpr uint32_t idesc = idesc_bf16_f32(BM, BN, true, true);
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t s = c % XI_STAGES;
			mbar_wait(&ring->full[s], (c / XI_STAGES) & 1);
			tc_fence_after();
			if (elect_one()) {
				const uint32_t base = smem_u32(smem + s * XI_STAGE);
				for (uint32_t j = 0; j < nt; j++) {
					const uint32_t a_off = (halo - tile_d(j)) * 16;              // the duration's row shift of A
					const uint32_t rb = base + 2 * XI_A_TILE + j * 2 * XI_R_TILE;
#pragma unroll
					for (int ks = 0; ks < KC / 16; ks++) {
						const uint64_t ah = smem_desc(base + a_off + ks * 256, 128, XI_ROWS * 16), al = smem_desc(base + XI_A_TILE + a_off + ks * 256, 128, XI_ROWS * 16);
						const uint64_t bh = smem_desc(rb + ks * 256, 128, 512), bl = smem_desc(rb + XI_R_TILE + ks * 256, 128, 512);
						mma_ss(tmem + j * BN, ah, bh, idesc, (c | ks) != 0);
						mma_ss(tmem + j * BN, al, bh, idesc, true);
						mma_ss(tmem + j * BN, ah, bl, idesc, true);
					}
				}
				mma_commit(&ring->empty[s]);
			}
			__syncwarp();
		}
		if (elect_one()) mma_commit(&ring->done);
		__syncwarp();
	}
	// ---- epilogue: lane = source label q; per column tile 64 destination labels; fp64 atomics into the gradient ----
	if (warp < 4 && n_chunks) {
		mbar_wait(&ring->done, 0);
		tc_fence_after();
		const uint32_t q = m0 + warp * 32 + lane;
		for (uint32_t j = 0; j < nt; j++) {
			const uint32_t c0 = tile_c0(j), w = tile_w(j);
#pragma unroll
			for (int cc = 0; cc < BN; cc += 16) {
				float v[16];
				tmem_ld16(tmem + ((warp * 32u) << 16) + j * BN + cc, v);
				tmem_ld_wait();
				if (q < p.L) {
#pragma unroll
					for (int jj = 0; jj < 16; jj++) {
						const uint32_t col = cc + jj;
						if (col >= w || v[jj] == 0.0f) continue;
						const uint32_t idx = __ldg(p.pair_idx + (uint64_t)q * p.L + c0 + col);
						if (idx != 0xffffffffu) atomicAdd(&p.out[idx], p.scale * (double)__ldg(p.Ew + (uint64_t)q * p.e_ld + c0 + col) * (double)v[jj]);
					}
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == XI_PW) tmem_dealloc(tmem, 512);
}

}  // namespace

cudaError_t launch_score_gemm_tc(const ScoreGemmParams& p, cudaStream_t s) {
	if (!p.M || !p.Ncols) return cudaSuccess;
	bool attr_done = false;      // (function attributes are per device: no process-wide cache, a multi-GPU process configures each one)
	if (!attr_done) {
		cudaError_t e = cudaFuncSetAttribute(score_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
		if (e != cudaSuccess) return e;
		attr_done = true;
	}
	// Many column tiles (the transition-score GEMM: labels^2 columns): the column tiles of one row tile run back to back, so the row
	// tile of X -- strided rows, far more bytes than the weights -- is read from DRAM once and from L2 by the other column tiles
	// (recipe shape, 36 column tiles: 10.9 GB of DRAM reads per launch with the row tiles fastest, profiles/r2v_recipe_kernels_full.md)
	const uint32_t mt = (p.M + BM - 1) / BM, nt = (p.Ncols + BN - 1) / BN;
	ScoreGemmParams q = p;
	q.n_fast = (nt > 1 && mt <= 65535) ? 1u : 0u;
	dim3 grid(q.n_fast ? nt : mt, q.n_fast ? mt : nt);
	score_gemm_tc_kernel<<<grid, N_THR, SMEM_BYTES, s>>>(q);
	return cudaGetLastError();
}

cudaError_t launch_xi_gemm_tc(const XiGemmParams& p, cudaStream_t s) {
	if (p.n1 <= p.n0 || !p.L || !p.P || p.D > 32) return p.D > 32 ? cudaErrorInvalidValue : cudaSuccess;
	bool attr_done = false;      // (function attributes are per device: no process-wide cache, a multi-GPU process configures each one)
	if (!attr_done) {
		cudaError_t e = cudaFuncSetAttribute(xi_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XI_SMEM);
		if (e != cudaSuccess) return e;
		attr_done = true;
	}
	const uint32_t n_tiles = p.D * ((p.P + BN - 1) / BN);
	dim3 grid((p.L + BM - 1) / BM, (n_tiles + XI_G - 1) / XI_G, (p.n1 - p.n0 + p.k_slab - 1) / p.k_slab);
	xi_gemm_tc_kernel<<<grid, XI_THR, XI_SMEM, s>>>(p);
	return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// The same product from operands that are ALREADY split and tiled (labels^2 rows: the transition-feature gradient, 342 Gflop at the
// TIMIT recipe's shape).  The register-staged kernel above spends its time in its four producer warps (strided fp32 loads, split,
// shared-memory stores: tensor pipe 9 %, profiles/r2v_recipe_kernels_full.md); here a tiling pass writes both operands once as bf16
// hi / lo halves in the byte order of the shared-memory tiles -- [column group][frame of the chunk][8 columns] per 32-frame chunk
// and column tile, hi then lo -- so a ring stage is two bulk copies (cp.async.bulk, 16 KB + 8 KB) issued by one thread, and every
// operand byte crosses L2 -> shared memory exactly as the MMA reads it.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// src[n][col0 + c], c < ncols (column ones_col reads as the constant 1), n < N  ->  tiles of T columns x 32 frames:
// dst + ((chunk * n_ct + ct) * 2 + {hi, lo}) * T * 64 bytes, element (column group cg, frame kk, j) at cg * 512 + kk * 16 + j * 2.
// lane = frame of the chunk (a thread reads whole 32-byte sectors of its row, a warp writes 512 contiguous bytes), warp = column groups.
template <int T>
__global__ void __launch_bounds__(256) tile_mn_kernel(const float* __restrict__ src, uint64_t ld, uint32_t ncols, uint32_t ones_col, uint32_t N, unsigned char* __restrict__ dst, uint32_t n_ct) {
	const uint32_t chunk = blockIdx.x, ct = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t n = chunk * KC + lane;
	const bool vec2 = ((reinterpret_cast<uintptr_t>(src) & 7) == 0) && (ld % 2 == 0);
	unsigned char* hi = dst + ((size_t)chunk * n_ct + ct) * 2 * (T * 64);
	for (uint32_t cg = warp; cg < T / 8; cg += 8) {
		const uint32_t col = ct * T + cg * 8;
		float x[8];
		const float* r = src + (uint64_t)n * ld + col;
		if (n < N && col + 8 <= ncols && ones_col >= col + 8) load8(r, true, vec2, x);
		else {
#pragma unroll
			for (int j = 0; j < 8; j++) x[j] = (n >= N || col + j >= ncols) ? 0.0f : (col + j == ones_col ? 1.0f : __ldg(r + j));
		}
		uint4 h, l; split8(x, h, l);
		const uint32_t o = cg * 512 + lane * 16;
		*reinterpret_cast<uint4*>(hi + o) = h; *reinterpret_cast<uint4*>(hi + T * 64 + o) = l;
	}
}

// tiles of the bulk-copy-fed kernels: 128 x 128 outputs (both operands as 128-wide tiles: half the L2 -> shared-memory traffic per MMA of
// the 128 x 64 tiles of the register-staged kernels), three ring stages of 32 KB, two CTAs per SM
constexpr int TBN = 128, TSTAGES = 3;
constexpr uint32_t TB_TILE = TBN * KC * 2, TSTAGE = 2 * A_TILE + 2 * TB_TILE, TSMEM = TSTAGES * TSTAGE + 1024;

__global__ void __launch_bounds__(192, 2) reduce_gemm_tiled_kernel(TiledReduceParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	Ring* ring = reinterpret_cast<Ring*>(smem + TSTAGES * TSTAGE);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t mt = blockIdx.x, nt = blockIdx.y, m0 = mt * BM, n0 = nt * TBN;
	const uint32_t c_lo = blockIdx.z * p.slab_chunks, c_hi = min(c_lo + p.slab_chunks, p.n_chunks);
	const uint32_t n_chunks = c_hi - c_lo;
	if (tid == 0) {
		for (int s = 0; s < TSTAGES; s++) { mbar_init(&ring->full[s], 1); mbar_init(&ring->empty[s], 1); }
		mbar_init(&ring->done, 1);
		fence_mbar_init();
	}
	if (warp == 4) tmem_alloc(&ring->tmem, TBN);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = ring->tmem;
	if (warp == 5) {
		if (lane == 0) {
			for (uint32_t c = 0; c < n_chunks; c++) {
				const uint32_t s = c % TSTAGES;
				if (c >= TSTAGES) mbar_wait(&ring->empty[s], ((c / TSTAGES) - 1) & 1);
				unsigned char* st = smem + s * TSTAGE;
				mbar_arrive_expect_tx(&ring->full[s], TSTAGE);
				bulk_g2s(st, p.At + ((size_t)(c_lo + c) * p.n_mt + mt) * (2 * A_TILE), 2 * A_TILE, &ring->full[s]);
				bulk_g2s(st + 2 * A_TILE, p.Bt + ((size_t)(c_lo + c) * p.n_nt + nt) * (2 * TB_TILE), 2 * TB_TILE, &ring->full[s]);
			}
		}
		__syncwarp();
	} else if (warp == 4) {
		constexpr uint32_t idesc = idesc_bf16_f32(BM, TBN, true, true);
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t s = c % TSTAGES;
			mbar_wait(&ring->full[s], (c / TSTAGES) & 1);
			tc_fence_after();
			const uint32_t base = smem_u32(smem + s * TSTAGE);
			if (elect_one()) {
#pragma unroll
				for (int ks = 0; ks < KC / 16; ks++) {
					const uint64_t ah = smem_desc(base + ks * 256, 128, 512), al = smem_desc(base + A_TILE + ks * 256, 128, 512);
					const uint64_t bh = smem_desc(base + 2 * A_TILE + ks * 256, 128, 512), bl = smem_desc(base + 2 * A_TILE + TB_TILE + ks * 256, 128, 512);
					mma_ss(tmem, ah, bh, idesc, (c | ks) != 0);
					mma_ss(tmem, al, bh, idesc, true);
					mma_ss(tmem, ah, bl, idesc, true);
				}
				mma_commit(&ring->empty[s]);
			}
			__syncwarp();
		}
		if (elect_one()) mma_commit(&ring->done);
		__syncwarp();
	}
	// ---- epilogue: lane = M-side column (output row), 64 N-side columns; fp64 atomics into the gradient ----
	if (warp < 4 && n_chunks) {
		mbar_wait(&ring->done, 0);
		tc_fence_after();
		const uint32_t gm = m0 + warp * 32 + lane;
		const uint32_t ri = gm < p.I ? __ldg(p.row_idx + gm) : 0xffffffffu;
#pragma unroll
		for (int c0 = 0; c0 < TBN; c0 += 16) {
			float v[16];
			tmem_ld16(tmem + ((warp * 32u) << 16) + c0, v);
			tmem_ld_wait();
			if (ri != 0xffffffffu) {
#pragma unroll
				for (int j = 0; j < 16; j++) {
					const uint32_t gn = n0 + c0 + j;
					if (gn >= p.J || v[j] == 0.0f) continue;
					const double sc = (gn == p.ones_col) ? p.ones_scale : p.scale;
					atomicAdd(&p.out[(uint64_t)ri + gn], sc * (double)v[j]);
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 4) tmem_dealloc(tmem, TBN);
}

// K-major twin for the transition SCORES  M[n][c] = sum_k X[n][k] * W[c][k] + bias[c]  (labels^2 columns): rows x 32-feature chunks of
// both operands as tiles in the byte order of score_gemm_tc_kernel's shared-memory tiles -- element (row r, k) at
// (r / 8) * 512 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2, hi then lo -- at dst + ((row tile * n_kc + chunk) * 2 + {hi, lo}) * T * 64.
// lane = row of a 32-row block (whole 32-byte sectors per thread, 128-byte runs per 8 lanes), warp = (row block, k group).
template <int T>
__global__ void __launch_bounds__(256) tile_k_kernel(const float* __restrict__ src, uint64_t ld, uint32_t rows, uint32_t K, unsigned char* __restrict__ dst, uint32_t n_kc) {
	const uint32_t rt = blockIdx.x, kc = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const bool vec2 = ((reinterpret_cast<uintptr_t>(src) & 7) == 0) && (ld % 2 == 0);
	unsigned char* hi = dst + ((size_t)rt * n_kc + kc) * 2 * (T * 64);
	for (uint32_t w = warp; w < (T / 32) * 4; w += 8) {
		const uint32_t rb = w >> 2, kg = w & 3, r = rb * 32 + lane, row = rt * T + r, k = kc * KC + kg * 8;
		float x[8];
		const float* p = src + (uint64_t)row * ld + k;
		if (row < rows && k + 8 <= K) load8(p, true, vec2, x);
		else {
#pragma unroll
			for (int j = 0; j < 8; j++) x[j] = (row < rows && k + j < K) ? __ldg(p + j) : 0.0f;
		}
		uint4 h, l; split8(x, h, l);
		const uint32_t o = (r >> 3) * 512 + kg * 128 + (r & 7) * 16;
		*reinterpret_cast<uint4*>(hi + o) = h; *reinterpret_cast<uint4*>(hi + T * 64 + o) = l;
	}
}

__global__ void __launch_bounds__(192, 2) score_gemm_tiled_kernel(TiledScoreParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	Ring* ring = reinterpret_cast<Ring*>(smem + TSTAGES * TSTAGE);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t nt = blockIdx.x, mt = blockIdx.y, m0 = mt * BM, n0 = nt * TBN;      // column tiles fastest: the row tile of X stays in L2
	const uint32_t n_chunks = p.n_kc;
	if (tid == 0) {
		for (int s = 0; s < TSTAGES; s++) { mbar_init(&ring->full[s], 1); mbar_init(&ring->empty[s], 1); }
		mbar_init(&ring->done, 1);
		fence_mbar_init();
	}
	if (warp == 4) tmem_alloc(&ring->tmem, TBN);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = ring->tmem;
	if (warp == 5) {
		if (lane == 0) {
			for (uint32_t c = 0; c < n_chunks; c++) {
				const uint32_t s = c % TSTAGES;
				if (c >= TSTAGES) mbar_wait(&ring->empty[s], ((c / TSTAGES) - 1) & 1);
				unsigned char* st = smem + s * TSTAGE;
				mbar_arrive_expect_tx(&ring->full[s], TSTAGE);
				bulk_g2s(st, p.At + ((size_t)mt * p.n_kc + c) * (2 * A_TILE), 2 * A_TILE, &ring->full[s]);
				bulk_g2s(st + 2 * A_TILE, p.Bt + ((size_t)nt * p.n_kc + c) * (2 * TB_TILE), 2 * TB_TILE, &ring->full[s]);
			}
		}
		__syncwarp();
	} else if (warp == 4) {
		constexpr uint32_t idesc = idesc_bf16_f32(BM, TBN, false, false);
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t s = c % TSTAGES;
			mbar_wait(&ring->full[s], (c / TSTAGES) & 1);
			tc_fence_after();
			const uint32_t base = smem_u32(smem + s * TSTAGE);
			if (elect_one()) {
#pragma unroll
				for (int ks = 0; ks < KC / 16; ks++) {
					const uint64_t ah = smem_desc(base + ks * 256, 128, 512), al = smem_desc(base + A_TILE + ks * 256, 128, 512);
					const uint64_t bh = smem_desc(base + 2 * A_TILE + ks * 256, 128, 512), bl = smem_desc(base + 2 * A_TILE + TB_TILE + ks * 256, 128, 512);
					mma_ss(tmem, ah, bh, idesc, (c | ks) != 0);
					mma_ss(tmem, al, bh, idesc, true);
					mma_ss(tmem, ah, bl, idesc, true);
				}
				mma_commit(&ring->empty[s]);
			}
			__syncwarp();
		}
		if (elect_one()) mma_commit(&ring->done);
		__syncwarp();
	}
	// ---- epilogue: TMEM -> registers -> shared transpose -> (+bias) -> coalesced rows of the score matrix ----
	if (warp < 4) {
		mbar_wait(&ring->done, 0);
		tc_fence_after();
		float* Cs = reinterpret_cast<float*>(smem);       // [128][TBN + 1], the ring is idle now
		const uint32_t row = warp * 32 + lane;
#pragma unroll
		for (int c0 = 0; c0 < TBN; c0 += 16) {
			float v[16];
			tmem_ld16(tmem + ((warp * 32u) << 16) + c0, v);
			tmem_ld_wait();
#pragma unroll
			for (int j = 0; j < 16; j++) Cs[row * (TBN + 1) + c0 + j] = v[j];
		}
		tc_fence_before();
		asm volatile("bar.sync 1, 128;" ::: "memory");
		const uint32_t ncol = min((uint32_t)TBN, p.Ncols - n0);
		for (uint32_t i = tid; i < BM * TBN; i += 128) {
			const uint32_t r = i / TBN, j = i % TBN, gm = m0 + r;
			if (gm < p.M && j < ncol) p.C[(uint64_t)gm * p.ldc + n0 + j] = Cs[r * (TBN + 1) + j] + (p.bias ? __ldg(p.bias + n0 + j) : 0.0f);
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 4) tmem_dealloc(tmem, TBN);
}

size_t tiled_k_operand_bytes(uint32_t rows, uint32_t K, uint32_t T) { return (size_t)((rows + T - 1) / T) * ((K + KC - 1) / KC) * 2 * ((size_t)T * 64); }

cudaError_t launch_tile_k(const float* src, uint64_t ld, uint32_t rows, uint32_t K, bool m_side, unsigned char* dst, cudaStream_t s) {
	if (!rows || !K) return cudaSuccess;
	const uint32_t T = m_side ? BM : BN, n_kc = (K + KC - 1) / KC;
	dim3 grid((rows + T - 1) / T, n_kc);
	if (m_side) tile_k_kernel<BM><<<grid, 256, 0, s>>>(src, ld, rows, K, dst, n_kc);
	else tile_k_kernel<BN><<<grid, 256, 0, s>>>(src, ld, rows, K, dst, n_kc);
	return cudaGetLastError();
}

cudaError_t launch_score_gemm_tiled(const TiledScoreParams& p0, cudaStream_t s) {
	if (!p0.M || !p0.Ncols || !p0.K) return cudaSuccess;
	cudaError_t e = cudaFuncSetAttribute(score_gemm_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TSMEM);
	if (e != cudaSuccess) return e;
	TiledScoreParams p = p0;
	p.n_kc = (p.K + KC - 1) / KC;
	const uint32_t mt = (p.M + BM - 1) / BM, nt = (p.Ncols + TBN - 1) / TBN;
	if (mt > 65535) return cudaErrorInvalidValue;
	score_gemm_tiled_kernel<<<dim3(nt, mt), 192, TSMEM, s>>>(p);
	return cudaGetLastError();
}

size_t tiled_operand_bytes(uint32_t N, uint32_t ncols, uint32_t T) { return (size_t)((N + KC - 1) / KC) * ((ncols + T - 1) / T) * 2 * ((size_t)T * 64); }

cudaError_t launch_tile_mn(const float* src, uint64_t ld, uint32_t ncols, uint32_t ones_col, uint32_t N, bool m_side, unsigned char* dst, cudaStream_t s) {
	if (!N || !ncols) return cudaSuccess;
	const uint32_t T = m_side ? BM : BN, n_ct = (ncols + T - 1) / T;      // (the bulk-copy-fed GEMMs take 128-wide tiles on both sides)
	dim3 grid((N + KC - 1) / KC, n_ct);
	if (m_side) tile_mn_kernel<BM><<<grid, 256, 0, s>>>(src, ld, ncols, ones_col, N, dst, n_ct);
	else tile_mn_kernel<BN><<<grid, 256, 0, s>>>(src, ld, ncols, ones_col, N, dst, n_ct);
	return cudaGetLastError();
}

cudaError_t launch_reduce_gemm_tiled(const TiledReduceParams& p0, cudaStream_t s) {
	if (!p0.N || !p0.I || !p0.J) return cudaSuccess;
	cudaError_t e = cudaFuncSetAttribute(reduce_gemm_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TSMEM);
	if (e != cudaSuccess) return e;
	TiledReduceParams p = p0;
	p.n_mt = (p.I + BM - 1) / BM; p.n_nt = (p.J + TBN - 1) / TBN; p.n_chunks = (p.N + KC - 1) / KC;
	// frame slabs only until the grid is about six waves of 2 CTAs per SM: every slab ends in up to 8192 fp64 atomics per CTA
	const uint32_t tiles = p.n_mt * p.n_nt, slabs = std::max(1u, std::min(p.n_chunks, (6u * 296u + tiles - 1) / tiles));
	p.slab_chunks = (p.n_chunks + slabs - 1) / slabs;
	dim3 grid(p.n_mt, p.n_nt, (p.n_chunks + p.slab_chunks - 1) / p.slab_chunks);
	reduce_gemm_tiled_kernel<<<grid, 192, TSMEM, s>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_reduce_gemm_tc(const ReduceGemmParams& p, bool m_side_is_b, cudaStream_t s) {
	if (p.n1 <= p.n0 || !p.I || !p.J) return cudaSuccess;
	bool attr_done = false;      // (function attributes are per device: no process-wide cache, a multi-GPU process configures each one)
	if (!attr_done) {
		cudaError_t e = cudaFuncSetAttribute(reduce_gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
		if (e == cudaSuccess) e = cudaFuncSetAttribute(reduce_gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
		if (e != cudaSuccess) return e;
		attr_done = true;
	}
	const uint32_t Mext = m_side_is_b ? p.J : p.I, Next = m_side_is_b ? p.I : p.J;
	dim3 grid((Mext + BM - 1) / BM, (Next + BN - 1) / BN, (p.n1 - p.n0 + p.k_slab - 1) / p.k_slab);
	if (m_side_is_b) reduce_gemm_tc_kernel<true><<<grid, 160, SMEM_BYTES, s>>>(p);
	else reduce_gemm_tc_kernel<false><<<grid, 160, SMEM_BYTES, s>>>(p);
	return cudaGetLastError();
}

}  // namespace crfgpu
