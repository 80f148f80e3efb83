// TMA-fed tensor-core GEMMs over the window-feature stream X[N][D][Wp] (sm_100a), the two kernels that move
// almost all HBM bytes of a segmental training step:
//
//   score_gemm_tma   S[n][(d,y)] = sum_k X[n][d][k] * W[(d,y)][k] + bias[(d,y)]       all durations in ONE launch
//                    replaces CRF_StdFeatureMap::computeStateArrayValue x L per frame (CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:65-81)
//                    and the per-duration score maxima the lattice kernels use as scale bounds
//   state_grad_tma   grad[sidx(d,y) + j] += sum_n Dm[n][(d,y)] * X[n][d][j]            all durations in ONE launch
//                    replaces the per-frame scatter of computeStateExpF (:130-175) and `grad -= ExpF`
//                    (CRF/src/trainers/gradbuilders/CRF_NewGradBuilder.cpp:374-376)
//
// Data path of one CTA (two CTAs per SM):
//   TMA warp        cp.async.bulk.tensor.3d boxes of raw fp32 window features (32 floats = 128 B inner, SWIZZLE_128B) into a
//                   4-deep ring: 64 KB per CTA in flight with no registers and no issue slots spent on the loads
//   8 converter warps  raw tile (conflict-free through the swizzle) -> bf16 hi/lo split -> canonical no-swizzle UMMA tiles
//   MMA warp        3 tcgen05.mma per 16-wide k-step (hi*hi + lo*hi + hi*lo) into a TMEM accumulator, commits free the stages
//   warps 0-3       epilogue
// The small operand never passes through the converters in the score kernel: the state weights are split into bf16 hi/lo UMMA
// tiles once per lambda (lambda_tables_kernel, crf_lambda.cu) and land in the operand stage with one 8 KB bulk copy per k-chunk.
#include <cuda.h>   // CUtensorMap and its enums only; cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint (no -lcuda)

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "crf_kernels.cuh"
#include "tc05.cuh"

namespace crfgpu {

using namespace tc05;

namespace {

constexpr int BM = 128, BN = 64, KC = 32;
constexpr int RA = 4, OS = 2;                                       // raw (TMA) stages, operand (UMMA) stages
constexpr uint32_t RAW_BYTES = 128 * 32 * 4;                        // one raw fp32 tile, 16 KB
constexpr uint32_t A_TILE = BM * KC * 2, B_TILE = BN * KC * 2;      // bf16 tiles of the 128-row and the 64-row operand
constexpr uint32_t OP_BYTES = 2 * A_TILE + 2 * B_TILE;              // hi + lo of both = 24 KB
constexpr uint32_t OP_OFF = RA * RAW_BYTES, CTL_OFF = OP_OFF + OS * OP_BYTES;
constexpr uint32_t SMEM_BYTES = CTL_OFF + 128;                      // 112 KB + barriers: two CTAs per SM
constexpr int CONV_WARPS = 8;

struct Ctl {
	uint64_t raw_full[RA], raw_empty[RA], op_full[OS], op_empty[OS], done;
	uint32_t tmem;
};
static_assert(sizeof(Ctl) <= 128, "control block");

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint32_t c0, uint32_t c1, uint32_t c2, uint64_t* bar) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
	             :
	             : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :
	             : "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
	asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// 8 consecutive floats of row `row` (128 B, SWIZZLE_128B: 16-byte chunk j of row r sits at chunk j ^ (r & 7)) starting at chunk 2*pair
__device__ __forceinline__ void load_raw8(const unsigned char* tile, uint32_t row, uint32_t pair, float (&x)[8]) {
	const unsigned char* r = tile + row * 128;
	const uint32_t sw = row & 7;
	const float4 a = *reinterpret_cast<const float4*>(r + (((2 * pair) ^ sw) << 4));
	const float4 b = *reinterpret_cast<const float4*>(r + (((2 * pair + 1) ^ sw) << 4));
	x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}

// the same for a 64-column tile held as two boxes when the 8 floats start `sh` (1..3, the same for the whole CTA) columns behind the
// aligned column col4 (a multiple of 4): three aligned 16-byte loads and a register selection instead of eight scalar loads with their
// own swizzle arithmetic (columns >= 64 read as 0)
__device__ __forceinline__ void load_raw8_shifted(const unsigned char* tiles, uint32_t row, uint32_t col4, uint32_t sh, float (&x)[8]) {
	const unsigned char* r = tiles + row * 128;
	const uint32_t sw = row & 7, q = col4 >> 2;                     // first aligned 16-byte chunk (0..15 over the two boxes)
	float4 v[3];
#pragma unroll
	for (uint32_t i = 0; i < 3; i++) {
		const uint32_t qi = q + i;
		v[i] = qi < 16u ? *reinterpret_cast<const float4*>(r + (qi >> 3) * 4096 + (((qi & 7) ^ sw) << 4)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}
	const float f[12] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w, v[2].x, v[2].y, v[2].z, v[2].w};
	if (sh == 1) {
#pragma unroll
		for (int e = 0; e < 8; e++) x[e] = f[e + 1];
	} else if (sh == 2) {
#pragma unroll
		for (int e = 0; e < 8; e++) x[e] = f[e + 2];
	} else {
#pragma unroll
		for (int e = 0; e < 8; e++) x[e] = f[e + 3];
	}
}

__device__ __forceinline__ void setup(Ctl* ctl, uint32_t tid, uint32_t warp, uint32_t alloc_warp, uint32_t op_full_count) {
	if (tid == 0) {
		if (smem_u32(ctl) & 127) __trap();
		for (int s = 0; s < RA; s++) { mbar_init(&ctl->raw_full[s], 1); mbar_init(&ctl->raw_empty[s], CONV_WARPS); }
		for (int s = 0; s < OS; s++) { mbar_init(&ctl->op_full[s], op_full_count); mbar_init(&ctl->op_empty[s], 1); }
		mbar_init(&ctl->done, 1);
		fence_mbar_init();
	}
	if (warp == alloc_warp) tmem_alloc(&ctl->tmem, BN);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
}

template <bool MN_MAJOR>
__device__ __forceinline__ void mma_walk(unsigned char* op, Ctl* ctl, uint32_t tmem, uint32_t n_chunks) {
	constexpr uint32_t idesc = idesc_bf16_f32(BM, BN, MN_MAJOR, MN_MAJOR);
	for (uint32_t c = 0; c < n_chunks; c++) {
		const uint32_t s = c % OS;
		mbar_wait(&ctl->op_full[s], (c / OS) & 1);
		tc_fence_after();
		const uint32_t base = smem_u32(op + s * OP_BYTES);
		if (elect_one()) {
#pragma unroll
			for (int ks = 0; ks < KC / 16; ks++) {
				const uint64_t ah = smem_desc(base + ks * 256, 128, 512), al = smem_desc(base + A_TILE + ks * 256, 128, 512);
				const uint64_t bh = smem_desc(base + 2 * A_TILE + ks * 256, 128, 512), bl = smem_desc(base + 2 * A_TILE + B_TILE + ks * 256, 128, 512);
				mma_ss(tmem, ah, bh, idesc, (c | ks) != 0);
				mma_ss(tmem, al, bh, idesc, true);
				mma_ss(tmem, ah, bl, idesc, true);
			}
			mma_commit(&ctl->op_empty[s]);
		}
		__syncwarp();
	}
	if (elect_one()) mma_commit(&ctl->done);
	__syncwarp();
}

// ------------------------------------------------------------------------------------------------
// score_gemm_tma: grid (durations, label tiles of 64 inside a duration, frame tiles of 128)
// warps 0-7 converters, 8 MMA, 9 window TMA, 10 weight-tile bulk copies
// ------------------------------------------------------------------------------------------------
constexpr int SC_THR = (CONV_WARPS + 3) * 32;

__global__ void __launch_bounds__(SC_THR, 2) score_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmX, ScoreTmaParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	Ctl* ctl = reinterpret_cast<Ctl*>(smem + CTL_OFF);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t d = blockIdx.x, jt = blockIdx.y, m0 = blockIdx.z * BM;   // durations of one frame tile run back to back: they share the base rows and the rows of S
	const uint32_t n_chunks = p.n_chunks;
	if (tid == 0 && (smem_u32(smem) & 1023)) __trap();
	setup(ctl, tid, warp, CONV_WARPS, CONV_WARPS + 1);
	const uint32_t tmem = ctl->tmem;

	if (warp < CONV_WARPS) {
		const uint32_t r8 = lane & 7, g = lane >> 3;
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t r = c % RA, s = c % OS;
			mbar_wait(&ctl->raw_full[r], (c / RA) & 1);
			const unsigned char* raw = smem + r * RAW_BYTES;
			float x0[8], x1[8];
			load_raw8(raw, warp * 8 + r8, g, x0);
			load_raw8(raw, (warp + 8) * 8 + r8, g, x1);
			uint4 h0, l0, h1, l1;
			split8(x0, h0, l0); split8(x1, h1, l1);
			if (c >= OS) mbar_wait(&ctl->op_empty[s], ((c / OS) - 1) & 1);
			unsigned char* st = smem + OP_OFF + s * OP_BYTES;
			const uint32_t o = warp * 512 + g * 128 + r8 * 16;
			*reinterpret_cast<uint4*>(st + o) = h0; *reinterpret_cast<uint4*>(st + A_TILE + o) = l0;
			*reinterpret_cast<uint4*>(st + o + 8 * 512) = h1; *reinterpret_cast<uint4*>(st + A_TILE + o + 8 * 512) = l1;
			fence_proxy_async_smem();
			__syncwarp();
			if (lane == 0) { mbar_arrive(&ctl->op_full[s]); mbar_arrive(&ctl->raw_empty[r]); }
		}
	} else if (warp == CONV_WARPS) {
		mma_walk<false>(smem + OP_OFF, ctl, tmem, n_chunks);
	} else if (warp == CONV_WARPS + 1) {
		if (lane == 0) {
			prefetch_tmap(&tmX);
			for (uint32_t c = 0; c < n_chunks; c++) {
				const uint32_t r = c % RA;
				if (c >= RA) mbar_wait(&ctl->raw_empty[r], ((c / RA) - 1) & 1);
				mbar_arrive_expect_tx(&ctl->raw_full[r], RAW_BYTES);
				tma_load_3d(smem + r * RAW_BYTES, &tmX, c * KC, d, m0, &ctl->raw_full[r]);
			}
		}
	} else {
		if (lane == 0) {
			const unsigned char* src = p.Bt + ((uint64_t)((p.shared_w ? 0u : d) * p.ntile + jt) * n_chunks) * (2 * B_TILE);
			for (uint32_t c = 0; c < n_chunks; c++) {
				const uint32_t s = c % OS;
				if (c >= OS) mbar_wait(&ctl->op_empty[s], ((c / OS) - 1) & 1);
				mbar_arrive_expect_tx(&ctl->op_full[s], 2 * B_TILE);
				bulk_g2s(smem + OP_OFF + s * OP_BYTES + 2 * A_TILE, src + (uint64_t)c * (2 * B_TILE), 2 * B_TILE, &ctl->op_full[s]);
			}
		}
	}
	// ---- epilogue: TMEM -> shared transpose -> (+bias) -> coalesced rows of S, and the row maximum of the duration block ----
	if (warp < 4) {
		mbar_wait(&ctl->done, 0);
		tc_fence_after();
		float* Cs = reinterpret_cast<float*>(smem);       // [128][65] over the raw ring, idle now
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
		__syncwarp();
		const uint32_t y0 = jt * BN, ncol = min((uint32_t)BN, p.P - y0);
		const uint32_t col0 = d * p.P + y0, bcol = (p.shared_w ? 0u : d * p.P) + y0;
		const float b0 = (p.bias && lane < ncol) ? __ldg(p.bias + bcol + lane) : 0.0f;
		const float b1 = (p.bias && lane + 32 < ncol) ? __ldg(p.bias + bcol + lane + 32) : 0.0f;
		for (uint32_t rr = 0; rr < 32; rr++) {
			const uint32_t rloc = warp * 32 + rr, gm = m0 + rloc;
			if (gm >= p.M) break;
			float mx = -INFINITY;
			float* crow = p.C + (uint64_t)gm * p.ldc + col0;
			if (lane < ncol) { const float v = Cs[rloc * 65 + lane] + b0; crow[lane] = v; mx = v; }
			if (lane + 32 < ncol) { const float v = Cs[rloc * 65 + lane + 32] + b1; crow[lane + 32] = v; mx = fmaxf(mx, v); }
			if (p.smaxd) {
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
				if (lane == 0) p.smaxd[(uint64_t)gm * p.D + d] = (d <= __ldg(p.frame_t + gm)) ? mx : -INFINITY;
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == CONV_WARPS) tmem_dealloc(tmem, BN);
}

// ------------------------------------------------------------------------------------------------
// score_gemm_tma, A operand from TENSOR MEMORY: the converters write the split window tile with tcgen05.st (thread = frame row,
// 16 consecutive features = one k-step = 8 packed columns) instead of into shared memory, so that per 16 KB of raw input the CTA moves
// 52 KB through shared memory (TMA write 16 + converter read 16 + weight tile 8 + MMA reads of the weight tile 12) instead of 92 KB,
// and the MMAs cost 44 instead of 77 cycles (tools/mma_bench.cu).  TMEM per CTA: 64 accumulator columns + 4 stages x (16 hi + 16 lo)
// = 192 -> 256 allocated, two CTAs per SM.  Shared memory: 4 x 16 KB raw + 4 x 8 KB weight tiles = 96 KB.
// warps 0-7 converters (warp w: lane quadrant w % 4, k-step w / 4), 8 MMA, 9 window TMA, 10 weight-tile bulk copies
// ------------------------------------------------------------------------------------------------
constexpr int TS = 4;                                               // operand stages (A in TMEM, B in shared memory)
constexpr uint32_t T_B_OFF = RA * RAW_BYTES, T_CTL_OFF = T_B_OFF + TS * 2 * B_TILE, T_SMEM = T_CTL_OFF + 256;
constexpr uint32_t T_ACOL = 64, T_COLS = 256;

struct TCtl {
	uint64_t raw_full[RA], raw_empty[RA], op_full[TS], op_empty[TS], done;
	uint32_t tmem;
};
static_assert(sizeof(TCtl) <= 256, "control block");

__global__ void __launch_bounds__(SC_THR, 2) score_gemm_tmem_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB, ScoreTmaParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	TCtl* ctl = reinterpret_cast<TCtl*>(smem + T_CTL_OFF);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t d = blockIdx.x, jt = blockIdx.y, m0 = blockIdx.z * BM;   // durations of one frame tile run back to back: they share the base rows and the rows of S
	const uint32_t n_chunks = p.n_chunks;
	if (tid == 0) {
		if (smem_u32(smem) & 1023) __trap();
		for (int s = 0; s < RA; s++) { mbar_init(&ctl->raw_full[s], 1); mbar_init(&ctl->raw_empty[s], CONV_WARPS); }
		for (int s = 0; s < TS; s++) { mbar_init(&ctl->op_full[s], CONV_WARPS + 1); mbar_init(&ctl->op_empty[s], 1); }
		mbar_init(&ctl->done, 1);
		fence_mbar_init();
	}
	if (warp == CONV_WARPS) tmem_alloc(&ctl->tmem, T_COLS);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = ctl->tmem;

	if (warp < CONV_WARPS) {
		const uint32_t q4 = warp & 3, ks = warp >> 2, row = q4 * 32 + lane;
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t r = c % RA, s = c % TS;
			mbar_wait(&ctl->raw_full[r], (c / RA) & 1);
			const unsigned char* raw = smem + r * RAW_BYTES;
			float x0[8], x1[8];
			load_raw8(raw, row, 2 * ks, x0);
			load_raw8(raw, row, 2 * ks + 1, x1);
			uint4 h0, l0, h1, l1;
			split8(x0, h0, l0); split8(x1, h1, l1);
			if (c >= TS) { mbar_wait(&ctl->op_empty[s], ((c / TS) - 1) & 1); tc_fence_after(); }
			const uint32_t hi[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w}, lo[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
			const uint32_t col = tmem + ((q4 * 32u) << 16) + T_ACOL + s * 32 + ks * 8;
			tmem_st8(col, hi);
			tmem_st8(col + 16, lo);
			tmem_st_wait();
			tc_fence_before();
			__syncwarp();
			if (lane == 0) { mbar_arrive(&ctl->op_full[s]); mbar_arrive(&ctl->raw_empty[r]); }
		}
	} else if (warp == CONV_WARPS) {
		constexpr uint32_t idesc = idesc_bf16_f32(BM, BN, false, false);
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t s = c % TS;
			mbar_wait(&ctl->op_full[s], (c / TS) & 1);
			tc_fence_after();
			const uint32_t bbase = smem_u32(smem + T_B_OFF + s * 2 * B_TILE), acol = tmem + T_ACOL + s * 32;
			if (elect_one()) {
#pragma unroll
				for (int ks = 0; ks < KC / 16; ks++) {
					const uint64_t bh = smem_desc(bbase + ks * 256, 128, 512), bl = smem_desc(bbase + B_TILE + ks * 256, 128, 512);
					mma_ts(tmem, acol + ks * 8, bh, idesc, (c | ks) != 0);
					mma_ts(tmem, acol + 16 + ks * 8, bh, idesc, true);
					mma_ts(tmem, acol + ks * 8, bl, idesc, true);
				}
				mma_commit(&ctl->op_empty[s]);
			}
			__syncwarp();
		}
		if (elect_one()) mma_commit(&ctl->done);
		__syncwarp();
	} else if (warp == CONV_WARPS + 1) {
		if (lane == 0) {
			prefetch_tmap(&tmX);
			if (p.virt) prefetch_tmap(&tmB);
			for (uint32_t c = 0; c < n_chunks; c++) {
				const uint32_t r = c % RA;
				if (c >= RA) mbar_wait(&ctl->raw_empty[r], ((c / RA) - 1) & 1);
				mbar_arrive_expect_tx(&ctl->raw_full[r], RAW_BYTES);
				if (!p.virt) tma_load_3d(smem + r * RAW_BYTES, &tmX, c * KC, d, m0, &ctl->raw_full[r]);
				else if (c < 5 * p.cpb) {
					// a sampled-frame block of the window = a ROW SHIFT of the (padded) base stream: the window of duration d+1 that ends on
					// frame n starts on frame n - d and its b-th sample sits steps[d][b] frames further on
					// (CRF_InFtrStream_SeqMultiWindow::sample_ftrs).  Rows before the batch are zero-filled by the TMA unit; rows of the
					// previous utterance belong to windows that do not exist (d > t) and are never read downstream.
					const uint32_t b = c / p.cpb;
					const uint32_t shift = __ldg(p.steps + d * 5 + b) - d;
					tma_load_3d(smem + r * RAW_BYTES, &tmB, (c - b * p.cpb) * KC, 0, m0 + shift, &ctl->raw_full[r]);
				} else tma_load_3d(smem + r * RAW_BYTES, &tmX, (c - 5 * p.cpb) * KC, d, m0, &ctl->raw_full[r]);      // avg | max | min, materialised
			}
		}
	} else {
		if (lane == 0) {
			const unsigned char* src = p.Bt + ((uint64_t)((p.shared_w ? 0u : d) * p.ntile + jt) * n_chunks) * (2 * B_TILE);
			for (uint32_t c = 0; c < n_chunks; c++) {
				const uint32_t s = c % TS;
				if (c >= TS) mbar_wait(&ctl->op_empty[s], ((c / TS) - 1) & 1);
				mbar_arrive_expect_tx(&ctl->op_full[s], 2 * B_TILE);
				bulk_g2s(smem + T_B_OFF + s * 2 * B_TILE, src + (uint64_t)c * (2 * B_TILE), 2 * B_TILE, &ctl->op_full[s]);
			}
		}
	}
	// ---- epilogue: thread = frame row: its 64 accumulators (+bias) and their maximum straight from tensor memory, then a shared-memory
	// transpose so that the rows of S leave as coalesced stores.  (The row maximum used to be a 5-shuffle reduction inside a per-row loop:
	// 32 dependent iterations per warp, a quarter of the CTA's lifetime with the other seven warps idle.)
	if (warp < 4) {
		const uint32_t y0 = jt * BN, ncol = min((uint32_t)BN, p.P - y0);
		const uint32_t col0 = d * p.P + y0, bcol = ((p.shared_w && !p.virt) ? 0u : d * p.P) + y0;   // virt: bias per (d,y), it carries the one-hot duration weight
		float* Cs = reinterpret_cast<float*>(smem);       // [128][65] over the raw ring, idle once the last MMA has committed
		float* sb = reinterpret_cast<float*>(smem + T_B_OFF);   // the tile's biases, over the weight ring (idle as well)
		mbar_wait(&ctl->done, 0);
		tc_fence_after();
		if (warp == 0) { sb[lane] = (p.bias && lane < ncol) ? __ldg(p.bias + bcol + lane) : 0.0f; sb[lane + 32] = (p.bias && lane + 32 < ncol) ? __ldg(p.bias + bcol + lane + 32) : 0.0f; }
		asm volatile("bar.sync 1, 128;" ::: "memory");
		const uint32_t row = warp * 32 + lane, gmr = m0 + row;
		float mx = -INFINITY;
#pragma unroll
		for (int c0 = 0; c0 < BN; c0 += 16) {
			float v[16];
			tmem_ld16(tmem + ((warp * 32u) << 16) + c0, v);
			tmem_ld_wait();
#pragma unroll
			for (int j = 0; j < 16; j++) {
				const float x = v[j] + sb[c0 + j];
				Cs[row * 65 + c0 + j] = x;
				if ((uint32_t)(c0 + j) < ncol) mx = fmaxf(mx, x);
			}
		}
		tc_fence_before();
		if (p.smaxd && gmr < p.M) p.smaxd[(uint64_t)gmr * p.D + d] = (d <= __ldg(p.frame_t + gmr)) ? mx : -INFINITY;
		__syncwarp();
		const uint32_t nrow = min(32u, p.M > m0 + warp * 32 ? p.M - (m0 + warp * 32) : 0u);
#pragma unroll 4
		for (uint32_t rr = 0; rr < nrow; rr++) {
			const uint32_t rloc = warp * 32 + rr;
			float* crow = p.C + (uint64_t)(m0 + rloc) * p.ldc + col0;
			if (lane < ncol) crow[lane] = Cs[rloc * 65 + lane];
			if (lane + 32 < ncol) crow[lane + 32] = Cs[rloc * 65 + lane + 32];
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == CONV_WARPS) tmem_dealloc(tmem, T_COLS);
}

// ------------------------------------------------------------------------------------------------
// frame_gemm_tma: products whose reduction index is the frame, both operands [frames][columns] (MN-major), both TMA-fed.
//   MODE 0  state gradient: 128-row side = window features of duration block d (3-D map, coordinate d), 64-column side = Dm columns
//           of block d.  The inner TMA coordinate must be a multiple of 16 bytes (measured: anything else faults), so the box starts at
//           the aligned column below the block (up to 3 early) and a column tile is 61 wide, which always fits the 64-wide box; the
//           converters read the tile at the residual offset, columns beyond the tile are masked in the epilogue.
//   MODE 1  Xi (CRF_StdFeatureMap::computeTransExpF summed over frames, CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:197-223):
//           128-row side = forward vectors A[n-d][q] -- the duration's row shift is just the TMA row coordinate (negative rows and
//           columns >= L are zero-filled) --, 64-column side = right factors R[n][(d,y)].
// grid (128-row tiles, (duration block, 64-column tile), frame slabs); rings: 3 x 16 KB + 2 x 8 KB raw, 2 x 24 KB operands: 112 KB,
// two CTAs per SM.  warps 0-7 converters, 8 MMA, 9 TMA
// ------------------------------------------------------------------------------------------------
constexpr int RM = 3, RN = 2;
constexpr uint32_t RAWN_BYTES = RAW_BYTES / 2;
constexpr uint32_t F_RAWN_OFF = RM * RAW_BYTES, F_OP_OFF = F_RAWN_OFF + RN * RAWN_BYTES, F_CTL_OFF = F_OP_OFF + OS * OP_BYTES, F_SMEM = F_CTL_OFF + 128;
constexpr int FG_THR = (CONV_WARPS + 2) * 32;

struct FCtl {
	uint64_t m_full[RM], m_empty[RM], n_full[RN], n_empty[RN], op_full[OS], op_empty[OS], done;
	uint32_t tmem;
};
static_assert(sizeof(FCtl) <= 128, "control block");

template <int MODE>
__global__ void __launch_bounds__(FG_THR, 2) frame_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmN, FrameGemmParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	FCtl* ctl = reinterpret_cast<FCtl*>(smem + F_CTL_OFF);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t m0 = blockIdx.x * BM, d = blockIdx.y / p.ntile, y0 = (blockIdx.y % p.ntile) * FRAME_GEMM_TILE;
	const uint32_t ns = blockIdx.z * p.k_slab, ne = min(ns + p.k_slab, p.N);
	const uint32_t n_chunks = (ne - ns + KC - 1) / KC;
	const uint32_t ncol = min((uint32_t)FRAME_GEMM_TILE, p.P - y0), col0 = d * p.P + y0, sh = col0 & 3;   // the box starts sh columns early
	if (tid == 0) {
		if (smem_u32(smem) & 1023) __trap();
		for (int s = 0; s < RM; s++) { mbar_init(&ctl->m_full[s], 1); mbar_init(&ctl->m_empty[s], CONV_WARPS); }
		for (int s = 0; s < RN; s++) { mbar_init(&ctl->n_full[s], 1); mbar_init(&ctl->n_empty[s], CONV_WARPS); }
		for (int s = 0; s < OS; s++) { mbar_init(&ctl->op_full[s], CONV_WARPS); mbar_init(&ctl->op_empty[s], 1); }
		mbar_init(&ctl->done, 1);
		fence_mbar_init();
	}
	if (warp == CONV_WARPS) tmem_alloc(&ctl->tmem, BN);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = ctl->tmem;

	if (warp < CONV_WARPS) {
		// unit = (frame k, group of 8 columns): two of the 128-row side and one of the 64-column side per thread
		const uint32_t k = (warp & 3) * 8 + (lane & 7), rg0 = (warp >> 2) * 8 + (lane >> 3) * 2, cg = (warp >> 2) * 4 + (lane >> 3);
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t rm = c % RM, rn = c % RN, s = c % OS;
			float x[3][8];
			mbar_wait(&ctl->m_full[rm], (c / RM) & 1);
			const unsigned char* rawm = smem + rm * RAW_BYTES;
#pragma unroll
			for (uint32_t i = 0; i < 2; i++) {
				const uint32_t rg = rg0 + i;
				load_raw8(rawm + (rg >> 2) * 4096, k, rg & 3, x[i]);
				if (MODE == 0) {
					const uint32_t rel = p.ones_col - (m0 + rg * 8);          // the constant-1 bias feature (TMA zero-fills beyond the window)
					if (rel < 8u) {
						const float one = (ns + c * KC + k < ne) ? 1.0f : 0.0f;
#pragma unroll
						for (uint32_t j = 0; j < 8; j++) if (j == rel) x[i][j] = one;
					}
				}
			}
			mbar_wait(&ctl->n_full[rn], (c / RN) & 1);
			if (sh == 0) load_raw8(smem + F_RAWN_OFF + rn * RAWN_BYTES + (cg >> 2) * 4096, k, cg & 3, x[2]);
			else load_raw8_shifted(smem + F_RAWN_OFF + rn * RAWN_BYTES, k, cg * 8, sh, x[2]);
			uint4 h[3], l[3];
#pragma unroll
			for (int i = 0; i < 3; i++) split8(x[i], h[i], l[i]);
			if (c >= OS) mbar_wait(&ctl->op_empty[s], ((c / OS) - 1) & 1);
			unsigned char* st = smem + F_OP_OFF + s * OP_BYTES;
#pragma unroll
			for (uint32_t i = 0; i < 2; i++) {
				const uint32_t o = (rg0 + i) * 512 + k * 16;
				*reinterpret_cast<uint4*>(st + o) = h[i]; *reinterpret_cast<uint4*>(st + A_TILE + o) = l[i];
			}
			{
				const uint32_t o = cg * 512 + k * 16;
				*reinterpret_cast<uint4*>(st + 2 * A_TILE + o) = h[2]; *reinterpret_cast<uint4*>(st + 2 * A_TILE + B_TILE + o) = l[2];
			}
			fence_proxy_async_smem();
			__syncwarp();
			if (lane == 0) { mbar_arrive(&ctl->op_full[s]); mbar_arrive(&ctl->m_empty[rm]); mbar_arrive(&ctl->n_empty[rn]); }
		}
	} else if (warp == CONV_WARPS) {
		constexpr uint32_t idesc = idesc_bf16_f32(BM, BN, true, true);
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t s = c % OS;
			mbar_wait(&ctl->op_full[s], (c / OS) & 1);
			tc_fence_after();
			const uint32_t base = smem_u32(smem + F_OP_OFF + s * OP_BYTES);
			if (elect_one()) {
#pragma unroll
				for (int ks = 0; ks < KC / 16; ks++) {
					const uint64_t ah = smem_desc(base + ks * 256, 128, 512), al = smem_desc(base + A_TILE + ks * 256, 128, 512);
					const uint64_t bh = smem_desc(base + 2 * A_TILE + ks * 256, 128, 512), bl = smem_desc(base + 2 * A_TILE + B_TILE + ks * 256, 128, 512);
					mma_ss(tmem, ah, bh, idesc, (c | ks) != 0);
					mma_ss(tmem, al, bh, idesc, true);
					mma_ss(tmem, ah, bl, idesc, true);
				}
				mma_commit(&ctl->op_empty[s]);
			}
			__syncwarp();
		}
		if (elect_one()) mma_commit(&ctl->done);
		__syncwarp();
	} else {
		if (lane == 0) {
			prefetch_tmap(&tmM); prefetch_tmap(&tmN);
			for (uint32_t c = 0; c < n_chunks; c++) {
				const uint32_t rm = c % RM, rn = c % RN, n = ns + c * KC;
				if (c >= RM) mbar_wait(&ctl->m_empty[rm], ((c / RM) - 1) & 1);
				mbar_arrive_expect_tx(&ctl->m_full[rm], RAW_BYTES);
#pragma unroll
				for (uint32_t b = 0; b < 4; b++)
					tma_load_3d(smem + rm * RAW_BYTES + b * 4096, &tmM, m0 + b * 32, MODE == 0 ? d : 0u, MODE == 0 ? n : n - (d + 1), &ctl->m_full[rm]);
				if (c >= RN) mbar_wait(&ctl->n_empty[rn], ((c / RN) - 1) & 1);
				mbar_arrive_expect_tx(&ctl->n_full[rn], RAWN_BYTES);
#pragma unroll
				for (uint32_t b = 0; b < 2; b++) tma_load_3d(smem + F_RAWN_OFF + rn * RAWN_BYTES + b * 4096, &tmN, (col0 & ~3u) + b * 32, 0, n, &ctl->n_full[rn]);
			}
		}
	}
	// ---- epilogue: lane = row of the 128-row side, 64 columns of the duration block; fp64 atomics into the gradient ----
	if (warp < 4 && n_chunks) {
		mbar_wait(&ctl->done, 0);
		tc_fence_after();
		const uint32_t gm = m0 + warp * 32 + lane;
		const double sc = (MODE == 0 && gm == p.ones_col) ? p.ones_scale : p.scale;
#pragma unroll
		for (int c0 = 0; c0 < BN; c0 += 16) {
			float v[16];
			tmem_ld16(tmem + ((warp * 32u) << 16) + c0, v);
			tmem_ld_wait();
			if (gm < p.Mext) {
#pragma unroll
				for (int j = 0; j < 16; j++) {
					const uint32_t y = c0 + j;
					if (y >= ncol || v[j] == 0.0f) continue;
					if (MODE == 0) atomicAdd(&p.out[(uint64_t)__ldg(p.row_idx + col0 + y) + gm], sc * (double)v[j]);
					else {
						const uint32_t idx = __ldg(p.pair_idx + (uint64_t)gm * p.L + col0 + y);
						if (idx != 0xffffffffu) atomicAdd(&p.out[idx], sc * (double)__ldg(p.Ew + (uint64_t)gm * p.e_ld + col0 + y) * (double)v[j]);
					}
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == CONV_WARPS) tmem_dealloc(tmem, BN);
}

// ------------------------------------------------------------------------------------------------
// frame_gemm, 128-row operand through TENSOR MEMORY (same idea as score_gemm_tmem_kernel): converter thread = one row of the
// 128-row side (a window feature / a source label), 16 frames = one k-step; it gathers its 16 values down the raw [frame][column]
// boxes (a warp reads 32 consecutive columns of one frame: conflict-free through the swizzle) and writes them with tcgen05.st.
// The 64-column operand still goes through the converters into shared memory (MN-major).  TMEM per CTA: 64 + 4 x 32 = 192 -> 256
// columns; shared memory 3 x 16 + 2 x 8 KB raw + 4 x 8 KB operand = 96 KB; two CTAs per SM.
// warps 0-7 converters (warp w: lane quadrant w % 4, k-step w / 4), 8 MMA, 9 TMA
// ------------------------------------------------------------------------------------------------
constexpr uint32_t G_RAWN_OFF = RM * RAW_BYTES, G_OP_OFF = G_RAWN_OFF + RN * RAWN_BYTES, G_CTL_OFF = G_OP_OFF + TS * 2 * B_TILE, G_SMEM = G_CTL_OFF + 256;

struct GCtl {
	uint64_t m_full[RM], m_empty[RM], n_full[RN], n_empty[RN], op_full[TS], op_empty[TS], done;
	uint32_t tmem;
};
static_assert(sizeof(GCtl) <= 256, "control block");

template <int MODE>
__global__ void __launch_bounds__(FG_THR, 2) frame_gemm_tmem_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmN, const __grid_constant__ CUtensorMap tmB, FrameGemmParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	GCtl* ctl = reinterpret_cast<GCtl*>(smem + G_CTL_OFF);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t d = blockIdx.y / p.ntile, y0 = (blockIdx.y % p.ntile) * FRAME_GEMM_TILE;
	// virtual windows (MODE 0, p.virt): the five sampled-frame blocks are walked in UNITS of 32 features -- unit u = 32 consecutive features
	// of block u / upb (a row shift of the padded base stream, upb = Fp / 32 units per block); a 128-row tile holds four consecutive units
	// (each of its four TMA boxes has its own block and shift, so narrow streams waste no tile: F = 64 -> 10 units = 3 tiles, not 5), the
	// first p.tpb tiles are sampled tiles, the others rows of the materialised avg | max | min array.  Per warp quadrant / epilogue warp
	// (32 rows = one unit): lam_row = index of its row 0 inside the label's state-weight block of lambda, vrows = its real feature rows.
	uint32_t m0 = blockIdx.x * BM, vb = 0xffffffffu, lam_row = m0 + (warp & 3) * 32, vrows = 32;
	const uint32_t upb = (MODE == 0 && p.virt) ? p.Fp / 32 : 1;
	if (MODE == 0 && p.virt) {
		const uint32_t mt = blockIdx.x;
		if (mt < p.tpb) {
			const uint32_t u = mt * 4 + (warp & 3), blk = u / upb, foff = (u - blk * upb) * 32;
			vb = 0;                                                              // a sampled tile
			lam_row = blk * p.F + foff; vrows = (blk < 5 && p.F > foff) ? min(32u, p.F - foff) : 0u;
		} else {
			m0 = (mt - p.tpb) * BM;
			const uint32_t r0 = m0 + (warp & 3) * 32;
			lam_row = 5 * p.F + r0; vrows = 3 * p.F > r0 ? min(32u, 3 * p.F - r0) : 0u;
		}
	}
	const uint32_t ns = blockIdx.z * p.k_slab, ne = min(ns + p.k_slab, p.N);
	const uint32_t n_chunks = (ne - ns + KC - 1) / KC;
	const uint32_t ncol = min((uint32_t)FRAME_GEMM_TILE, p.P - y0), col0 = d * p.P + y0, sh = col0 & 3;
	if (tid == 0) {
		if (smem_u32(smem) & 1023) __trap();
		for (int s = 0; s < RM; s++) { mbar_init(&ctl->m_full[s], 1); mbar_init(&ctl->m_empty[s], CONV_WARPS); }
		for (int s = 0; s < RN; s++) { mbar_init(&ctl->n_full[s], 1); mbar_init(&ctl->n_empty[s], CONV_WARPS); }
		for (int s = 0; s < TS; s++) { mbar_init(&ctl->op_full[s], CONV_WARPS); mbar_init(&ctl->op_empty[s], 1); }
		mbar_init(&ctl->done, 1);
		fence_mbar_init();
	}
	if (warp == CONV_WARPS) tmem_alloc(&ctl->tmem, T_COLS);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = ctl->tmem;

	if (warp < CONV_WARPS) {
		const uint32_t q4 = warp & 3, ks = warp >> 2;
		const uint32_t gm = m0 + q4 * 32 + lane;                                      // my row of the 128-row side
		const uint32_t k = (warp & 3) * 8 + (lane & 7), cg = (warp >> 2) * 4 + (lane >> 3);   // my unit of the 64-column side
		const bool is_ones = MODE == 0 && (p.virt ? (vb == 0xffffffffu && gm == 3 * p.F) : gm == p.ones_col);
		// byte offset of my column inside raw row kk of the swizzled box: 16-byte chunk (lane >> 2) ^ (kk & 7), word lane & 3
		uint32_t xoff[8];
#pragma unroll
		for (uint32_t j = 0; j < 8; j++) xoff[j] = (((lane >> 2) ^ j) << 4) + (lane & 3) * 4;
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t rm = c % RM, rn = c % RN, s = c % TS;
			mbar_wait(&ctl->m_full[rm], (c / RM) & 1);
			const unsigned char* box = smem + rm * RAW_BYTES + q4 * 4096;
			float x[16];
#pragma unroll
			for (uint32_t j = 0; j < 16; j++) {
				const uint32_t kk = ks * 16 + j;
				x[j] = *reinterpret_cast<const float*>(box + kk * 128 + xoff[j & 7]);
			}
			if (is_ones) {                                                              // one lane of one warp of the tile that holds the bias row
#pragma unroll
				for (uint32_t j = 0; j < 16; j++) x[j] = (ns + c * KC + ks * 16 + j < ne) ? 1.0f : 0.0f;   // the constant-1 bias feature (TMA zero-fills beyond the window)
			}
			uint32_t hi[8], lo[8];
#pragma unroll
			for (int j = 0; j < 8; j++) split2(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
			float xn[8];
			mbar_wait(&ctl->n_full[rn], (c / RN) & 1);
			if (sh == 0) load_raw8(smem + G_RAWN_OFF + rn * RAWN_BYTES + (cg >> 2) * 4096, k, cg & 3, xn);
			else load_raw8_shifted(smem + G_RAWN_OFF + rn * RAWN_BYTES, k, cg * 8, sh, xn);
			uint4 hn, ln;
			split8(xn, hn, ln);
			if (c >= TS) { mbar_wait(&ctl->op_empty[s], ((c / TS) - 1) & 1); tc_fence_after(); }
			const uint32_t col = tmem + ((q4 * 32u) << 16) + T_ACOL + s * 32 + ks * 8;
			tmem_st8(col, hi);
			tmem_st8(col + 16, lo);
			unsigned char* st = smem + G_OP_OFF + s * 2 * B_TILE;
			const uint32_t o = cg * 512 + k * 16;
			*reinterpret_cast<uint4*>(st + o) = hn; *reinterpret_cast<uint4*>(st + B_TILE + o) = ln;
			fence_proxy_async_smem();
			tmem_st_wait();
			tc_fence_before();
			__syncwarp();
			if (lane == 0) { mbar_arrive(&ctl->op_full[s]); mbar_arrive(&ctl->m_empty[rm]); mbar_arrive(&ctl->n_empty[rn]); }
		}
	} else if (warp == CONV_WARPS) {
		constexpr uint32_t idesc = idesc_bf16_f32(BM, BN, false, true);
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t s = c % TS;
			mbar_wait(&ctl->op_full[s], (c / TS) & 1);
			tc_fence_after();
			const uint32_t bbase = smem_u32(smem + G_OP_OFF + s * 2 * B_TILE), acol = tmem + T_ACOL + s * 32;
			if (elect_one()) {
#pragma unroll
				for (int ks = 0; ks < KC / 16; ks++) {
					const uint64_t bh = smem_desc(bbase + ks * 256, 128, 512), bl = smem_desc(bbase + B_TILE + ks * 256, 128, 512);
					mma_ts(tmem, acol + ks * 8, bh, idesc, (c | ks) != 0);
					mma_ts(tmem, acol + 16 + ks * 8, bh, idesc, true);
					mma_ts(tmem, acol + ks * 8, bl, idesc, true);
				}
				mma_commit(&ctl->op_empty[s]);
			}
			__syncwarp();
		}
		if (elect_one()) mma_commit(&ctl->done);
		__syncwarp();
	} else {
		if (lane == 0) {
			prefetch_tmap(&tmM); prefetch_tmap(&tmN);
			// sampled tiles: unit of each of the four boxes -> (feature offset, row shift); units behind the fifth block read beyond the
			// stream's last row, which the TMA unit zero-fills
			uint32_t ufoff[4] = {0, 0, 0, 0}, usft[4] = {0, 0, 0, 0}; bool ulive[4] = {false, false, false, false};
			if (MODE == 0 && p.virt && vb != 0xffffffffu) {
				prefetch_tmap(&tmB);
#pragma unroll
				for (uint32_t b = 0; b < 4; b++) {
					const uint32_t u = blockIdx.x * 4 + b, blk = u / upb;
					ufoff[b] = (u - blk * upb) * 32; ulive[b] = blk < 5;
					usft[b] = ulive[b] ? __ldg(p.steps + d * 5 + blk) - d : 0u;
				}
			}
			for (uint32_t c = 0; c < n_chunks; c++) {
				const uint32_t rm = c % RM, rn = c % RN, n = ns + c * KC;
				if (c >= RM) mbar_wait(&ctl->m_empty[rm], ((c / RM) - 1) & 1);
				mbar_arrive_expect_tx(&ctl->m_full[rm], RAW_BYTES);
#pragma unroll
				for (uint32_t b = 0; b < 4; b++) {
					if (MODE == 0 && p.virt && vb != 0xffffffffu) tma_load_3d(smem + rm * RAW_BYTES + b * 4096, &tmB, ufoff[b], 0, ulive[b] ? n + usft[b] : p.N, &ctl->m_full[rm]);
					else tma_load_3d(smem + rm * RAW_BYTES + b * 4096, &tmM, m0 + b * 32, MODE == 0 ? d : 0u, MODE == 0 ? n : n - (d + 1), &ctl->m_full[rm]);
				}
				if (c >= RN) mbar_wait(&ctl->n_empty[rn], ((c / RN) - 1) & 1);
				mbar_arrive_expect_tx(&ctl->n_full[rn], RAWN_BYTES);
#pragma unroll
				for (uint32_t b = 0; b < 2; b++) tma_load_3d(smem + G_RAWN_OFF + rn * RAWN_BYTES + b * 4096, &tmN, (col0 & ~3u) + b * 32, 0, n, &ctl->n_full[rn]);
			}
		}
	}
	// ---- epilogue: lane = row of the 128-row side, 64 columns of the duration block; fp64 atomics into the gradient ----
	if (warp < 4 && n_chunks) {
		mbar_wait(&ctl->done, 0);
		tc_fence_after();
		const uint32_t gm = m0 + warp * 32 + lane;
		const bool v_ones = MODE == 0 && p.virt && vb == 0xffffffffu && gm == 3 * p.F;      // sum_n Dm: the bias count AND the one-hot duration count
		const bool row_ok = (MODE == 0 && p.virt) ? (lane < vrows || v_ones) : gm < p.Mext;
		const uint32_t lrow = (MODE == 0 && p.virt) ? (v_ones ? p.ones_col : lam_row + lane) : gm;
		const double sc = (MODE == 0 && (p.virt ? v_ones : gm == p.ones_col)) ? p.ones_scale : p.scale;
#pragma unroll
		for (int c0 = 0; c0 < BN; c0 += 16) {
			float v[16];
			tmem_ld16(tmem + ((warp * 32u) << 16) + c0, v);
			tmem_ld_wait();
			if (row_ok) {
#pragma unroll
				for (int j = 0; j < 16; j++) {
					const uint32_t y = c0 + j;
					if (y >= ncol || v[j] == 0.0f) continue;
					if (MODE == 0) {
						const uint64_t rb = __ldg(p.row_idx + col0 + y);
						if (!v_ones || p.ones_col != 0xffffffffu) atomicAdd(&p.out[rb + lrow], sc * (double)v[j]);
						if (v_ones) atomicAdd(&p.out[rb + 8 * p.F + d], p.scale * (double)v[j]);          // the window's one-hot duration feature is 1 exactly at d
					}
					else {
						const uint32_t idx = __ldg(p.pair_idx + (uint64_t)gm * p.L + col0 + y);
						if (idx != 0xffffffffu) atomicAdd(&p.out[idx], sc * (double)__ldg(p.Ew + (uint64_t)gm * p.e_ld + col0 + y) * (double)v[j]);
					}
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == CONV_WARPS) tmem_dealloc(tmem, T_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encoder() {
	static EncodeTiledFn fn = nullptr;
	if (!fn) {
		void* sym = nullptr;
		cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
			fn = reinterpret_cast<EncodeTiledFn>(sym);
	}
	return fn;
}

// X viewed as [N][D][K] floats, window stride Wp; box = [box_n frames][1 duration][32 features], 128-byte swizzle, zero fill
bool window_map(CUtensorMap* tm, const float* X, uint32_t N, uint32_t D, uint32_t Wp, uint32_t K, uint32_t box_n, bool wide_promotion) {
	EncodeTiledFn enc = encoder();
	if (!enc) return false;
	const cuuint64_t dims[3] = {K, D, N};
	const cuuint64_t strides[2] = {(cuuint64_t)Wp * 4, (cuuint64_t)D * Wp * 4};
	const cuuint32_t box[3] = {32, 1, box_n}, estr[3] = {1, 1, 1};
	return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(X), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	           CU_TENSOR_MAP_SWIZZLE_128B, wide_promotion ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
	           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool tma_gemm_eligible(const float* X, uint32_t D, uint32_t Wp, uint32_t sf0) {
	return encoder() != nullptr && D > 1 && Wp % 4 == 0 && sf0 % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
}

uint32_t score_tma_chunks(uint32_t K) { return (K + KC - 1) / KC; }

cudaError_t launch_score_gemm_tma(const float* X, uint32_t Wp, const ScoreTmaParams& p, cudaStream_t s) {
	if (!p.M || !p.P) return cudaSuccess;
	bool attr_done = false;      // (function attributes are per device: no process-wide cache, a multi-GPU process configures each one)
	if (!attr_done) {
		cudaError_t e = cudaFuncSetAttribute(score_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
		if (e != cudaSuccess) return e;
		attr_done = true;
	}
	CUtensorMap tm, tb;
	// virtual windows: X is the aggregate array [M][D][Wp] (avg | max | min), base2 the padded base stream [M][cpb * 32]
	if (!window_map(&tm, X, p.M, p.D, Wp, p.virt ? Wp : p.K, BM, true)) return cudaErrorInvalidValue;
	if (p.virt) { if (!p.a_from_tmem || !window_map(&tb, p.base2, p.M, 1, p.cpb * KC, p.cpb * KC, BM, false)) return cudaErrorInvalidValue; }
	else tb = tm;
	dim3 grid(p.D, p.ntile, (p.M + BM - 1) / BM);
	if (grid.z > 65535u) return cudaErrorInvalidValue;      // 8.3 M frames per launch
	if (p.a_from_tmem) {
		bool attr2 = false;
		if (!attr2) {
			cudaError_t e = cudaFuncSetAttribute(score_gemm_tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM);
			if (e != cudaSuccess) return e;
			attr2 = true;
		}
		score_gemm_tmem_kernel<<<grid, SC_THR, T_SMEM, s>>>(tm, tb, p);
	} else score_gemm_tma_kernel<<<grid, SC_THR, SMEM_BYTES, s>>>(tm, p);
	return cudaGetLastError();
}

static cudaError_t frame_gemm_attrs() {
	bool attr_done = false;      // (function attributes are per device: no process-wide cache, a multi-GPU process configures each one)
	if (!attr_done) {
		cudaError_t e = cudaFuncSetAttribute(frame_gemm_tma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM);
		if (e == cudaSuccess) e = cudaFuncSetAttribute(frame_gemm_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM);
		if (e == cudaSuccess) e = cudaFuncSetAttribute(frame_gemm_tmem_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
		if (e == cudaSuccess) e = cudaFuncSetAttribute(frame_gemm_tmem_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
		if (e != cudaSuccess) return e;
		attr_done = true;
	}
	return cudaSuccess;
}

bool lattice_tma_eligible(const float* a, uint32_t ld) { return encoder() != nullptr && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0; }

cudaError_t launch_state_grad_tma(const float* X, uint32_t Wp, uint32_t K, const float* Dm, uint32_t ldd, const FrameGemmParams& p, cudaStream_t s) {
	if (!p.N || !p.P || !p.Mext) return cudaSuccess;
	if (p.k_slab % KC) return cudaErrorInvalidValue;
	cudaError_t e = frame_gemm_attrs();
	if (e != cudaSuccess) return e;
	// the Dm view ends at column D*P: what a box reads beyond the last block is zero-filled
	CUtensorMap tm, tn, tb;
	if (!window_map(&tm, X, p.N, p.D, Wp, p.virt ? Wp : K, KC, false) || !window_map(&tn, Dm, p.N, 1, ldd, p.D * p.P, KC, false)) return cudaErrorInvalidValue;
	uint32_t mtiles = (p.Mext + BM - 1) / BM;
	if (p.virt) {
		// virtual windows: X = aggregate array (avg | max | min, the constant-1 row right behind), base2 = padded base stream [N][Fp]
		if (!p.a_from_tmem || !window_map(&tb, p.base2, p.N, 1, p.Fp, p.Fp, KC, false)) return cudaErrorInvalidValue;
		mtiles = p.tpb + (3 * p.F + 1 + BM - 1) / BM;         // p.tpb = sampled tiles = ceil(5 * (Fp / 32) / 4)
	} else tb = tm;
	dim3 grid(mtiles, p.D * p.ntile, (p.N + p.k_slab - 1) / p.k_slab);
	if (p.a_from_tmem) frame_gemm_tmem_kernel<0><<<grid, FG_THR, G_SMEM, s>>>(tm, tn, tb, p);
	else frame_gemm_tma_kernel<0><<<grid, FG_THR, F_SMEM, s>>>(tm, tn, p);
	return cudaGetLastError();
}

cudaError_t launch_xi_gemm_tma(const float* A, const float* R, uint32_t ld, const FrameGemmParams& p, cudaStream_t s) {
	if (!p.N || !p.L) return cudaSuccess;
	if (p.k_slab % KC) return cudaErrorInvalidValue;
	cudaError_t e = frame_gemm_attrs();
	if (e != cudaSuccess) return e;
	// [N][1][L] views (row stride ld): columns >= L are never read (zero fill), so the pad of the lattice arrays may hold anything
	CUtensorMap ta, tr;
	if (!window_map(&ta, A, p.N, 1, ld, p.L, KC, false) || !window_map(&tr, R, p.N, 1, ld, p.L, KC, false)) return cudaErrorInvalidValue;
	dim3 grid((p.Mext + BM - 1) / BM, p.D * p.ntile, (p.N + p.k_slab - 1) / p.k_slab);
	if (p.a_from_tmem) frame_gemm_tmem_kernel<1><<<grid, FG_THR, G_SMEM, s>>>(ta, tr, ta, p);
	else frame_gemm_tma_kernel<1><<<grid, FG_THR, F_SMEM, s>>>(ta, tr, p);
	return cudaGetLastError();
}

}  // namespace crfgpu
