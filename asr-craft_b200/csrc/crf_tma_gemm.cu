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
// tiles once per lambda (crfgpu_set_lambda) and land in the operand stage with one 8 KB bulk copy per k-chunk.
#include <cuda.h>   // CUtensorMap and its enums only; cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint (no -lcuda)

#include <cmath>
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
__device__ __forceinline__ void mma_walk(unsigned char* smem, Ctl* ctl, uint32_t tmem, uint32_t n_chunks) {
	constexpr uint32_t idesc = idesc_bf16_f32(BM, BN, MN_MAJOR, MN_MAJOR);
	for (uint32_t c = 0; c < n_chunks; c++) {
		const uint32_t s = c % OS;
		mbar_wait(&ctl->op_full[s], (c / OS) & 1);
		tc_fence_after();
		const uint32_t base = smem_u32(smem + OP_OFF + s * OP_BYTES);
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
// score_gemm_tma: grid (frame tiles of 128, label tiles of 64 inside a duration, durations)
// warps 0-7 converters, 8 MMA, 9 window TMA, 10 weight-tile bulk copies
// ------------------------------------------------------------------------------------------------
constexpr int SC_THR = (CONV_WARPS + 3) * 32;

__global__ void __launch_bounds__(SC_THR, 2) score_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmX, ScoreTmaParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	Ctl* ctl = reinterpret_cast<Ctl*>(smem + CTL_OFF);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t m0 = blockIdx.x * BM, jt = blockIdx.y, d = blockIdx.z;
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
		mma_walk<false>(smem, ctl, tmem, n_chunks);
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
			const unsigned char* src = p.Bt + ((uint64_t)(d * p.ntile + jt) * n_chunks) * (2 * B_TILE);
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
		const uint32_t col0 = d * p.P + y0;
		const float b0 = (p.bias && lane < ncol) ? __ldg(p.bias + col0 + lane) : 0.0f;
		const float b1 = (p.bias && lane + 32 < ncol) ? __ldg(p.bias + col0 + lane + 32) : 0.0f;
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
// state_grad_tma: grid (feature tiles of 128, (duration, label tile of 64), frame slabs)
// MMA M side = window features (MN-major, frames are the reduction index), N side = Dm columns of the duration block.
// warps 0-7 converters of the TMA'd feature tile, 8-11 register loaders of the Dm tile (its columns start on 4-byte boundaries
// only; two groups of two warps take alternate chunks so that every load has two chunk-times to land), 12 MMA, 13 TMA
// ------------------------------------------------------------------------------------------------
constexpr int LOAD_WARPS = 4;
constexpr int SG_THR = (CONV_WARPS + LOAD_WARPS + 2) * 32;

__global__ void __launch_bounds__(SG_THR, 2) state_grad_tma_kernel(const __grid_constant__ CUtensorMap tmX, StateGradTmaParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	Ctl* ctl = reinterpret_cast<Ctl*>(smem + CTL_OFF);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t m0 = blockIdx.x * BM, d = blockIdx.y / p.ntile, jt = blockIdx.y % p.ntile;
	const uint32_t ns = blockIdx.z * p.k_slab, ne = min(ns + p.k_slab, p.N);
	const uint32_t n_chunks = (ne - ns + KC - 1) / KC;
	if (tid == 0 && (smem_u32(smem) & 1023)) __trap();
	setup(ctl, tid, warp, CONV_WARPS + LOAD_WARPS, CONV_WARPS + 2);
	const uint32_t tmem = ctl->tmem;
	const uint32_t y0 = jt * BN, ncol = min((uint32_t)BN, p.P - y0);

	if (warp < CONV_WARPS) {
		// raw stage = 4 boxes [32 frames][32 features]; unit = (frame k, feature group rg of 8)
		const uint32_t k = (warp & 3) * 8 + (lane & 7), rg0 = (warp >> 2) * 8 + (lane >> 3) * 2;
		for (uint32_t c = 0; c < n_chunks; c++) {
			const uint32_t r = c % RA, s = c % OS;
			mbar_wait(&ctl->raw_full[r], (c / RA) & 1);
			const unsigned char* raw = smem + r * RAW_BYTES;
			float x[2][8];
#pragma unroll
			for (uint32_t i = 0; i < 2; i++) {
				const uint32_t rg = rg0 + i;
				load_raw8(raw + (rg >> 2) * 4096, k, rg & 3, x[i]);
				const uint32_t rel = p.ones_col - (m0 + rg * 8);          // the constant-1 bias feature (TMA zero-fills beyond the window)
				if (rel < 8u) {
					const float one = (ns + c * KC + k < ne) ? 1.0f : 0.0f;
#pragma unroll
					for (uint32_t j = 0; j < 8; j++) if (j == rel) x[i][j] = one;
				}
			}
			uint4 h[2], l[2];
			split8(x[0], h[0], l[0]); split8(x[1], h[1], l[1]);
			if (c >= OS) mbar_wait(&ctl->op_empty[s], ((c / OS) - 1) & 1);
			unsigned char* st = smem + OP_OFF + s * OP_BYTES;
#pragma unroll
			for (uint32_t i = 0; i < 2; i++) {
				const uint32_t o = (rg0 + i) * 512 + k * 16;
				*reinterpret_cast<uint4*>(st + o) = h[i]; *reinterpret_cast<uint4*>(st + A_TILE + o) = l[i];
			}
			fence_proxy_async_smem();
			__syncwarp();
			if (lane == 0) { mbar_arrive(&ctl->op_full[s]); mbar_arrive(&ctl->raw_empty[r]); }
		}
	} else if (warp < CONV_WARPS + LOAD_WARPS) {
		// Dm tile [32 frames][64 columns]: unit = (frame, column group of 8); 4 units per thread; group lg owns chunks c = lg (mod 2)
		const uint32_t lg = (warp - CONV_WARPS) >> 1, lw = (warp - CONV_WARPS) & 1, k8 = lane & 7, cq = lane >> 3;
		const float* colp = p.Dm + (uint64_t)d * p.P + y0;
		const bool vec2 = ((reinterpret_cast<uintptr_t>(colp) & 7) == 0) && (p.ldd % 2 == 0);
		float x[4][8];
		auto fetch = [&](uint32_t c) {
#pragma unroll
			for (uint32_t it = 0; it < 4; it++) {
				const uint32_t cg = cq + 4 * (it & 1), kk = (lw * 2 + (it >> 1)) * 8 + k8, n = ns + c * KC + kk;
				const float* src = colp + (uint64_t)n * p.ldd + cg * 8;
				if (n < ne && cg * 8 + 8 <= ncol) {
					if (vec2) {
#pragma unroll
						for (int j = 0; j < 4; j++) { const float2 v = __ldg(reinterpret_cast<const float2*>(src) + j); x[it][2 * j] = v.x; x[it][2 * j + 1] = v.y; }
					} else {
#pragma unroll
						for (int j = 0; j < 8; j++) x[it][j] = __ldg(src + j);
					}
				} else {
#pragma unroll
					for (uint32_t j = 0; j < 8; j++) x[it][j] = (n < ne && cg * 8 + j < ncol) ? __ldg(src + j) : 0.0f;
				}
			}
		};
		if (lg < n_chunks) fetch(lg);
		for (uint32_t c = lg; c < n_chunks; c += 2) {
			const uint32_t s = c % OS;
			if (c >= OS) mbar_wait(&ctl->op_empty[s], ((c / OS) - 1) & 1);
			unsigned char* st = smem + OP_OFF + s * OP_BYTES + 2 * A_TILE;
#pragma unroll
			for (uint32_t it = 0; it < 4; it++) {
				const uint32_t cg = cq + 4 * (it & 1), kk = (lw * 2 + (it >> 1)) * 8 + k8;
				const uint32_t o = cg * 512 + kk * 16;
				uint4 h, l;
				split8(x[it], h, l);
				*reinterpret_cast<uint4*>(st + o) = h; *reinterpret_cast<uint4*>(st + B_TILE + o) = l;
			}
			fence_proxy_async_smem();
			__syncwarp();
			if (lane == 0) mbar_arrive(&ctl->op_full[s]);
			if (c + 2 < n_chunks) fetch(c + 2);
		}
	} else if (warp == CONV_WARPS + LOAD_WARPS) {
		mma_walk<true>(smem, ctl, tmem, n_chunks);
	} else {
		if (lane == 0) {
			prefetch_tmap(&tmX);
			for (uint32_t c = 0; c < n_chunks; c++) {
				const uint32_t r = c % RA;
				if (c >= RA) mbar_wait(&ctl->raw_empty[r], ((c / RA) - 1) & 1);
				mbar_arrive_expect_tx(&ctl->raw_full[r], RAW_BYTES);
#pragma unroll
				for (uint32_t b = 0; b < 4; b++) tma_load_3d(smem + r * RAW_BYTES + b * 4096, &tmX, m0 + b * 32, d, ns + c * KC, &ctl->raw_full[r]);
			}
		}
	}
	// ---- epilogue: lane = feature j, 64 labels of the duration block; fp64 atomics into the gradient ----
	if (warp < 4 && n_chunks) {
		mbar_wait(&ctl->done, 0);
		tc_fence_after();
		const uint32_t gj = m0 + warp * 32 + lane;
		const double sc = (gj == p.ones_col) ? p.ones_scale : p.scale;
#pragma unroll
		for (int c0 = 0; c0 < BN; c0 += 16) {
			float v[16];
			tmem_ld16(tmem + ((warp * 32u) << 16) + c0, v);
			tmem_ld_wait();
			if (gj < p.J) {
#pragma unroll
				for (int j = 0; j < 16; j++) {
					const uint32_t y = c0 + j;
					if (y >= ncol || v[j] == 0.0f) continue;
					atomicAdd(&p.out[(uint64_t)__ldg(p.row_idx + d * p.P + y0 + y) + gj], sc * (double)v[j]);
				}
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == CONV_WARPS + LOAD_WARPS) tmem_dealloc(tmem, BN);
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

inline uint16_t bf16_rn(float x) {
	uint32_t u; memcpy(&u, &x, 4);
	if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);
	u += 0x7fffu + ((u >> 16) & 1u);
	return (uint16_t)(u >> 16);
}
inline float bf16_f(uint16_t h) { const uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

}  // namespace

bool tma_gemm_eligible(const float* X, uint32_t D, uint32_t Wp, uint32_t sf0) {
	return encoder() != nullptr && D > 1 && Wp % 4 == 0 && sf0 % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
}

uint32_t score_tma_chunks(uint32_t K) { return (K + KC - 1) / KC; }

void split_weight_tiles(const float* Ws, uint32_t P, uint32_t D, uint32_t K, std::vector<unsigned char>* out) {
	const uint32_t ntile = (P + BN - 1) / BN, n_chunks = score_tma_chunks(K);
	out->assign((size_t)D * ntile * n_chunks * 2 * B_TILE, 0);
	for (uint32_t d = 0; d < D; d++)
		for (uint32_t jt = 0; jt < ntile; jt++)
			for (uint32_t c = 0; c < n_chunks; c++) {
				unsigned char* tile = out->data() + ((size_t)(d * ntile + jt) * n_chunks + c) * 2 * B_TILE;
				for (uint32_t r = 0; r < (uint32_t)BN; r++) {
					const uint32_t y = jt * BN + r;
					if (y >= P) break;
					const float* w = Ws + (size_t)(d * P + y) * K;
					for (uint32_t kk = 0; kk < (uint32_t)KC; kk++) {
						const uint32_t k = c * KC + kk;
						if (k >= K) break;
						const uint16_t hi = bf16_rn(w[k]), lo = bf16_rn(w[k] - bf16_f(hi));
						const size_t o = (size_t)(r / 8) * 512 + (kk / 8) * 128 + (r % 8) * 16 + (kk % 8) * 2;
						memcpy(tile + o, &hi, 2); memcpy(tile + B_TILE + o, &lo, 2);
					}
				}
			}
}

cudaError_t launch_score_gemm_tma(const float* X, uint32_t Wp, const ScoreTmaParams& p, cudaStream_t s) {
	if (!p.M || !p.P) return cudaSuccess;
	static bool attr_done = false;
	if (!attr_done) {
		cudaError_t e = cudaFuncSetAttribute(score_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
		if (e != cudaSuccess) return e;
		attr_done = true;
	}
	CUtensorMap tm;
	if (!window_map(&tm, X, p.M, p.D, Wp, p.K, BM, true)) return cudaErrorInvalidValue;
	dim3 grid((p.M + BM - 1) / BM, p.ntile, p.D);
	score_gemm_tma_kernel<<<grid, SC_THR, SMEM_BYTES, s>>>(tm, p);
	return cudaGetLastError();
}

cudaError_t launch_state_grad_tma(const float* X, uint32_t Wp, uint32_t K, const StateGradTmaParams& p, cudaStream_t s) {
	if (!p.N || !p.P || !p.J) return cudaSuccess;
	if (p.k_slab % KC) return cudaErrorInvalidValue;
	static bool attr_done = false;
	if (!attr_done) {
		cudaError_t e = cudaFuncSetAttribute(state_grad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
		if (e != cudaSuccess) return e;
		attr_done = true;
	}
	CUtensorMap tm;
	if (!window_map(&tm, X, p.N, p.D, Wp, K, KC, false)) return cudaErrorInvalidValue;
	dim3 grid((p.J + BM - 1) / BM, p.D * p.ntile, (p.N + p.k_slab - 1) / p.k_slab);
	state_grad_tma_kernel<<<grid, SG_THR, SMEM_BYTES, s>>>(tm, p);
	return cudaGetLastError();
}

}  // namespace crfgpu
