// Device kernels of libcrfgpu (sm_100a).  See crf_kernels.cuh for the parameter blocks and DESIGN.md
// for the data layout and the roofline that bounds each kernel.  Citations: ASR-CRaFT tree.
#include "crf_kernels.cuh"

#include <cfloat>
#include <math_constants.h>

namespace crfgpu {

// =================================================================================================
// helpers
// =================================================================================================
__device__ __forceinline__ int float_key(float f) {
	// order-preserving map float -> int so that atomicMax(int) implements a float max
	int i = __float_as_int(f);
	return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) {
	return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}
constexpr int KEY_NEG_INF = (int)0x807fffff;  // float_key(-inf)

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

__global__ void fill_f32_kernel(float* p, uint64_t n, float v) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill_f64_kernel(double* p, uint64_t n, double v) {
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
void launch_fill_f32(float* p, uint64_t n, float v, cudaStream_t s) {
	if (!n) return;
	uint64_t b = (n + 255) / 256; if (b > 148 * 16) b = 148 * 16;
	fill_f32_kernel<<<(unsigned)b, 256, 0, s>>>(p, n, v);
}
void launch_fill_f64(double* p, uint64_t n, double v, cudaStream_t s) {
	if (!n) return;
	uint64_t b = (n + 255) / 256; if (b > 148 * 16) b = 148 * 16;
	fill_f64_kernel<<<(unsigned)b, 256, 0, s>>>(p, n, v);
}

// =================================================================================================
// K7: segment windows.  One CTA per frame; thread f walks the window back from the last frame so the
// running sum / max / min are accumulated in the reference's order
// (CRF_InFtrStream_SeqMultiWindow.cpp: sample_ftrs :556-590, avg_ftrs :601-626, max_ftrs :637-666,
//  min_ftrs :677-706, dur_ftrs :790-812; first_frame_ftrs for the non-segment mode).
// =================================================================================================
// One CTA per frame.  The base rows the frame's windows reach back to are fetched with ONE batch of independent loads into shared
// memory, the D windows are built there (running sum / max / min in the reference's order) and leave the SM as 16-byte vectors of
// whole rows when the row stride allows it (Wp % 4 == 0: every window row is 16-byte aligned), otherwise element by element.
__global__ void __launch_bounds__(128) expand_windows_kernel(ExpandParams p) {
	// blockIdx.y selects a group of p.dpart durations (the whole window set when dpart == D); small groups keep the CTA's shared memory
	// low enough to run beside a resident lattice CTA (crfgpu_prefetch_batch)
	extern __shared__ __align__(16) float xw_raw[];         // [dpart][Wp] windows, then [D][F] base rows (row j = frame n - j), then [D*5] offsets
	const uint32_t DP = p.dpart, dlo = blockIdx.y * DP + 1, dhi = min(dlo + DP - 1, p.D);   // durations [dlo, dhi] of this CTA
	// three disjoint regions: telling the compiler so lets it hoist the loads of a duration above the stores of the previous one
	float* __restrict__ xw = xw_raw;
	const float* __restrict__ rows = xw_raw + (size_t)DP * p.Wp;
	const uint32_t* __restrict__ stp = reinterpret_cast<const uint32_t*>(xw_raw + (size_t)DP * p.Wp + (size_t)p.D * p.F);
	float* rows_w = xw_raw + (size_t)DP * p.Wp;
	uint32_t* stp_w = reinterpret_cast<uint32_t*>(rows_w + (size_t)p.D * p.F);
	const uint32_t n = p.n0 + blockIdx.x;
	if (n >= p.N) return;
	const uint32_t t = p.frame_t[n];
	const uint32_t dtop = min(t + 1, p.D), dmax = min(dtop, dhi);       // durations that exist for this frame, up to this CTA's last one
	const float* cur = p.base + (uint64_t)n * p.F;
	// rows n, n-1, ... are contiguous in the base stream read backwards: row j, feature f sits at cur[f - j*F].  Eight independent loads per
	// thread are in flight before the first one is stored.
	const float inv_f = 1.0f / (float)p.F;
	for (uint32_t i0 = 0; i0 < dmax * p.F; i0 += 8 * blockDim.x) {
		float v[8];
#pragma unroll
		for (uint32_t k = 0; k < 8; k++) {
			const uint32_t i = i0 + k * blockDim.x + threadIdx.x;
			uint32_t j = __float2uint_rz((float)i * inv_f);          // i / F for i < 2^20, corrected below
			if ((j + 1) * p.F <= i) j++;
			if (j * p.F > i) j--;
			const uint32_t f = i - j * p.F;
			v[k] = i < dmax * p.F ? __ldg(cur - (uint64_t)j * p.F + f) : 0.0f;
		}
#pragma unroll
		for (uint32_t k = 0; k < 8; k++) {
			const uint32_t i = i0 + k * blockDim.x + threadIdx.x;
			if (i < dmax * p.F) rows_w[i] = v[k];
		}
	}
	for (uint32_t i = threadIdx.x; i < p.D * 5; i += blockDim.x) stp_w[i] = p.steps[i];
	// windows that would start before the utterance are never read by the lattice; keep them zero
	for (uint32_t i = (dmax >= dlo ? dmax - dlo + 1 : 0) * p.Wp + threadIdx.x; i < (dhi - dlo + 1) * p.Wp; i += blockDim.x) xw[i] = 0.0f;
	__syncthreads();
	if (p.seg_ftrs) {
		for (uint32_t f = threadIdx.x; f < p.F; f += blockDim.x) {
			float acc = 0.0f, amax = rows[f], amin = rows[f];
			for (uint32_t d = 1; d <= dmax; d++) {
				// the window of duration d covers frames n-d+1 .. n = rows d-1 .. 0; its first frame is row d-1
				const float v = rows[(d - 1) * p.F + f];
				acc += v;
				amax = v > amax ? v : amax;
				amin = v < amin ? v : amin;
				if (d < dlo) continue;                               // the running statistics still cover the shorter windows
				float* o = xw + (d - dlo) * p.Wp;
				float smp[5];
#pragma unroll
				for (int k = 0; k < 5; k++) smp[k] = rows[(d - 1 - stp[(d - 1) * 5 + k]) * p.F + f];
#pragma unroll
				for (int k = 0; k < 5; k++) o[k * p.F + f] = smp[k];
				o[5 * p.F + f] = acc / (float)d;
				o[6 * p.F + f] = amax;
				o[7 * p.F + f] = amin;
			}
		}
		// one-hot duration + the pad behind each window
		const uint32_t tail = p.Wp - 8 * p.F;
		for (uint32_t i = threadIdx.x; i < (dmax >= dlo ? dmax - dlo + 1 : 0) * tail; i += blockDim.x) {
			const uint32_t dl = i / tail, j = i - dl * tail;
			xw[dl * p.Wp + 8 * p.F + j] = (j == dl + dlo - 1) ? 1.0f : 0.0f;
		}
	} else {
		for (uint32_t i = threadIdx.x; i < (dmax >= dlo ? dmax - dlo + 1 : 0) * p.Wp; i += blockDim.x) {
			const uint32_t dl = i / p.Wp, f = i - dl * p.Wp;
			xw[i] = f < p.F ? rows[(dl + dlo - 1) * p.F + f] : 0.0f;
		}
	}
	__syncthreads();
	float* out = p.X + ((uint64_t)n * p.D + (dlo - 1)) * p.Wp;
	const uint32_t tot = (dhi - dlo + 1) * p.Wp;
	if (p.Wp % 4 == 0 && (reinterpret_cast<uintptr_t>(p.X) & 15) == 0) {
		float4* o4 = reinterpret_cast<float4*>(out);
		const float4* s4 = reinterpret_cast<const float4*>(xw);
		for (uint32_t i = threadIdx.x; i < tot / 4; i += blockDim.x) __stcs(o4 + i, s4[i]);
	} else {
		for (uint32_t i = threadIdx.x; i < tot; i += blockDim.x) out[i] = xw[i];
	}
}
// =================================================================================================
// "Virtual windows" for the training GEMMs.  Five of the eight feature blocks of a segment window are SAMPLED FRAMES -- plain rows of the
// base stream at an offset that depends only on (duration, sample) -- so the score and state-gradient GEMMs read them as row-shifted TMA
// boxes of a PADDED copy of the base stream and only the three aggregate blocks (average, maximum, minimum over the window's frames,
// CRF_InFtrStream_SeqMultiWindow.cpp:592-706) are materialised: [N][D][Wa] with Wa = roundup(3F + 1, 32) instead of [N][D][8F + D]
// (cfg4: 12.8 instead of 34 kB per frame).  The one-hot duration block (:790-812) is 1 exactly at the window's own duration: it enters the
// scores through a per-(duration, label) bias and the gradient through the constant-1 row that also counts the bias.
// =================================================================================================
// per-frame index tables of a ragged batch from its utterance offsets: frame_t = position inside the utterance, frame_utt = utterance,
// frame_len = its length (the host used to build and upload 12 bytes per frame before the first feature byte could leave)
__global__ void __launch_bounds__(256) frame_tables_kernel(const uint32_t* __restrict__ off, uint32_t n_utt, uint32_t N, uint32_t* ft, uint32_t* fu, uint32_t* fl) {
	const uint32_t n = blockIdx.x * 256 + threadIdx.x;
	if (n >= N) return;
	uint32_t lo = 0, hi = n_utt;             // the utterance u with off[u] <= n < off[u+1]
	while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(off + mid) <= n) lo = mid; else hi = mid; }
	const uint32_t o = __ldg(off + lo);
	if (ft) ft[n] = n - o;
	if (fu) fu[n] = lo;
	if (fl) fl[n] = __ldg(off + lo + 1) - o;
}

__global__ void __launch_bounds__(256) pad_base_kernel(const float* base, float* base2, uint32_t n0, uint32_t n1, uint32_t F, uint32_t Fp) {
	const uint64_t tot = (uint64_t)(n1 - n0) * Fp;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < tot; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t n = n0 + i / Fp; const uint32_t f = (uint32_t)(i % Fp);
		base2[n * Fp + f] = f < F ? __ldg(base + n * F + f) : 0.0f;
	}
}
// one CTA per frame: running sum / maximum / minimum over the window's frames in the reference's order (longest reach last), exactly the
// arithmetic of expand_windows_kernel for these three blocks
__global__ void __launch_bounds__(128) expand_agg_kernel(const float* base, const uint32_t* frame_t, float* Xa, uint32_t N, uint32_t F, uint32_t D, uint32_t Wa, uint32_t n0) {
	const uint32_t n = n0 + blockIdx.x;
	if (n >= N) return;
	const uint32_t dmax = min(__ldg(frame_t + n) + 1, D);
	float* out = Xa + (uint64_t)n * D * Wa;
	const float* cur = base + (uint64_t)n * F;
	for (uint32_t f = threadIdx.x; f < F; f += blockDim.x) {
		float acc = 0.0f, amax = 0.0f, amin = 0.0f;
		for (uint32_t d = 1; d <= D; d++) {
			float* o = out + (uint64_t)(d - 1) * Wa;
			if (d <= dmax) {
				const float v = __ldg(cur - (uint64_t)(d - 1) * F + f);
				acc += v;
				amax = (d == 1 || v > amax) ? v : amax;
				amin = (d == 1 || v < amin) ? v : amin;
				o[f] = acc / (float)d; o[F + f] = amax; o[2 * F + f] = amin;
			} else { o[f] = 0.0f; o[F + f] = 0.0f; o[2 * F + f] = 0.0f; }
		}
	}
	for (uint32_t i = threadIdx.x; i < D * (Wa - 3 * F); i += blockDim.x) out[(uint64_t)(i / (Wa - 3 * F)) * Wa + 3 * F + i % (Wa - 3 * F)] = 0.0f;
}
void launch_virtual_windows(const float* base, const uint32_t* frame_t, float* base2, float* Xa, uint32_t N, uint32_t F, uint32_t Fp, uint32_t D,
                            uint32_t Wa, uint32_t n0, uint32_t n1, cudaStream_t s) {
	if (n1 <= n0) return;
	pad_base_kernel<<<148 * 4, 256, 0, s>>>(base, base2, n0, n1, F, Fp);
	expand_agg_kernel<<<n1 - n0, 128, 0, s>>>(base, frame_t, Xa, N, F, D, Wa, n0);
}

// =================================================================================================
// General window expansion: context frames around the window, boundary deltas, a second stream joined behind the first
// (CRF_InFtrStream_SeqMultiWindow::read_ftrs, CRF/src/io/CRF_InFtrStream_SeqMultiWindow.cpp:325-471: first_frame_left_ctx_ftrs :897-935 |
// sample / avg / max / min / dur or first_frame_ftrs :1017-1048 | last_frame_right_ctx_ftrs :977-1015 (segment features) or
// first_frame_right_ctx_ftrs :937-975, or boundary_delta_ftrs :1050-1110 alone; QN_InFtrStream_JoinFtrs puts stream 2's vector behind
// stream 1's).  One CTA per frame, one thread per output element: a gather with the reference's running sum / max / min order.
// Labelled frame t of utterance u is row  n + u (lc + rc) + lc  of a stream with lc + rc context frames per utterance.
// =================================================================================================
__global__ void __launch_bounds__(256) expand_joined_kernel(ExpandJoinedParams p) {
	// One CTA per frame, one thread per window COLUMN, the durations as the inner loop: which stream / block / feature a column is
	// is worked out once (it is the same for every duration), and the average / maximum / minimum blocks grow by one frame per
	// duration in the reference's own order (from the window's last frame back to its first) instead of being re-summed per duration.
	const uint32_t n = blockIdx.x;
	if (n >= p.N) return;
	const uint32_t t = __ldg(p.frame_t + n), u = __ldg(p.frame_utt + n);
	const uint32_t dmax = min(t + 1, p.D);
	float* out = p.X + (uint64_t)n * p.D * p.Wp;
	for (uint32_t j0 = threadIdx.x; j0 < p.Wp; j0 += blockDim.x) {
		uint32_t j = j0; int q = -1;
		for (uint32_t qq = 0; qq < p.n_parts; qq++) { if (j < p.part[qq].width) { q = (int)qq; break; } j -= p.part[qq].width; }
		if (q < 0) {                                       // padding behind the joined vector
			for (uint32_t d0 = 0; d0 < p.D; d0++) out[(uint64_t)d0 * p.Wp + j0] = 0.0f;
			continue;
		}
		const JoinedPart& a = p.part[q];
		const uint32_t F = a.F, k = j / F, f = j - k * F;
		const float* x = a.x + f;
		const int64_t last = (int64_t)n + (int64_t)u * (a.lc + a.rc) + a.lc;      // row of the window's last frame
		// columns no kernel reads in windows longer than one frame (the transition features of a model whose transition scores come
		// from the duration-1 window): zeros instead of the gather
		const bool d1_only = p.keep_hi != 0 && (j0 < p.keep_lo || j0 >= p.keep_hi);
		const bool delta = a.bdelta && !(p.D > 1 && a.seg);
		const bool body = !delta && k >= a.lc && !(p.D == 1 || !a.seg);
		const uint32_t b = body ? k - a.lc : 0;                                    // block inside the segment-feature body
		const uint32_t jj = j - (a.lc + 8) * F;                                     // (b >= 8) one-hot duration, then the right context of the last frame
		float acc = 0.0f, amax = 0.0f, amin = 0.0f;
		for (uint32_t d0 = 0; d0 < p.D; d0++) {
			float v = 0.0f;
			if (d0 < dmax && !(d1_only && d0 >= 1)) {
				const uint32_t d = d0 + 1;
				const int64_t first = last - d0;                                        // row of the window's first frame
				if (delta) {
					const float l = __ldg(x + (first - 1 - k) * F), r = __ldg(x + (first + k) * F);
					v = l >= r ? l - r : r - l;
				} else if (k < a.lc) v = __ldg(x + (first - a.lc + k) * F);
				else if (!body) v = __ldg(x + (first + (k - a.lc)) * F);                 // the first frame, then its right context
				else if (b < 5) v = __ldg(x + (first + __ldg(p.steps + d0 * 5 + b)) * F);
				else if (b < 8) {
					const float w = __ldg(x + first * F);                                  // the frame this duration adds
					acc += w; amax = (d0 == 0 || w > amax) ? w : amax; amin = (d0 == 0 || w < amin) ? w : amin;
					v = b == 5 ? acc / (float)d : (b == 6 ? amax : amin);
				} else if (jj < p.D) v = jj == d0 ? 1.0f : 0.0f;
				else { const uint32_t kk = (jj - p.D) / F, ff = (jj - p.D) - kk * F; v = __ldg(a.x + ff + (last + 1 + kk) * F); }
			}
			out[(uint64_t)d0 * p.Wp + j0] = v;
		}
	}
}
void launch_expand_joined(const ExpandJoinedParams& p, cudaStream_t s) {
	if (!p.N) return;
	expand_joined_kernel<<<p.N, 256, 0, s>>>(p);
}

void launch_expand_windows(const ExpandParams& p, uint32_t n1, cudaStream_t s) {
	if (n1 <= p.n0) return;
	ExpandParams q = p;
	if (q.dpart == 0 || q.dpart > q.D) q.dpart = q.D;
	const size_t smem = sizeof(float) * ((size_t)q.dpart * q.Wp + (size_t)q.D * (q.F + 5));
	cudaFuncSetAttribute(expand_windows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device, so no process-wide cache
	dim3 grid(n1 - q.n0, (q.D + q.dpart - 1) / q.dpart);
	expand_windows_kernel<<<grid, 128, smem, s>>>(q);
}

// =================================================================================================
// GEMM-1 (state scores): C[n][j] = sum_k A[n][k]*B[j][k] + bias[j]   (fp32 FFMA, 64x64x16 tiles)
// replaces L virtual calls of CRF_StdFeatureMap::computeStateArrayValue per frame
// (CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:65-81)
// =================================================================================================
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;
__global__ void __launch_bounds__(256) score_gemm_kernel(ScoreGemmParams p) {
	__shared__ __align__(16) float As[SG_BK][SG_BM + 4];
	__shared__ __align__(16) float Bs[SG_BK][SG_BN + 4];
	const uint32_t m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * SG_BN;
	const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
	float acc[4][4];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
	for (uint32_t k0 = 0; k0 < p.K; k0 += SG_BK) {
#pragma unroll
		for (int i = 0; i < 4; i++) {
			const int idx = threadIdx.x + i * 256;
			const int r = idx / SG_BK, kk = idx % SG_BK;
			const uint32_t gm = m0 + r, gn = n0 + r, gk = k0 + kk;
			As[kk][r] = (gm < p.M && gk < p.K) ? p.A[(uint64_t)gm * p.lda + gk] : 0.0f;
			Bs[kk][r] = (gn < p.Ncols && gk < p.K) ? p.B[(uint64_t)gn * p.ldb + gk] : 0.0f;
		}
		__syncthreads();
#pragma unroll
		for (int kk = 0; kk < SG_BK; kk++) {
			const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
			const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
			const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
		}
		__syncthreads();
	}
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const uint32_t gm = m0 + ty * 4 + i;
		if (gm >= p.M) continue;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const uint32_t gn = n0 + tx * 4 + j;
			if (gn < p.Ncols) p.C[(uint64_t)gm * p.ldc + gn] = acc[i][j] + (p.bias ? p.bias[gn] : 0.0f);
		}
	}
}
void launch_score_gemm(const ScoreGemmParams& p, cudaStream_t s) {
	if (!p.M || !p.Ncols) return;
	dim3 grid((p.M + SG_BM - 1) / SG_BM, (p.Ncols + SG_BN - 1) / SG_BN);
	score_gemm_kernel<<<grid, 256, 0, s>>>(p);
}

// =================================================================================================
// Lattice recursions (K2/K3/K4), probability domain with per-frame log scales.
//
// Reference semantics (log domain, fp64): CRF_StdSegStateNode::computeAlpha / computeBeta / computeExpF /
// computeAlphaSum (CRF/src/nodes/CRF_StdSegStateNode.cpp:135-186, 219-308, 343-438, 447-462); with
// max_dur == 1 these are exactly CRF_StdStateNode's (CRF/src/nodes/CRF_StdStateNode.cpp:81-299), and the
// N-state topology (CRF_StdNStateNode.cpp:110-365) is the same recursion with E == 0 on illegal pairs.
//
//   alpha_t[(d,y)] = S_t[(d,y)] + log sum_{q<avail(t-d)} exp(alpha_{t-d}[q] + M[q,(d,y)])     d <= t
//                  = S_t[(d,y)]                                                               d == t+1
// Here alpha_t[c] = m_t + log A_t[c] with max_c A_t[c] == 1, and G_t[c] = sum_q A_t[q]*E[q][c] is pushed
// once per frame and consumed at t+d by block d, so each frame costs one [avail x L] mat-vec.
// =================================================================================================
__host__ __device__ __forceinline__ size_t dp_vec_bytes(uint32_t L, int U) {
	return ((size_t)L * U * sizeof(float) + 15) / 16 * 16;
}
// the U per-slot values of row q of a [rows][U] shared-memory vector, as one vector load
template <int U> __device__ __forceinline__ void load_slots(const float* base, uint32_t q, float (&out)[U]);
template <> __device__ __forceinline__ void load_slots<1>(const float* base, uint32_t q, float (&out)[1]) { out[0] = base[q]; }
template <> __device__ __forceinline__ void load_slots<2>(const float* base, uint32_t q, float (&out)[2]) {
	const float2 v = reinterpret_cast<const float2*>(base)[q]; out[0] = v.x; out[1] = v.y;
}
template <> __device__ __forceinline__ void load_slots<4>(const float* base, uint32_t q, float (&out)[4]) {
	const float4 v = reinterpret_cast<const float4*>(base)[q]; out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <> __device__ __forceinline__ void load_slots<8>(const float* base, uint32_t q, float (&out)[8]) {
	const float4 v = reinterpret_cast<const float4*>(base)[2 * q], w = reinterpret_cast<const float4*>(base)[2 * q + 1];
	out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; out[4] = w.x; out[5] = w.y; out[6] = w.z; out[7] = w.w;
}

template <int U>
__global__ void __launch_bounds__(1024) forward_kernel(DpParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float* a_s = reinterpret_cast<float*>(smem_raw);                              // [L][U], 16B aligned
	double* m_ring = reinterpret_cast<double*>(smem_raw + dp_vec_bytes(p.L, U));  // [U][D]
	double* mu_s = m_ring + (size_t)U * p.D;                                      // [U]
	double* zsum_s = mu_s + U;                                                    // [U]
	int* key_s = reinterpret_cast<int*>(zsum_s + U);                              // [2][U]
	float* delta_s = reinterpret_cast<float*>(key_s + 2 * U);                     // [U][D+1] (index d, d==0 unused)
	__shared__ uint32_t s_off[U], s_len[U];

	const uint32_t c = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
	const uint32_t L = p.L, Lp = p.Lp, P = p.P, D = p.D;
	if (threadIdx.x < U) {
		const uint32_t utt = p.grp_utt[blockIdx.x * U + threadIdx.x];
		s_off[threadIdx.x] = utt == LAB_BAD ? 0 : p.off[utt];
		s_len[threadIdx.x] = utt == LAB_BAD ? 0 : p.off[utt + 1] - p.off[utt];
		key_s[threadIdx.x] = KEY_NEG_INF; key_s[U + threadIdx.x] = KEY_NEG_INF;
		zsum_s[threadIdx.x] = 0.0;
	}
	__syncthreads();
	uint32_t Tmax = 0;
#pragma unroll
	for (int u = 0; u < U; u++) Tmax = max(Tmax, s_len[u]);
	const uint32_t my_d = (c < L) ? c / P + 1 : 0xffffu;  // duration block of my column

	for (uint32_t t = 0; t < Tmax; t++) {
		const uint32_t avail = P * min(t + 1, D);
		// [A] per-slot scale bookkeeping: warp u serves slot u, lane d-1 serves duration d
		for (uint32_t u = warp; u < U; u += n_warps) {
			if (t < s_len[u]) {
				double val = -DBL_MAX;
				const uint32_t d = lane + 1;
				if (d <= D) {
					if (d <= t) val = m_ring[u * D + (t - d) % D] + p.Mmax;
					else if (d == t + 1) val = 0.0;   // segment starts the utterance: alpha = S
				}
				const double mu = warp_max_d(val);
				if (d <= D) delta_s[u * (D + 1) + d] = (val == -DBL_MAX) ? -INFINITY : (float)(val - mu);
				if (lane == 0) mu_s[u] = mu;
			}
		}
		__syncthreads();
		// [B] log-domain candidate, block max per slot
		float lr[U];
		if (threadIdx.x < U) key_s[((t + 1) & 1) * U + threadIdx.x] = KEY_NEG_INF;   // reset next step's key
#pragma unroll
		for (int u = 0; u < U; u++) {
			lr[u] = -INFINITY;
			if (t < s_len[u] && c < avail) {
				const uint64_t n = (uint64_t)s_off[u] + t;
				float lg = 0.0f;
				if (my_d <= t) lg = logf(p.G[(n - my_d) * Lp + c]);
				lr[u] = p.S[n * Lp + c] + lg + delta_s[u * (D + 1) + my_d];
			}
			const float wm = warp_max(lr[u]);
			if (lane == 0 && wm > -INFINITY) atomicMax(&key_s[(t & 1) * U + u], float_key(wm));
		}
		__syncthreads();
		// [C] normalise, publish A_t
		float part[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			part[u] = 0.0f;
			if (t < s_len[u]) {
				const float mx = key_float(key_s[(t & 1) * U + u]);
				const float a = (c < avail) ? expf(lr[u] - mx) : 0.0f;
				if (c < L) {
					a_s[c * U + u] = a;
					p.A[((uint64_t)s_off[u] + t) * Lp + c] = a;
				}
				part[u] = a;
				if (c == 0) {
					const double mt = mu_s[u] + (double)mx;
					m_ring[u * D + t % D] = mt;
					p.m[(uint64_t)s_off[u] + t] = mt;
				}
			} else if (c < L) a_s[c * U + u] = 0.0f;
		}
		// logZ = m_{T-1} + log sum_{q<avail(T-1)} A_{T-1}[q]   (computeAlphaSum, :447-462)
		bool any_last = false;
#pragma unroll
		for (int u = 0; u < U; u++) any_last |= (t + 1 == s_len[u]);
		if (any_last) {
#pragma unroll
			for (int u = 0; u < U; u++) {
				if (t + 1 == s_len[u]) {
					const double ws = warp_sum_d((double)part[u]);
					if (lane == 0) atomicAdd(&zsum_s[u], ws);
				}
			}
		}
		__syncthreads();
		if (any_last && threadIdx.x < U && t + 1 == s_len[threadIdx.x]) {
			const uint32_t utt = p.grp_utt[blockIdx.x * U + threadIdx.x];
			p.logZ[utt] = m_ring[threadIdx.x * D + t % D] + log(zsum_s[threadIdx.x]);
		}
		// [D] push: G_t[c] = sum_{q<avail} A_t[q] * E[q][c]
		bool any_next = false;
#pragma unroll
		for (int u = 0; u < U; u++) any_next |= (t + 1 < s_len[u]);
		if (any_next && c < L) {
			float acc[U];
#pragma unroll
			for (int u = 0; u < U; u++) acc[u] = 0.0f;
			const float* Ec = p.E + c;
			uint32_t q = 0;
			for (; q + 4 <= avail; q += 4) {
				const float e0 = __ldg(Ec + (uint64_t)(q + 0) * Lp), e1 = __ldg(Ec + (uint64_t)(q + 1) * Lp);
				const float e2 = __ldg(Ec + (uint64_t)(q + 2) * Lp), e3 = __ldg(Ec + (uint64_t)(q + 3) * Lp);
				float a0[U], a1[U], a2[U], a3[U];
				load_slots<U>(a_s, q + 0, a0); load_slots<U>(a_s, q + 1, a1);
				load_slots<U>(a_s, q + 2, a2); load_slots<U>(a_s, q + 3, a3);
#pragma unroll
				for (int u = 0; u < U; u++) {
					acc[u] = fmaf(a0[u], e0, acc[u]);
					acc[u] = fmaf(a1[u], e1, acc[u]);
					acc[u] = fmaf(a2[u], e2, acc[u]);
					acc[u] = fmaf(a3[u], e3, acc[u]);
				}
			}
			for (; q < avail; q++) {
				const float e0 = __ldg(Ec + (uint64_t)q * Lp);
				float a0[U];
				load_slots<U>(a_s, q, a0);
#pragma unroll
				for (int u = 0; u < U; u++) acc[u] = fmaf(a0[u], e0, acc[u]);
			}
#pragma unroll
			for (int u = 0; u < U; u++)
				if (t + 1 < s_len[u]) p.G[((uint64_t)s_off[u] + t) * Lp + c] = acc[u];
		}
		// G_t[c] is read back only by this same thread (column c), A/a_s hazards are covered by the
		// barriers of the next step before a_s is rewritten in [C].
	}
}

static size_t dp_smem_bytes(const DpParams& p, int U) {
	size_t b = dp_vec_bytes(p.L, U);                                  // a_s / v_s
	b += sizeof(double) * ((size_t)U * p.D + 2 * U);                  // ring + two per-slot doubles
	b += sizeof(int) * 2 * U;                                         // keys
	b += sizeof(float) * 2 * (size_t)U * (p.D + 1);                   // delta, rsc
	return b;
}

template <int U>
static void launch_forward_u(const DpParams& p, cudaStream_t s) {
	const unsigned threads = (p.L + 31) / 32 * 32;
	const size_t smem = dp_smem_bytes(p, U);
	cudaFuncSetAttribute(forward_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	forward_kernel<U><<<p.n_groups, threads, smem, s>>>(p);
}
void launch_forward(const DpParams& p, int U, cudaStream_t s) {
	if (!p.n_groups) return;
	switch (U) {
		case 1: launch_forward_u<1>(p, s); break;
		case 2: launch_forward_u<2>(p, s); break;
		case 4: launch_forward_u<4>(p, s); break;
		default: launch_forward_u<8>(p, s); break;
	}
}

// -------------------------------------------------------------------------------------------------
// backward + posteriors.
//   beta_t[q] = log sum_{d<=numNext} sum_y exp(M[q,(d,y)] + S_{t+d}[(d,y)] + beta_{t+d}[(d,y)])   (:219-308)
// With w~_{t'}[c] = exp(S_{t'}[c] + beta_{t'}[c] - kappa_{t'}) (max 1) kept in G, frame t gathers
// v[(d,y)] = w~_{t+d}[(d,y)] * exp(kappa_{t+d} - kappa*) and does one [L x numNext*P] mat-vec with E^T:
// u[q] = sum_c E[q][c] v[c], beta_t[q] = kappa* + Mmax + log u[q] =: bbase_t + log u[q].
//   gamma_t[c] = exp(alpha+beta-logZ) = A_t[c]*u[c]*exp(m_t + bbase_t - logZ)                       (:369-372)
//   xi_t(q,c)  = A_{t-d}[q] * E[q][c] * R_t[c],  R_t[c] = exp(S_t[c]+beta_t[c]+m_{t-d}+Mmax-logZ)  (:389-397)
// The kernel emits Dm = onehot(reference label) - gamma and R; the sums over frames are the two
// reduce-GEMMs (state weights: Dm^T X, transition bias: A^T R).
// -------------------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(1024) backward_kernel(DpParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float* v_s = reinterpret_cast<float*>(smem_raw);                              // [L][U], 16B aligned
	double* k_ring = reinterpret_cast<double*>(smem_raw + dp_vec_bytes(p.L, U));  // [U][D]
	double* base_s = k_ring + (size_t)U * p.D;                                    // [U]
	double* sg_s = base_s + U;                                                    // [U] log gamma scale
	int* key_s = reinterpret_cast<int*>(sg_s + U);                                // [2][U]
	float* delta_s = reinterpret_cast<float*>(key_s + 2 * U);                     // [U][D+1]
	float* rsc_s = delta_s + (size_t)U * (p.D + 1);                               // [U][D+1]
	__shared__ uint32_t s_off[U], s_len[U];
	__shared__ double s_logZ[U];

	const uint32_t c = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
	const uint32_t L = p.L, Lp = p.Lp, P = p.P, D = p.D;
	if (threadIdx.x < U) {
		const uint32_t utt = p.grp_utt[blockIdx.x * U + threadIdx.x];
		s_off[threadIdx.x] = utt == LAB_BAD ? 0 : p.off[utt];
		s_len[threadIdx.x] = utt == LAB_BAD ? 0 : p.off[utt + 1] - p.off[utt];
		s_logZ[threadIdx.x] = utt == LAB_BAD ? 0.0 : p.logZ[utt];
		key_s[threadIdx.x] = KEY_NEG_INF; key_s[U + threadIdx.x] = KEY_NEG_INF;
	}
	__syncthreads();
	uint32_t Tmax = 0;
#pragma unroll
	for (int u = 0; u < U; u++) Tmax = max(Tmax, s_len[u]);
	const uint32_t my_d = (c < L) ? c / P + 1 : 0xffffu;

	for (uint32_t t = Tmax; t-- > 0;) {
		const uint32_t avail = P * min(t + 1, D);
		uint32_t nn_max = 0;   // largest numNext among the slots active at this step
		// [A] scales: warp u serves slot u, lane d-1 serves duration d
		for (uint32_t u = warp; u < U; u += n_warps) {
			if (t < s_len[u]) {
				const uint32_t numNext = min(s_len[u] - 1 - t, D);
				const uint64_t n = (uint64_t)s_off[u] + t;
				const uint32_t d = lane + 1;
				double kv = -DBL_MAX;
				if (d <= numNext) kv = k_ring[u * D + (t + d) % D];
				const double kstar = warp_max_d(kv);
				const double base = numNext ? kstar + p.Mmax : 0.0;   // tail: beta = 0 (setTailBeta)
				if (d <= D) {
					delta_s[u * (D + 1) + d] = (d <= numNext) ? (float)(kv - kstar) : -INFINITY;
					// xi needs alpha of frame t-d: scale m_{t-d}
					rsc_s[u * (D + 1) + d] = (d <= t) ? (float)(base + p.m[n - d] + p.Mmax - s_logZ[u]) : -INFINITY;
				}
				if (lane == 0) {
					base_s[u] = base;
					sg_s[u] = p.m[n] + base - s_logZ[u];
					p.bbase[n] = base;
				}
			}
		}
#pragma unroll
		for (int u = 0; u < U; u++)
			if (t < s_len[u]) nn_max = max(nn_max, min(s_len[u] - 1 - t, D));
		__syncthreads();
		// [B] gather v from the next numNext frames
		if (threadIdx.x < U) key_s[((t + 1) & 1) * U + threadIdx.x] = KEY_NEG_INF;
		if (c < L) {
#pragma unroll
			for (int u = 0; u < U; u++) {
				float v = 0.0f;
				if (t < s_len[u]) {
					const uint32_t numNext = min(s_len[u] - 1 - t, D);
					if (my_d <= numNext)
						v = p.G[((uint64_t)s_off[u] + t + my_d) * Lp + c] * expf(delta_s[u * (D + 1) + my_d]);
				}
				v_s[c * U + u] = v;
			}
		}
		__syncthreads();
		// [C] u[q] = sum_c E[q][c] v[c]; posteriors
		float uq[U];
#pragma unroll
		for (int u = 0; u < U; u++) uq[u] = 0.0f;
		if (c < avail && nn_max) {
			const float* Eq = p.ET + c;
			const uint32_t ncol = nn_max * P;
			uint32_t k = 0;
			for (; k + 4 <= ncol; k += 4) {
				const float e0 = __ldg(Eq + (uint64_t)(k + 0) * Lp), e1 = __ldg(Eq + (uint64_t)(k + 1) * Lp);
				const float e2 = __ldg(Eq + (uint64_t)(k + 2) * Lp), e3 = __ldg(Eq + (uint64_t)(k + 3) * Lp);
				float v0[U], v1[U], v2[U], v3[U];
				load_slots<U>(v_s, k + 0, v0); load_slots<U>(v_s, k + 1, v1);
				load_slots<U>(v_s, k + 2, v2); load_slots<U>(v_s, k + 3, v3);
#pragma unroll
				for (int u = 0; u < U; u++) {
					uq[u] = fmaf(v0[u], e0, uq[u]);
					uq[u] = fmaf(v1[u], e1, uq[u]);
					uq[u] = fmaf(v2[u], e2, uq[u]);
					uq[u] = fmaf(v3[u], e3, uq[u]);
				}
			}
			for (; k < ncol; k++) {
				const float e0 = __ldg(Eq + (uint64_t)k * Lp);
				float v0[U];
				load_slots<U>(v_s, k, v0);
#pragma unroll
				for (int u = 0; u < U; u++) uq[u] = fmaf(v0[u], e0, uq[u]);
			}
		}
		float lw[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			lw[u] = -INFINITY;
			float gamma = 0.0f;
			if (t < s_len[u] && c < L) {
				const uint64_t n = (uint64_t)s_off[u] + t;
				const bool tail = (t + 1 == s_len[u]);
				if (c < avail) {
					const float uu = tail ? 1.0f : uq[u];
					const float lu = logf(uu);
					lw[u] = p.S[n * Lp + c] + lu;
					gamma = p.A[n * Lp + c] * expf(lu + (float)sg_s[u]);
					p.Dm[n * Lp + c] = ((p.node_lab[n] == c) ? 1.0f : 0.0f) - gamma;
					p.R[n * Lp + c] = (my_d <= t) ? expf(lw[u] + rsc_s[u * (D + 1) + my_d]) : 0.0f;
					if (p.Uvec) p.Uvec[n * Lp + c] = uu;
				} else {
					p.Dm[n * Lp + c] = 0.0f;
					p.R[n * Lp + c] = 0.0f;
					if (p.Uvec) p.Uvec[n * Lp + c] = 0.0f;
				}
			}
			const float wm = warp_max(lw[u]);
			if (lane == 0 && wm > -INFINITY) atomicMax(&key_s[(t & 1) * U + u], float_key(wm));
			if (p.mass) {
				const float gs = warp_sum(gamma);
				if (lane == 0 && t < s_len[u] && gs != 0.0f) atomicAdd(&p.mass[(uint64_t)s_off[u] + t], (double)gs);
			}
		}
		__syncthreads();
		// [D] publish w~_t and its scale
#pragma unroll
		for (int u = 0; u < U; u++) {
			if (t < s_len[u]) {
				const float mx = key_float(key_s[(t & 1) * U + u]);
				const uint64_t n = (uint64_t)s_off[u] + t;
				if (c < L) p.G[n * Lp + c] = (c < avail) ? expf(lw[u] - mx) : 0.0f;
				if (c == 0) {
					const double kap = base_s[u] + (double)mx;
					k_ring[u * D + t % D] = kap;
					p.kappa[n] = kap;
				}
			}
		}
		__syncthreads();   // k_ring / G of this frame are read by other threads in the next step
	}
}

template <int U>
static void launch_backward_u(const DpParams& p, cudaStream_t s) {
	const unsigned threads = (p.L + 31) / 32 * 32;
	const size_t smem = dp_smem_bytes(p, U);
	cudaFuncSetAttribute(backward_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	backward_kernel<U><<<p.n_groups, threads, smem, s>>>(p);
}
void launch_backward(const DpParams& p, int U, cudaStream_t s) {
	if (!p.n_groups) return;
	switch (U) {
		case 1: launch_backward_u<1>(p, s); break;
		case 2: launch_backward_u<2>(p, s); break;
		case 4: launch_backward_u<4>(p, s); break;
		default: launch_backward_u<8>(p, s); break;
	}
}

__global__ void dump_alpha_beta_kernel(DpParams p, uint32_t N, const uint32_t* frame_t, const uint32_t* frame_len,
                                       double* alpha, double* beta) {
	const uint64_t total = (uint64_t)N * p.L;
	for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t n = i / p.L; const uint32_t c = (uint32_t)(i % p.L);
		const uint32_t t = frame_t[n];
		const uint32_t avail = p.P * min(t + 1, p.D);
		double a = -DBL_MAX, b = -DBL_MAX;
		if (c < avail) {
			const float av = p.A[n * p.Lp + c];
			a = av > 0.0f ? p.m[n] + log((double)av) : -INFINITY;
			const float uv = p.Uvec[n * p.Lp + c];
			b = uv > 0.0f ? p.bbase[n] + log((double)uv) : -INFINITY;
		} else if (t + 1 == frame_len[n]) b = 0.0;   // setTailBeta writes all labels
		alpha[i] = a; beta[i] = b;
	}
}
void launch_dump_alpha_beta(const DpParams& p, uint32_t N, const uint32_t* frame_t, const uint32_t* frame_len,
                            double* alpha, double* beta, cudaStream_t s) {
	if (N) dump_alpha_beta_kernel<<<148 * 4, 256, 0, s>>>(p, N, frame_t, frame_len, alpha, beta);
}

// =================================================================================================
// Posterior-mass assertion of the nodes' computeExpF: mass_n = sum_c gamma_n(c) = [a reference label sits on frame n] - sum_c Dm[n][c].
// Frame-level nodes throw unless 0.9 <= mass <= 1.1 (CRF_StdStateNode.cpp:252-275, CRF_StdNStateNode.cpp); segmental nodes -- where the
// mass is the probability that a segment ENDS on the frame -- unless -1e-6 <= mass <= 1.000001 (CRF_StdSegStateNode.cpp:417-436); the
// device holds posteriors in fp32 and sums up to 1024 of them, so its segmental band is [-tol, 1 + tol] with tol = 1e-3.  NaN counts as a
// violation.  One warp per frame; violating frames are counted into *n_bad (the 4th tail scalar behind the gradient, so that it is
// all-reduced with it), mass[] is kept for crfgpu_fetch_posterior_mass.
// =================================================================================================
__global__ void __launch_bounds__(256) posterior_mass_kernel(const float* Dm, uint64_t ld, uint32_t L, const uint32_t* node_lab, uint32_t N,
                                                             float lo, float hi, float* mass, double* n_bad) {
	const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
	uint32_t bad = 0;
	for (uint32_t n = w; n < N; n += nw) {
		const float* row = Dm + (uint64_t)n * ld;
		float sum = 0.0f;
		for (uint32_t c = lane; c < L; c += 32) sum += row[c];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
		const float m = ((node_lab[n] != LAB_BAD) ? 1.0f : 0.0f) - sum;
		if (lane == 0) {
			mass[n] = m;
			if (!(m >= lo && m <= hi)) bad++;
		}
	}
	if (lane == 0 && bad) atomicAdd(n_bad, (double)bad);
}
void launch_posterior_mass(const float* Dm, uint64_t ld, uint32_t L, const uint32_t* node_lab, uint32_t N, float lo, float hi, float* mass,
                           double* n_bad, cudaStream_t s) {
	if (!N) return;
	unsigned blocks = (N + 7) / 8; if (blocks > 148 * 8) blocks = 148 * 8;
	posterior_mass_kernel<<<blocks, 256, 0, s>>>(Dm, ld, L, node_lab, N, lo, hi, mass, n_bad);
}

// =================================================================================================
// Reduce-GEMM: out[map(i,j)] += scale * sum_n A[n-shift][i] * B[n][j]   (fp32 tiles, fp64 atomics)
// replaces the per-frame scatter of CRF_StdFeatureMap::computeStateExpF / computeTransExpF
// (CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:130-223) and `grad -= ExpF` (CRF_NewGradBuilder.cpp:374-376).
// =================================================================================================
constexpr int RG_BI = 64, RG_BJ = 64, RG_BK = 16;
__global__ void __launch_bounds__(256) reduce_gemm_kernel(ReduceGemmParams p) {
	__shared__ __align__(16) float As[RG_BK][RG_BI];
	__shared__ __align__(16) float Bs[RG_BK][RG_BJ];
	const uint32_t i0 = blockIdx.x * RG_BI, j0 = blockIdx.y * RG_BJ;
	const uint32_t ns = p.n0 + blockIdx.z * p.k_slab;
	const uint32_t ne = min(ns + p.k_slab, p.n1);
	const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
	// fp32 products are chained over one 16-row tile only, then folded into fp64 accumulators: with
	// ~10^5 frames per launch a pure fp32 chain drifts by ~1e-7 per add (systematically when many frames
	// carry the same posterior), which would eat the 1e-4 gradient tolerance.
	double acc[4][4];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
	for (uint32_t nb = ns; nb < ne; nb += RG_BK) {
#pragma unroll
		for (int r = 0; r < 4; r++) {
			const int idx = threadIdx.x + r * 256;
			const int kk = idx / 64, col = idx % 64;
			const uint32_t n = nb + kk;
			float a = 0.0f, b = 0.0f;
			if (n < ne) {
				if (i0 + col < p.I) a = p.A[(uint64_t)(n - p.a_row_shift) * p.lda + i0 + col];
				if (j0 + col < p.J) b = (j0 + col == p.ones_col) ? 1.0f : p.B[(uint64_t)n * p.ldb + j0 + col];
			}
			As[kk][col] = a; Bs[kk][col] = b;
		}
		__syncthreads();
		float part[4][4];
#pragma unroll
		for (int i = 0; i < 4; i++)
#pragma unroll
			for (int j = 0; j < 4; j++) part[i][j] = 0.0f;
#pragma unroll
		for (int kk = 0; kk < RG_BK; kk++) {
			const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
			const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
			const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
		}
#pragma unroll
		for (int i = 0; i < 4; i++)
#pragma unroll
			for (int j = 0; j < 4; j++) acc[i][j] += (double)part[i][j];
		__syncthreads();
	}
#pragma unroll
	for (int i = 0; i < 4; i++) {
		const uint32_t gi = i0 + ty * 4 + i;
		if (gi >= p.I) continue;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const uint32_t gj = j0 + tx * 4 + j;
			if (gj >= p.J) continue;
			const double v = acc[i][j];
			if (v == 0.0) continue;
			if (p.mode == 0) {
				const double sc = (gj == p.ones_col) ? p.ones_scale : p.scale;
				if (p.row_idx[gi] != 0xffffffffu) atomicAdd(&p.out[(uint64_t)p.row_idx[gi] + gj], sc * v);
			} else {
				const uint32_t idx = p.pair_idx[(uint64_t)gi * p.pair_ld + gj];
				if (idx != 0xffffffffu) atomicAdd(&p.out[idx], p.scale * (double)p.Ew[(uint64_t)gi * p.e_ld + gj] * v);
			}
		}
	}
}
void launch_reduce_gemm(const ReduceGemmParams& p, cudaStream_t s) {
	if (p.n1 <= p.n0 || !p.I || !p.J) return;
	dim3 grid((p.I + RG_BI - 1) / RG_BI, (p.J + RG_BJ - 1) / RG_BJ, (p.n1 - p.n0 + p.k_slab - 1) / p.k_slab);
	reduce_gemm_kernel<<<grid, 256, 0, s>>>(p);
}

// =================================================================================================
// Empirical counts on the reference path and the numerator sum lambda.f (fp64, one warp per labelled node)
// (CRF_StdFeatureMap::computeStateExpF/computeTransExpF `t_clab==clab` branches, :150-171, :205-220)
// The state-feature empirical counts themselves ride in Dm (the onehot term); here only the
// transition-bias counts and the numerator are produced.
// =================================================================================================
__global__ void __launch_bounds__(256) empirical_kernel(EmpiricalParams p) {
	const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t n = warp_global; n < p.N; n += n_warps) {
		const uint32_t lab = p.node_lab[n];
		if (lab == LAB_BAD || lab >= p.L) continue;
		const uint32_t d = lab / p.P;   // duration block (0-based); 0 for frame-level models
		const double* lam = p.lambda + p.sidx[lab];
		double acc = 0.0;
		if (!p.virt) {
			const float* x = p.X + (uint64_t)n * p.ldx + (uint64_t)d * p.W + p.sf0;
			for (uint32_t f = lane; f < p.nSf; f += 32) acc += lam[f] * (double)x[f];
		} else {
			// the window rebuilt from its parts: five sampled base rows, the aggregate array, the one-hot duration feature
			const uint32_t F = p.F;
			for (uint32_t b = 0; b < 5; b++) {
				const float* x = p.base + ((uint64_t)n - d + p.steps[d * 5 + b]) * F;
				for (uint32_t f = lane; f < F; f += 32) acc += lam[b * F + f] * (double)x[f];
			}
			const float* xa = p.X + ((uint64_t)n * p.D + d) * p.W;
			for (uint32_t f = lane; f < 3 * F; f += 32) acc += lam[5 * F + f] * (double)xa[f];
			if (lane == 0) acc += lam[8 * F + d];
		}
		acc = warp_sum_d(acc);
		if (lane == 0) {
			if (p.use_state_bias) acc += lam[p.nSf] * p.state_bias_val;
			const uint32_t pl = p.prev_lab[n];
			if (pl != LAB_BAD && pl < p.L && p.use_trans_bias) {
				const uint32_t ti = (p.tL == p.L) ? p.tidx[(uint64_t)pl * p.L + lab] : p.tidx[(uint64_t)(pl % p.P) * p.tL + lab % p.P];
				if (ti != 0xffffffffu) {
					acc += p.lambda[ti] * p.trans_bias_val;
					atomicAdd(&p.grad[ti], p.trans_bias_val);
				}
			}
			atomicAdd(&p.numer[p.frame_utt[n]], acc);
		}
	}
}
void launch_empirical(const EmpiricalParams& p, cudaStream_t s) {
	if (!p.N) return;
	unsigned blocks = (p.N + 7) / 8; if (blocks > 148 * 8) blocks = 148 * 8;
	empirical_kernel<<<blocks, 256, 0, s>>>(p);
}

// =================================================================================================
// Viterbi scores in the reference's arithmetic: fp64, features in index order, separate multiply and
// add (no FMA), bias last, then negate and narrow to float
// (CRF_StdFeatureMap.cpp:65-81; CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:143)
// =================================================================================================
// One CTA = VS_ROWS (frame, duration) windows x 256 labels: the windows are staged in shared memory and every weight is loaded
// once per CTA and applied to all rows, so the kernel is bound by the fp64 pipe instead of by two loads per multiply-add.
// The accumulation order of every output is unchanged (features in index order, bias last).
constexpr int VS_ROWS = 8;
__global__ void __launch_bounds__(256) vit_scores_kernel(VitScoreParams p) {
	extern __shared__ __align__(16) double vs_x[];           // [nSf][VS_ROWS], widened once per CTA (the float -> double conversion runs
	                                                         // at a quarter of the fp64 multiply-add rate, so it must not sit in the inner loop)
	__shared__ uint32_t row_ok[VS_ROWS];
	const uint64_t n_rows = (uint64_t)p.N * p.D, r0 = (uint64_t)blockIdx.x * VS_ROWS;
	for (uint32_t i = 0; i < VS_ROWS; i++) {
		const uint64_t nd = r0 + i;
		bool ok = nd < n_rows;
		uint64_t n = 0; uint32_t d = 0;
		if (ok) { n = nd / p.D; d = (uint32_t)(nd % p.D); ok = d <= p.frame_t[n]; }
		if (threadIdx.x == 0) row_ok[i] = ok ? 1u : 0u;
		const float* x = p.X + n * p.ldx + (uint64_t)d * p.W + p.sf0;
		for (uint32_t f = threadIdx.x; f < p.nSf; f += blockDim.x) vs_x[f * VS_ROWS + i] = ok ? (double)x[f] : 0.0;
	}
	__syncthreads();
	const uint32_t lab = blockIdx.y * blockDim.x + threadIdx.x;
	if (lab >= p.L) return;
	double acc[VS_ROWS];
#pragma unroll
	for (int i = 0; i < VS_ROWS; i++) acc[i] = 0.0;
	for (uint32_t f = 0; f < p.nSf; f++) {
		const double w = p.Wd[(uint64_t)f * p.L + lab];
#pragma unroll
		for (int i = 0; i < VS_ROWS; i += 2) {
			const double2 xv = *reinterpret_cast<const double2*>(vs_x + f * VS_ROWS + i);     // broadcast, two windows per load
			acc[i] = __dadd_rn(acc[i], __dmul_rn(xv.x, w));
			acc[i + 1] = __dadd_rn(acc[i + 1], __dmul_rn(xv.y, w));
		}
	}
	const double wb = p.use_bias ? __dmul_rn(p.Wd[(uint64_t)p.nSf * p.L + lab], p.bias_val) : 0.0;
#pragma unroll
	for (int i = 0; i < VS_ROWS; i++) {
		const uint64_t nd = r0 + i;
		if (nd >= n_rows) break;
		const double v = p.use_bias ? __dadd_rn(acc[i], wb) : acc[i];
		p.negS[nd * p.L + lab] = row_ok[i] ? (float)(-v) : 0.0f;
	}
}
void launch_vit_scores(const VitScoreParams& p, cudaStream_t s) {
	const uint64_t n_rows = (uint64_t)p.N * p.D;
	if (!n_rows || !p.L) return;
	const size_t smem = sizeof(double) * VS_ROWS * (size_t)(p.nSf ? p.nSf : 1);
	cudaFuncSetAttribute(vit_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	// thread = label: label sets below 256 get a CTA of their own size (cfg3: 183 labels -> 192 threads instead of 256 with 73 idle ones,
	// so more windows are resident per SM to hide the fp64 pipe's latency)
	const unsigned threads = p.L >= 256 ? 256u : (p.L + 31) / 32 * 32;
	dim3 grid((unsigned)((n_rows + VS_ROWS - 1) / VS_ROWS), (p.L + threads - 1) / threads);
	vit_scores_kernel<<<grid, threads, smem, s>>>(p);
}

// =================================================================================================
// Viterbi recursion + traceback, one CTA per utterance, thread = (phone, sub-state) label.
// Array restatement of the token-passing decoder with the free-phone LM and no beam
// (CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:116-166 state update, :246-415 within-phone,
//  :435-543 cross-phone, :545-733 expansion order, :976-1106 pruning/kept list, :2156-2171 final
//  argmin, :2204-2349 traceback; CRF_ViterbiNode::addNonEpsVtbState / choose_nState_Best_Seg .h:193-465).
// All costs are float with the finite sentinel 99999.0; candidate order reproduces the reference's
// first-arrival-wins tie-breaks (see DESIGN.md "Viterbi exactness").
// =================================================================================================
__device__ __forceinline__ uint32_t kept_phone(uint32_t i, uint32_t P, uint32_t g, uint32_t none = 0xffu) {
	// i-th phone of the kept list described by g: identity if g is the "none" sentinel, else increasing order with g moved last
	if (g == none) return i;
	if (i + 1 == P) return g;
	return i < g ? i : i + 1;
}

template <bool HAS_LM, bool BEAM>
__global__ void __launch_bounds__(1024) viterbi_kernel(VitParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	float* Wprev = reinterpret_cast<float*>(smem_raw);   // [L] kept weights of the previous frame
	float* partW = Wprev + p.L;                           // [NS][P] partial cross-phone minima of the frame (all threads share the scan)
	int32_t* partP = reinterpret_cast<int32_t*>(partW + p.L);   // their back pointers, -1: no candidate in that part of the list
	float* We = partW + 2 * p.L;                          // [P] kept weight of every phone's END state + 0.0f (the LM arc weight of the free-phone loop)
	float* crossS = We + p.P;                             // [P][Pt] when it fits, TRANSPOSED: crossS[tq*Pt + pp] = crossT[pp][tq], Pt = P | 1
	__shared__ uint32_t s_g;        // descriptor of the kept-list order of the previous frame
	__shared__ uint8_t s_move[256]; // arrival-order descriptors a[s] of the last D start frames (ring; D <= 255)
	__shared__ int s_best;
	// beam pruning (BEAM, one state per phone): a node keeps the hypotheses whose weight is < min_weight + beam (a float against a double
	// sum, pruning() .cpp:1013-1100) and only kept hypotheses are expanded -- across phones (:573; a pruned phone's entry of We[] is +inf,
	// so it never wins the scan and the list order of the others is untouched) and within the phone -- or end the path.  Every kept
	// hypothesis still reaches every other phone, so every phone has a candidate at every start frame and the kept list is the closed-form
	// order with the pruned phones left out; its HEAD (the phone whose own candidate arrives last at the next start frame) is the first
	// kept phone in that order, found from the warps' ballots.
	__shared__ float s_wmin[32];
	__shared__ uint32_t s_keptmask[32];
	__shared__ uint32_t s_move_w[256];     // BEAM: the arrival-order descriptors are phone ids (any phone can head a pruned list), 0xffffffff = identity
	__shared__ uint32_t s_gw;
	constexpr uint32_t GNONE = BEAM ? 0xffffffffu : 0xffu;
	bool my_kept = true; float my_best = 0.0f;
	const uint32_t u = p.order ? p.order[blockIdx.x] : blockIdx.x;      // longest utterances first
	const uint32_t L = p.L, P = p.P, NS = p.NS, D = p.D;
	const uint32_t off = p.off[u], T = p.off[u + 1] - p.off[u];
	const uint32_t lab = threadIdx.x;
	const uint32_t Pt = P | 1u;      // odd row stride: the rows of 32 consecutive target phones start in 32 different banks
	const bool cross_in_smem = p.negMt == nullptr && ((size_t)P * Pt * sizeof(float) <= 96 * 1024);
	if (cross_in_smem) for (uint32_t i = threadIdx.x; i < P * P; i += blockDim.x) crossS[(i % P) * Pt + i / P] = p.crossT[i];
	// phone-bigram LM (one state per phone): the weight of arc pp -> tq is added to the expanding hypothesis BEFORE the transition score,
	// float by float -- (prev + lm) + trans, expandCrossStateFromPrevNode :629 / crossStateTransUpdate :467 -- so We[] then holds the raw
	// kept weights (the "+ 0.0f" of the free-phone loop IS this add).  The table sits behind the cross table in shared memory when both fit.
	constexpr bool has_lm = HAS_LM;       // compile-time: the free-phone instantiation is the kernel it was before the LM existed
	const bool lm_in_smem = has_lm && NS == 1 && cross_in_smem && (2 * (size_t)P * Pt * sizeof(float) <= 96 * 1024);
	float* lmS = crossS + (size_t)P * Pt;
	if (lm_in_smem) for (uint32_t i = threadIdx.x; i < P * P; i += blockDim.x) lmS[(i / P) * Pt + i % P] = p.lm_bigT[i];
	const float* crossT = p.crossT;      // global table (per-frame tables replace it below); the shared-memory copy is crossS
	const bool per_frame = p.negMt != nullptr;      // transition FEATURES: the tables of the frame a segment starts in (global memory)
	float* candW = p.candW + (uint64_t)u * D * L;
	int32_t* candP = p.candP + (uint64_t)u * D * L;
	const uint32_t q = lab / NS, k = lab % NS;
	// cross-phone scan: thread = (part of the kept list, target phone); constant over the frames
	// (p.Ppad = P rounded up to the warp size when NS * Ppad threads fit the CTA, so that no warp straddles two parts of the list and
	// runs both loops; else P)
	const uint32_t part = threadIdx.x / p.Ppad, tq = threadIdx.x - part * p.Ppad;
	const bool scan_ok = part < NS && tq < P;
	const uint32_t chunk = (P + NS - 1) / NS, i_lo = min(P, part * chunk), i_hi = min(P, i_lo + chunk);
	float my_diag = (lab < L && p.negMt == nullptr) ? p.negDiag[lab] : 0.0f;
	float my_off = (lab < L && k > 0 && p.negMt == nullptr) ? p.negOff[lab] : 0.0f;
	if (threadIdx.x == 0) { s_g = 0xffu; s_gw = 0xffffffffu; }
	__syncthreads();
	const bool timing = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
	unsigned long long tacc[4] = {0, 0, 0, 0};
	long long tlast = timing ? clock64() : 0;
#define VKTICK(i) do { if (timing) { const long long now_ = clock64(); tacc[i] += (unsigned long long)(now_ - tlast); tlast = now_; } } while (0)

	// the state scores of a node do not depend on the recursion: those of node s+1 are fetched while node s is processed, so the
	// per-frame chain holds no global-memory round trip (VPF durations in registers, the rest read in place)
	constexpr uint32_t VPF = 4;
	float nsv[VPF];
#pragma unroll
	for (uint32_t j = 0; j < VPF; j++) nsv[j] = (lab < L && j < D && T > 0) ? p.negS[((uint64_t)off * D + j) * L + lab] : 0.0f;

	// transition FEATURES: the cross-phone table of frame s+1 is copied (transposed like the constant table) into the other half of a
	// double buffer while frame s is processed, its self-loop / advance entries wait in registers: the scan reads shared memory instead
	// of 48 rounds to L2 per frame
	const bool pf_smem = per_frame && (2 * (size_t)P * Pt * sizeof(float) <= 96 * 1024);
	auto fetch_table = [&](uint32_t frame) {      // row `frame` of negMt -> buffer frame & 1 (cp.async, 4 bytes per element)
		const float* tb = p.negMt + (uint64_t)(off + frame) * p.E;
		float* dst = crossS + (size_t)(frame & 1) * P * Pt;
		uint32_t r = threadIdx.x / P, c = threadIdx.x - r * P;      // element i = r * P + c -> dst[c * Pt + r]
		const uint32_t dr = blockDim.x / P, dc = blockDim.x - dr * P;
		for (uint32_t i = threadIdx.x; i < P * P; i += blockDim.x) {
			asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + (size_t)c * Pt + r)), "l"(tb + i) : "memory");
			c += dc; r += dr;
			if (c >= P) { c -= P; r++; }
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	float nx_diag = 0.0f, nx_off = 0.0f;
	if (pf_smem && T > 1) {
		fetch_table(1);
		const float* tb = p.negMt + (uint64_t)(off + 1) * p.E;
		if (lab < L) { nx_diag = tb[P * P + lab]; nx_off = k > 0 ? tb[P * P + L + lab] : 0.0f; }
	}
	const float* crossF = crossS;

	for (uint32_t s = 0; s < T; s++) {
		float nsn[VPF];
#pragma unroll
		for (uint32_t j = 0; j < VPF; j++) nsn[j] = (lab < L && j < D && s + 1 < T) ? p.negS[((uint64_t)(off + s + 1) * D + j) * L + lab] : 0.0f;
		// ---- cross-phone candidates for segments starting at frame s: the scan of the kept list of frame s-1 is shared by ALL threads
		//      of the CTA (with N states per phone only every N-th thread owns a start state): thread = (part of the list, target phone);
		//      strict '<' keeps the first arrival inside a part, and the parts are merged in list order below ----
		if (pf_smem && s > 0) {
			asm volatile("cp.async.wait_group 0;" ::: "memory");
			__syncthreads();                                     // the table of frame s is complete; the other buffer was last read in frame s-1
			crossF = crossS + (size_t)(s & 1) * P * Pt;
			my_diag = nx_diag; my_off = nx_off;
			if (s + 1 < T) {
				fetch_table(s + 1);
				const float* tb = p.negMt + (uint64_t)(off + s + 1) * p.E;
				if (lab < L) { nx_diag = tb[P * P + lab]; nx_off = k > 0 ? tb[P * P + L + lab] : 0.0f; }
			}
		} else if (per_frame && s > 0) {
			const float* tb = p.negMt + (uint64_t)(off + s) * p.E;
			crossT = tb;
			if (lab < L) { my_diag = tb[P * P + lab]; my_off = k > 0 ? tb[P * P + L + lab] : 0.0f; }
		}
		VKTICK(3);   // loop top: prefetch issue (+ list bookkeeping of the previous frame)
		float pw = VIT_INF; int32_t pptr = -1;
		if (s > 0 && scan_ok) {
			const uint32_t g = BEAM ? s_gw : s_g;
			// the table pointer keeps its address space (shared or global) in each instantiation, the per-thread part of the index is
			// hoisted out of the frame loop, and the two list orders have their own loops: per element this is two loads, two adds and
			// one compare-select (generic loads and per-element list arithmetic made the scan 70 % of the frame)
			// The scan is bound by instruction issue (a few warps per scheduler), so it is written for the fewest instructions per list
			// entry: We[] already carries "+ 0.0f", the shared-memory table is transposed (one row per target phone: consecutive entries at
			// consecutive addresses, immediate offsets after unrolling), and the running minimum starts at +inf -- every real candidate is
			// finite, so that equals the reference's "the first candidate is always taken".
			auto scan = [&](const float* ct, const uint32_t stride) {     // entry pp of the target phone's column at ct[pp * stride]
				pw = CUDART_INF_F;
				if (has_lm && NS > 1) {
					// N states per phone: the hypothesis returns to the LM's start state through the epsilon arc of its phone (exit cost, already
					// in We[]) and leaves it on the unigram arc of the target phone: (prev + exit) + unigram, then the transition score
					const float uq = __ldg(p.lm_start + tq);
#pragma unroll 4
					for (uint32_t pp = i_lo; pp < i_hi; pp++) {
						const float cc = (We[pp] + uq) + ct[(size_t)pp * stride];
						if (cc < pw) { pw = cc; pptr = (int32_t)pp; }
					}
				} else if (has_lm) {
					// one state per phone with LM weights: the same list order, two adds per entry
					const float* lr = lm_in_smem ? lmS + (size_t)tq * Pt : p.lm_bigT + (size_t)tq * P;
#pragma unroll 4
					for (uint32_t pp = 0; pp < P; pp++) {
						const float cc = (We[pp] + lr[pp]) + ct[(size_t)pp * stride];
						if (pp != tq && pp != g && cc < pw) { pw = cc; pptr = (int32_t)pp; }
					}
					if (g != GNONE && g != tq) {
						const float cc = (We[g] + lr[g]) + ct[(size_t)g * stride];
						if (cc < pw) { pw = cc; pptr = (int32_t)g; }
					}
				} else if (NS > 1) {
					// N states per phone: the kept list is always in phone order (s_g stays 0xff) and every phone may follow every phone
#pragma unroll 4
					for (uint32_t pp = i_lo; pp < i_hi; pp++) {
						const float cc = We[pp] + ct[(size_t)pp * stride];
						if (cc < pw) { pw = cc; pptr = (int32_t)pp; }
					}
				} else {
					// one state per phone: increasing phone order with phone g (if any) moved to the back; no arc to the same phone (:1332-1346)
#pragma unroll 4
					for (uint32_t pp = 0; pp < P; pp++) {
						const float cc = We[pp] + ct[(size_t)pp * stride];
						if (pp != tq && pp != g && cc < pw) { pw = cc; pptr = (int32_t)pp; }
					}
					if (g != GNONE && g != tq) {
						const float cc = We[g] + ct[(size_t)g * stride];
						if (cc < pw) { pw = cc; pptr = (int32_t)g; }
					}
				}
				if (pptr < 0) pw = VIT_INF;                       // no candidate (P == 1)
				else pptr = pptr * (int32_t)NS + (int32_t)NS - 1;   // the phone's end state
			};
			if (cross_in_smem) scan(crossS + (size_t)tq * Pt, 1u);
			else if (pf_smem) scan(crossF + (size_t)tq * Pt, 1u);
			else scan(crossT + tq, P);
			if (NS > 1) { partW[part * P + tq] = pw; partP[part * P + tq] = pptr; }
		}
		VKTICK(0);   // cross-phone scan
		if (NS > 1) __syncthreads();        // one state per phone: every thread scanned the whole list for its own phone
		// ---- candidates for segments starting at frame s ----
		float cw = VIT_INF; int32_t cp = -1;
		if (lab < L) {
			if (s == 0) {
				if (k == 0) cw = (0.0f + (has_lm ? __ldg(p.lm_start + q) : 0.0f)) + 0.0f;   // lm_start hypothesis 0 + arc weight, trans_wt = 0.0 at node 0 (:444-447)
			} else {
				bool seen = false;
				if (k == 0) {
					if (NS == 1) { if (pptr >= 0) { cw = pw; cp = pptr; seen = true; } }
					else for (uint32_t part = 0; part < NS; part++) {
						const int32_t qp = partP[part * P + q];
						const float qw = partW[part * P + q];
						if (qp >= 0 && (!seen || qw < cw)) { cw = qw; cp = qp; seen = true; }
					}
				} else seen = true;   // the cross update created the slot with 99999.0 / -1 for inner sub-states
				// within-phone: self vs advance from the previous sub-state, self only if strictly smaller (:338)
				if (!BEAM || my_kept) {                              // (a pruned hypothesis is not expanded)
					float w; int32_t ptr;
					const float n1 = Wprev[lab] + my_diag;
					if (k == 0) { w = n1; ptr = (int32_t)lab; }
					else {
						const float n2 = Wprev[lab - 1] + my_off;
						if (n1 < n2) { w = n1; ptr = (int32_t)lab; } else { w = n2; ptr = (int32_t)lab - 1; }
					}
					if (k == 0 && !seen) { cw = w; cp = ptr; }         // no cross arc reached this phone (P == 1, or it is the only kept one)
					else if (w < cw) { cw = w; cp = ptr; }
				}
			}
			if (D > 1) { candW[(uint64_t)(s % D) * L + lab] = cw; candP[(uint64_t)(s % D) * L + lab] = cp; }   // read back d-1 frames later
		}
		VKTICK(1);   // merge + within-phone candidates
		__syncthreads();   // Wprev fully consumed
		// ---- node s: add state values, best duration per (phone, sub-state); longest duration first ----
		if (lab < L) {
			const uint32_t dmax = min(s + 1, D);
			float best = 0.0f; int32_t bptr = -1; uint32_t bdur = 0;
			for (uint32_t d = dmax; d >= 1; d--) {
				const uint32_t s0 = s - d + 1;
				float w = (d == 1) ? cw : candW[(uint64_t)(s0 % D) * L + lab];
				const int32_t ptr = (d == 1) ? cp : candP[(uint64_t)(s0 % D) * L + lab];
				if (w < VIT_INF) {
					float sv;
					if (d <= VPF) {
						sv = nsv[0];
#pragma unroll
						for (uint32_t j = 1; j < VPF; j++) if (d == j + 1) sv = nsv[j];
					} else sv = p.negS[((uint64_t)(off + s) * D + (d - 1)) * L + lab];
					w = w + sv;
				}
				if (d == dmax || w < best) { best = w; bptr = ptr; bdur = d; }
			}
			Wprev[lab] = best;
			if (!BEAM && k == NS - 1) We[q] = has_lm ? (NS == 1 ? best : best + __ldg(p.lm_exit + q)) : best + 0.0f;
			if (BEAM) my_best = best;
			p.bp[(uint64_t)(off + s) * L + lab] = bptr < 0 ? (uint16_t)0xffff : (uint16_t)bptr;
			p.bd[(uint64_t)(off + s) * L + lab] = (uint8_t)bdur;
		}
		if (BEAM) {
			// minimum over the node (findMinWeight), then the kept flags and their ballots
			float mn = lab < L ? my_best : CUDART_INF_F;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
			if ((threadIdx.x & 31) == 0) s_wmin[threadIdx.x >> 5] = mn;
			__syncthreads();
			const uint32_t nw = (blockDim.x + 31) / 32;
			float bm = (threadIdx.x & 31) < nw ? s_wmin[threadIdx.x & 31] : CUDART_INF_F;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) bm = fminf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
			my_kept = lab < L && (double)my_best < (double)bm + p.beam;
			const uint32_t bal = __ballot_sync(0xffffffffu, my_kept);
			if ((threadIdx.x & 31) == 0) s_keptmask[threadIdx.x >> 5] = bal;
			if (lab < L) We[q] = my_kept ? (has_lm ? my_best : my_best + 0.0f) : CUDART_INF_F;
		}
		__syncthreads();
		VKTICK(2);   // node update + barrier
		// the kept-list order only moves for one state per phone (N > 1: always the identity, s_g stays 0xff): no bookkeeping, no barrier
		if (BEAM) {
			if (threadIdx.x == 0) {
				// order of kept(s) = a[max(0, s-D+1)] with the pruned phones left out; its head = the first kept phone in that order =
				// the smallest kept phone unless that is the one the descriptor moves to the back
				const uint32_t j2 = s + 1 >= D ? s + 1 - D : 0;
				const uint32_t g2 = (j2 == 0) ? 0xffffffffu : s_move_w[j2 & 255];
				uint32_t m1 = 0xffffffffu, m2 = 0xffffffffu;
				for (uint32_t w = 0; w < (P + 31) / 32 && m2 == 0xffffffffu; w++) {
					uint32_t b = s_keptmask[w];
					while (b && m2 == 0xffffffffu) { const uint32_t i = w * 32 + (uint32_t)__ffs((int)b) - 1; b &= b - 1; if (m1 == 0xffffffffu) m1 = i; else m2 = i; }
				}
				const uint32_t head = (m1 != g2 || m2 == 0xffffffffu) ? m1 : m2;
				s_move_w[(s + 1) & 255] = P > 1 ? head : 0xffffffffu;      // a[s+1]: ascending with the head of kept(s) arriving last
				s_gw = g2;                                                  // order of kept(s) feeds the cross scan of frame s+1
			}
			__syncthreads();
		} else if (NS == 1) {
			if (threadIdx.x == 0) {
				// a[s] := descriptor of the ARRIVAL order at start frame s; a[0] = identity.
				// a[s>=1] (NS==1): head of kept(s-1) moved to the back; kept(s-1) = a[max(0, s-1-D+1)] = a[max(0, s-D)].
				uint32_t a_s = 0xffu;
				if (P > 1 && s >= 1) {
					const uint32_t j = s >= D ? s - D : 0;
					const uint32_t aj = j == 0 ? 0xffu : s_move[j & 255];
					a_s = (aj == 0xffu) ? 0u : (aj == 0u ? 1u : 0u);     // head of the list described by aj
				}
				s_move[s & 255] = (uint8_t)a_s;
				// order of kept(s) feeds the cross scan of frame s+1: kept(s) = a[max(0, s-D+1)]
				const uint32_t j2 = s + 1 >= D ? s + 1 - D : 0;
				s_g = (j2 == 0) ? 0xffu : (j2 == s ? a_s : s_move[j2 & 255]);
			}
			__syncthreads();
		}
#pragma unroll
		for (uint32_t j = 0; j < VPF; j++) nsv[j] = nsn[j];
	}
	if (timing) { for (int i = 0; i < 4; i++) p.dbg[i] = tacc[i]; p.dbg[4] = T; }
	// ---- final argmin over the kept list (first wins) and traceback ----
	// The walk is one dependent load per segment; from global memory that was 0.3 us per segment (a fifth of the kernel at cfg3).
	// The CTA now copies the back-pointer rows of a WINDOW of frames into shared memory (coalesced, all threads), thread 0 walks
	// inside the window at shared-memory latency, and the segments (found last to first) are reversed by all threads at the end.
	__shared__ int s_end, s_cur; __shared__ uint32_t s_nseg;
	if (threadIdx.x == 0) {
		float minw = VIT_INF; int best = -1;
		if (T > 0) {
			const uint32_t g = BEAM ? s_gw : s_g;
			for (uint32_t i = 0; i < P; i++) {
				// free-phone LM: kept-list order, first wins.  Input LM: the decoder's finalStateSet, ordered by LM state = phone, weight =
				// hypothesis + final weight of its state, states that are not final left out (expandFinalNode :746-758, :2138-2153)
				const uint32_t ph = has_lm ? i : kept_phone(i, P, g, GNONE), e = ph * NS + NS - 1;
				if (BEAM && !((s_keptmask[ph >> 5] >> (ph & 31)) & 1u)) continue;      // pruned at the last node
				float w = Wprev[e];
				if (has_lm) { const float fw = __ldg(p.lm_final + i); if (isinf(fw)) continue; w = w + fw; }
				if (w < minw) { minw = w; best = (int)e; }
			}
		}
		p.cost[u] = minw;
		s_cur = best; s_end = best >= 0 ? (int)T - 1 : -1; s_nseg = 0;
	}
	__syncthreads();                       // the recursion's shared memory is free from here on
	// window buffers: the rows lo..end of bp / bd are ONE contiguous byte range each; it is copied in 16-byte vectors from the
	// 16-byte-aligned address below its start, so the data sits `shift` bytes into the (aligned) buffer
	const size_t bp_cap = ((size_t)p.tbW * L * 2 + 47) / 16 * 16;
	unsigned char* bpw_raw = smem_raw;                                   // [<= tbW * L * 2 + 32]
	unsigned char* bdw_raw = smem_raw + bp_cap;                          // [<= tbW * L + 32]
	auto copy_window = [&](unsigned char* dst, const unsigned char* src, size_t nbytes) -> uint32_t {
		const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
		const uint4* sv = reinterpret_cast<const uint4*>(src - shift);
		const uint32_t nvec = (uint32_t)((shift + nbytes + 15) / 16);
		for (uint32_t i = threadIdx.x; i < nvec; i += blockDim.x) reinterpret_cast<uint4*>(dst)[i] = sv[i];
		return shift;
	};
	uint32_t* ol = p.out_lab + off; uint32_t* od = p.out_dur + off; uint32_t* op = p.out_phn + off;
	while (s_end >= 0) {
		const int end0 = s_end, lo = max(0, end0 - (int)p.tbW + 1);
		const size_t cnt = (size_t)(end0 - lo + 1) * L;
		const uint32_t sh_p = copy_window(bpw_raw, reinterpret_cast<const unsigned char*>(p.bp + (uint64_t)(off + lo) * L), cnt * 2);
		const uint32_t sh_d = copy_window(bdw_raw, p.bd + (uint64_t)(off + lo) * L, cnt);
		const uint16_t* bpw = reinterpret_cast<const uint16_t*>(bpw_raw + sh_p);     // 2-byte aligned: bp rows are
		const uint8_t* bdw = bdw_raw + sh_d;
		__syncthreads();
		if (threadIdx.x == 0) {
			int end = end0, cur = s_cur; uint32_t nseg = s_nseg;
			while (end >= lo) {
				const uint32_t d = bdw[(uint32_t)(end - lo) * L + cur];
				const uint16_t prev = bpw[(uint32_t)(end - lo) * L + cur];
				const int start = end + 1 - (int)d;
				ol[nseg] = (uint32_t)cur; od[nseg] = d;
				if (start == 0) op[nseg] = (uint32_t)cur / NS;
				else op[nseg] = ((uint32_t)cur % NS == 0 && (int)prev != cur) ? (uint32_t)cur / NS : LAB_BAD;
				nseg++;
				if (start == 0) { end = -1; break; }
				cur = (int)prev; end = start - 1;
			}
			s_end = end; s_cur = cur; s_nseg = nseg;
		}
		__syncthreads();
	}
	const uint32_t nseg = s_nseg;
	for (uint32_t i = threadIdx.x; i < nseg / 2; i += blockDim.x) {
		uint32_t a;
		a = ol[i]; ol[i] = ol[nseg - 1 - i]; ol[nseg - 1 - i] = a;
		a = od[i]; od[i] = od[nseg - 1 - i]; od[nseg - 1 - i] = a;
		a = op[i]; op[i] = op[nseg - 1 - i]; op[nseg - 1 - i] = a;
	}
	if (threadIdx.x == 0) p.n_seg[u] = nseg;
}
void launch_frame_tables(const uint32_t* off, uint32_t n_utt, uint32_t N, uint32_t* ft, uint32_t* fu, uint32_t* fl, cudaStream_t s) {
	if (!N) return;
	frame_tables_kernel<<<(N + 255) / 256, 256, 0, s>>>(off, n_utt, N, ft, fu, fl);
}

void launch_viterbi(const VitParams& p, cudaStream_t s) {
	if (!p.n_utt) return;
	VitParams q = p;
	const uint32_t Pw = (p.P + 31) / 32 * 32;
	q.Ppad = (p.NS > 1 && p.NS * Pw <= 1024) ? Pw : p.P;
	const unsigned threads = (std::max(p.L, p.NS * q.Ppad) + 31) / 32 * 32;
	size_t smem = sizeof(float) * (3 * (size_t)p.L + p.P);
	if (p.negMt != nullptr && 2 * (size_t)p.P * (p.P | 1u) * sizeof(float) <= 96 * 1024) smem += 2 * sizeof(float) * (size_t)p.P * (p.P | 1u);   // per-frame tables: double buffer
	if (p.negMt == nullptr && (size_t)p.P * (p.P | 1u) * sizeof(float) <= 96 * 1024) {
		smem += sizeof(float) * (size_t)p.P * (p.P | 1u);
		if (p.lm_bigT != nullptr && 2 * (size_t)p.P * (p.P | 1u) * sizeof(float) <= 96 * 1024) smem += sizeof(float) * (size_t)p.P * (p.P | 1u);
	}
	// traceback window: rows of back pointers (2 + 1 bytes per label) of as many frames as fit 48 KB, at least 8
	q.tbW = std::max<uint32_t>(8u, (48u * 1024u) / (3u * p.L));
	smem = std::max(smem, ((size_t)q.tbW * p.L * 2 + 47) / 16 * 16 + (size_t)q.tbW * p.L + 48);
	auto go = [&](auto kern) {
		cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		kern<<<p.n_utt, threads, smem, s>>>(q);
	};
	const bool lm = p.lm_start != nullptr, beam = p.beam > 0.0;      // (beam: one state per phone, checked by the caller)
	if (lm && beam) go(viterbi_kernel<true, true>);
	else if (lm) go(viterbi_kernel<true, false>);
	else if (beam) go(viterbi_kernel<false, true>);
	else go(viterbi_kernel<false, false>);
}

}  // namespace crfgpu
