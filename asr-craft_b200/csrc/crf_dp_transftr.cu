// Frame-level CRF with transition FEATURES (crf_featuremap=stdtrans, one state per label) for sm_100a.
//
// Reference semantics: CRF_StdStateNode::computeTransMatrix / computeAlpha / computeBeta / computeExpF
// (CRF/src/nodes/CRF_StdStateNode.cpp:58-72, 81-128, 140-204, 221-277) with the frame-dependent transition scores of
// CRF_StdFeatureMap::computeTransMatrixValue (CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:94-110):
//   M_t[p][c] = x_t(trans slice) . lambda_t[p->c] + lambda_bias[p->c] * transBiasVal
//   alpha_t[c] = S_t[c] + logsum_p (alpha_{t-1}[p] + M_t[p][c]),  alpha_0 = S_0;   logZ = logsum_c alpha_{T-1}[c]
//   beta_t[p]  = logsum_c (M_{t+1}[p][c] + S_{t+1}[c] + beta_{t+1}[c]),  beta_{T-1} = 0
//   gamma_t[c] = exp(alpha_t[c] + beta_t[c] - logZ),   xi_t[p][c] = exp(alpha_{t-1}[p] + M_t[p][c] + S_t[c] + beta_t[c] - logZ)  (t >= 1)
// and computeTransExpF (:197-223): ExpF_trans[p->c][f] += xi_t[p][c] * x_t[f].
//
// Device plan: the L*L transition scores of every frame are ONE tensor-core GEMM [frames x features] . [features x L*L]
// (launch_score_gemm_tc, crf_tc_gemm.cu) into M[N][L*L]; the recursions below stream that matrix, one CTA per utterance, the
// next frame's matrix fetched with one bulk copy (cp.async.bulk, mbarrier) while the current one is used; the posteriors leave as Dm = [ref] - gamma and
// Xd = [ref pair] - xi, so that the state and transition gradients (and their empirical counts) are two more tensor-core GEMMs
// (launch_reduce_gemm_tc).  gamma and xi are normalised by their own per-frame sums, which are 1 in exact arithmetic.
#include <cfloat>
#include <cstdlib>
#include <type_traits>

#include "crf_kernels.cuh"
#include "tc05.cuh"

namespace crfgpu {

namespace {

// CTA sizes: thread = label.  128 threads up to 128 labels, 192 beyond (the double-buffered L x L tile bounds the label count by shared memory:
// 169 labels frame-level, 160 segmental)
constexpr int TF_MAX_THR = 192;

template <int TF_THR>
__device__ __forceinline__ float block_max(float v, float* scratch) {
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	__syncthreads();
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	float r = scratch[0];
	for (int w = 1; w < TF_THR / 32; w++) r = fmaxf(r, scratch[w]);
	return r;
}
template <int TF_THR>
__device__ __forceinline__ float block_sum(float v, float* scratch) {
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	float r = 0.0f;
	for (int w = 0; w < TF_THR / 32; w++) r += scratch[w];
	return r;
}

// Pre-pass over all frames, fully parallel: E_n = exp(M_n - max M_n) and the row maximum.  Inside the recursions that was L*L exp per
// frame on the dependent chain (48 of them per thread and frame in the forward matrix-vector product alone).
__global__ void __launch_bounds__(256) transftr_exp_kernel(const float* __restrict__ M, float* __restrict__ E, float* __restrict__ rowmax, uint32_t N, uint32_t LL, uint32_t Lq) {
	const uint32_t warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (warp >= N) return;
	const float* m = M + (size_t)warp * Lq;
	float* e = E + (size_t)warp * Lq;
	float mx = -INFINITY;
	for (uint32_t i = lane; i < LL; i += 32) mx = fmaxf(mx, m[i]);
	for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	for (uint32_t i = lane; i < LL; i += 32) e[i] = __expf(m[i] - mx);
	if (lane == 0) rowmax[warp] = mx;
}

// The same pre-pass with ONE read of M: a CTA per row, the row's RPT * 256 elements wait in registers between the maximum and the
// exponentials (the warp-per-row kernel above reads every row twice, and with 8 rows of 15 KB per CTA and 8 CTAs per SM the second
// read misses L1 and L2: ncu 4.1 GB read for 2.0 GB of scores, profiles/r2z4_frame_transftr_kernels_full.md).  Same values bit for bit.
template <int RPT>
__global__ void __launch_bounds__(256) transftr_exp_row_kernel(const float* __restrict__ M, float* __restrict__ E, float* __restrict__ rowmax, uint32_t N, uint32_t LL, uint32_t Lq) {
	__shared__ float wmax[2][8];
	uint32_t it = 0;
	for (uint32_t n = blockIdx.x; n < N; n += gridDim.x, it ^= 1) {
		const float* m = M + (size_t)n * Lq;
		float* e = E + (size_t)n * Lq;
		float x[RPT], mx = -INFINITY;
#pragma unroll
		for (int j = 0; j < RPT; j++) { const uint32_t i = threadIdx.x + j * 256; x[j] = i < LL ? __ldcs(m + i) : -INFINITY; }
#pragma unroll
		for (int j = 0; j < RPT; j++) mx = fmaxf(mx, x[j]);
		for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
		if ((threadIdx.x & 31) == 0) wmax[it][threadIdx.x >> 5] = mx;
		__syncthreads();      // (one barrier per row: the other half of wmax is only rewritten behind the next row's barrier)
		float r = wmax[it][0];
#pragma unroll
		for (int w = 1; w < 8; w++) r = fmaxf(r, wmax[it][w]);
#pragma unroll
		for (int j = 0; j < RPT; j++) { const uint32_t i = threadIdx.x + j * 256; if (i < LL) e[i] = __expf(x[j] - r); }
		if (threadIdx.x == 0) rowmax[n] = r;
	}
}

// The recursions keep the frame's matrix DENSE in shared memory -- exactly the Lq floats of the frame's row in global memory -- so a
// frame is ONE bulk copy issued by one thread and counted on an mbarrier (cp.async.bulk), instead of L * L / threads 4-byte cp.async
// with their index arithmetic per thread (about 40 % of the forward step's instructions at 61 labels).
__device__ __forceinline__ void bulk_matrix(float* dst, const float* src, uint32_t Lq, uint64_t* bar) {
	tc05::mbar_arrive_expect_tx(bar, Lq * 4u);
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(tc05::smem_u32(dst)), "l"(src), "r"(Lq * 4u), "r"(tc05::smem_u32(bar)) : "memory");
}

template <int TF_THR>
__global__ void __launch_bounds__(TF_THR) transftr_forward_kernel(TransFtrParams p) {
	extern __shared__ __align__(16) float sm[];
	const uint32_t L = p.L, Lq = p.Lq;
	float* Ms = sm;                              // [2][Lq] dense: element (q, c) at q * L + c
	float* a_prev = sm + 2 * Lq;                 // [L]
	float* scratch = a_prev + L;                 // [8]
	__shared__ uint64_t mbar[2];                 // matrix of frame t >= 1 arrived in buffer t & 1 (use (t - 1) >> 1 of that buffer: frame 0 has no matrix)
	const uint32_t u = blockIdx.x, off = p.off[u], T = p.off[u + 1] - off, c = threadIdx.x;
	double rho = 0.0;
	if (c == 0) { tc05::mbar_init(&mbar[0], 1); tc05::mbar_init(&mbar[1], 1); tc05::fence_mbar_init(); }
	__syncthreads();
	if (c == 0 && T > 1) bulk_matrix(Ms + Lq, p.E + (size_t)(off + 1) * Lq, Lq, &mbar[1]);       // frame 1 -> buffer 1
	// the frame's score and matrix maximum do not depend on the recursion: those of frame t+1 are requested while frame t is processed
	float s_next = (c < L && T > 0) ? p.S[(size_t)off * p.Lp + c] : 0.0f, mmax_next = 0.0f;
	for (uint32_t t = 0; t < T; t++) {
		const size_t n = (size_t)off + t;
		float* Mt = Ms + (t & 1) * Lq;
		float w = -INFINITY;
		const float s = s_next;
		float mmax = mmax_next;
		if (t + 1 < T) { s_next = c < L ? p.S[(n + 1) * p.Lp + c] : 0.0f; mmax_next = p.rowmax[n + 1]; }
		if (t == 0) { if (c < L) w = s; }
		else {
			// (the other buffer was last read in step t-1, which every thread left through the step's closing barrier)
			if (c == 0 && t + 1 < T) bulk_matrix(Ms + ((t + 1) & 1) * Lq, p.E + (n + 1) * Lq, Lq, &mbar[(t + 1) & 1]);
			tc05::mbar_wait(&mbar[t & 1], ((t - 1) >> 1) & 1);
			if (c < L) {
				float v = 0.0f;
#pragma unroll 8
				for (uint32_t q = 0; q < L; q++) v = fmaf(a_prev[q], Mt[q * L + c], v);      // Mt = exp(M_t - mmax) from the pre-pass
				w = __logf(v) + s;
			}
		}
		const float wmax = block_max<TF_THR>(w, scratch);
		const float a = c < L ? __expf(w - wmax) : 0.0f;
		const float asum = block_sum<TF_THR>(a, scratch);
		rho += (double)mmax + (double)wmax + (double)__logf(asum);
		if (c < L) { a_prev[c] = a / asum; p.A[n * p.Lp + c] = a / asum; }
		if (c == 0) p.rho[n] = rho;      // (the numerator -- dependent global loads of ONE thread per frame -- is a pass of its own: transftr_numer_kernel)
		__syncthreads();
	}
	if (c == 0) p.logZ[u] = rho;       // sum_c a_{T-1}[c] = 1: alpha_{T-1} sums to exp(rho)
}

// numerator of an utterance on the reference path: state score of every frame's label + transition score from the previous frame's
// label (pairs an N-state map lacks count nothing); one warp per utterance, lanes stride the frames, fixed summation order
__global__ void __launch_bounds__(128) transftr_numer_kernel(TransFtrParams p) {
	const uint32_t u = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (u >= p.n_utt) return;
	const uint32_t off = p.off[u], T = p.off[u + 1] - off, L = p.L;
	double acc = 0.0;
	for (uint32_t t = lane; t < T; t += 32) {
		const size_t n = (size_t)off + t;
		const uint32_t y = p.labs[n];
		if (y >= L) continue;
		acc += (double)p.S[n * p.Lp + y];
		if (t > 0) { const uint32_t yp = p.labs[n - 1]; if (yp < L && p.tidx[yp * L + y] != 0xffffffffu) acc += (double)p.M[n * p.Lq + yp * L + y]; }
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if (lane == 0) p.numer[u] = acc;
}

// One step of the backward recursion needs three reductions over the CTA -- sum_{q,c} alpha_{t-1}[q] E_t[q][c] w_t[c] (the normaliser of
// xi_t), the maximum of the new beta vector, and the maximum of the NEXT frame's scores (data-independent, in registers one frame
// ahead) -- and they are ONE pass with one barrier: the scratch words are only rewritten behind the step's closing barrier.
template <int TF_THR>
__device__ __forceinline__ void block_sum_max_max(float& s, float& m1, float& m2, float* scratch) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		s += __shfl_xor_sync(0xffffffffu, s, o);
		m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
		m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
	}
	if ((threadIdx.x & 31) == 0) { scratch[threadIdx.x >> 5] = s; scratch[8 + (threadIdx.x >> 5)] = m1; scratch[16 + (threadIdx.x >> 5)] = m2; }
	__syncthreads();
	s = scratch[0]; m1 = scratch[8]; m2 = scratch[16];
#pragma unroll
	for (int w = 1; w < TF_THR / 32; w++) { s += scratch[w]; m1 = fmaxf(m1, scratch[8 + w]); m2 = fmaxf(m2, scratch[16 + w]); }
}

template <int TF_THR>
__global__ void __launch_bounds__(TF_THR) transftr_backward_kernel(TransFtrParams p) {
	extern __shared__ __align__(16) float sm[];
	const uint32_t L = p.L, Lq = p.Lq;
	float* Ms = sm;                              // [2][Lq] dense: element (q, c) at q * L + c
	__shared__ uint64_t mbar[2];                 // matrix of frame t arrived in buffer t & 1 (use (T - 1 - t) >> 1 of that buffer)
	float* b = sm + 2 * Lq;                      // [L] row sums of the step = beta_{t-1} before its normalisation (row threads -> label threads)
	float* wv = b + L;                           // [L] exp(S_t - smax) * beta_t
	float* av = wv + L;                          // [L] alpha_{t-1}
	float* scratch = av + L;                     // [24]
	const uint32_t u = blockIdx.x, off = p.off[u], T = p.off[u + 1] - off, c = threadIdx.x;
	// (pair of element i = c + k * TF_THR of the L x L matrix: a constant step with a carry, no division per element)
	const uint32_t q_first = c / L, c_first = c - q_first * L, dq_ = TF_THR / L, dc_ = TF_THR - dq_ * L;
	// second identity of a thread: TPR threads share row rq of the matrix (E_t[rq][.] * w_t summed over the row IS beta_{t-1}[rq]:
	// the row sums fall out of the pass that scales the matrix, not out of a loop of their own over L dependent additions)
	const uint32_t TPR = TF_THR / L >= 4 ? 4u : TF_THR / L >= 2 ? 2u : 1u;
	const uint32_t rq = c / TPR, rpart = c - rq * TPR;
	const bool row_ok = rq < L;
	// rows of an even label count all start in the same few banks: row rq then starts its pass rq columns further on (wrapping round)
	const uint32_t r_first = row_ok ? (rpart + ((L & 1u) ? 0u : rq)) % L : 0u, r_cnt = (row_ok && rpart < L) ? (L - rpart + TPR - 1) / TPR : 0u;
	if (!T) return;
	if (c == 0) { tc05::mbar_init(&mbar[0], 1); tc05::mbar_init(&mbar[1], 1); tc05::fence_mbar_init(); }
	__syncthreads();
	if (c == 0 && T > 1) bulk_matrix(Ms + ((T - 1) & 1) * Lq, p.E + (size_t)(off + T - 1) * Lq, Lq, &mbar[(T - 1) & 1]);
	// the frame's label, alpha entry and score do not depend on the recursion: those of frame t-1 are requested while frame t is processed
	uint32_t y_cur = p.labs[(size_t)off + T - 1];
	float a_cur = c < L ? p.A[((size_t)off + T - 1) * p.Lp + c] : 0.0f, s_cur = c < L ? p.S[((size_t)off + T - 1) * p.Lp + c] : -INFINITY;
	float bc = c < L ? 1.0f : 0.0f;                              // beta_t[c], max-normalised (setTailBeta)
	float gsum = block_sum<TF_THR>(a_cur * bc, scratch);         // sum_c alpha_t[c] beta_t[c]
	float smax = block_max<TF_THR>(s_cur, scratch);
	__syncthreads();
	for (uint32_t t = T; t-- > 0;) {
		const size_t n = (size_t)off + t;
		const uint32_t y = y_cur;
		uint32_t y_prev = LAB_BAD; float a_prev1 = 0.0f, s_prev = -INFINITY;
		if (t > 0) { y_prev = p.labs[n - 1]; if (c < L) { a_prev1 = p.A[(n - 1) * p.Lp + c]; s_prev = p.S[(n - 1) * p.Lp + c]; } }
		// gamma_t = A_t * b_t / sum
		if (c < L) p.Dm[n * p.Lp + c] = ((y == c) ? 1.0f : 0.0f) - a_cur * bc / gsum;
		float* xrow = p.Xd + n * p.Lq;
		if (t == 0) {
			for (uint32_t i = c; i < L * L; i += TF_THR) xrow[i] = 0.0f;      // no transition enters the first frame
			break;
		}
		float* Mt = Ms + (t & 1) * Lq;
		if (c < L) { wv[c] = __expf(s_cur - smax) * bc; av[c] = a_prev1; }
		// (the other buffer was last touched in step t+1, which every thread left through the step's closing barrier behind its proxy fence)
		if (c == 0 && t > 1) bulk_matrix(Ms + ((t - 1) & 1) * Lq, p.E + (n - 1) * Lq, Lq, &mbar[(t - 1) & 1]);
		tc05::mbar_wait(&mbar[t & 1], ((T - 1 - t) >> 1) & 1);
		__syncthreads();
		// Mt = E_t = exp(M_t - max) from the pre-pass (the normalisations below are scale-free) -> E_t[q][cc] * w_t[cc] in place
		float rs = 0.0f;
		{
			float* mrow = Mt + rq * L;
			uint32_t cc = r_first;
#pragma unroll 4
			for (uint32_t k = 0; k < r_cnt; k++) {
				const float e = mrow[cc] * wv[cc]; mrow[cc] = e; rs += e;
				cc += TPR; if (cc >= L) cc -= L;
			}
		}
		for (uint32_t o = 1; o < TPR; o <<= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
		if (row_ok && rpart == 0) b[rq] = rs;
		// xsum = sum_{q,cc} alpha_{t-1}[q] E_t[q][cc] w_t[cc];  bmax = max_q beta_{t-1}[q];  the next frame's score maximum
		float xsum = (row_ok && rpart == 0) ? av[rq] * rs : 0.0f, bmax = row_ok ? rs : 0.0f, smax_next = s_prev;
		block_sum_max_max<TF_THR>(xsum, bmax, smax_next, scratch);
		uint32_t yp = y_prev;
		if (yp < L && y < L && p.tidx[yp * L + y] == 0xffffffffu) yp = LAB_BAD;      // a reference pair the N-state map does not have
		const float inv = 1.0f / xsum;
		uint32_t q = q_first, cc = c_first;
		for (uint32_t i = c; i < L * L; i += TF_THR) {
			xrow[i] = ((q == yp && cc == y) ? 1.0f : 0.0f) - av[q] * Mt[i] * inv;
			cc += dc_; q += dq_;
			if (cc >= L) { cc -= L; q++; }
		}
		// beta_{t-1}, max-normalised; sum_q alpha_{t-1}[q] beta_{t-1}[q] is the xi normaliser over the same maximum
		bc = c < L ? b[c] / bmax : 0.0f;
		gsum = xsum / bmax; smax = smax_next;
		y_cur = y_prev; a_cur = a_prev1; s_cur = s_prev;
		tc05::fence_proxy_async_smem();      // this step's in-place scaling (generic stores) before the bulk copy that reuses the buffer
		__syncthreads();
	}
}

}  // namespace

void launch_transftr_exp(const float* M, float* E, float* rowmax, uint32_t N, uint32_t LL, uint32_t Lq, cudaStream_t s) {
	if (!N) return;
	const uint32_t grid = N < 148u * 8u ? N : 148u * 8u;      // rows strided over the resident CTAs
	if (LL <= 256u * 10u) transftr_exp_row_kernel<10><<<grid, 256, 0, s>>>(M, E, rowmax, N, LL, Lq);       // up to 50 labels (the recipe's 48 phones)
	else if (LL <= 256u * 16u) transftr_exp_row_kernel<16><<<grid, 256, 0, s>>>(M, E, rowmax, N, LL, Lq);  // up to 64 labels
	else if (LL <= 256u * 40u) transftr_exp_row_kernel<40><<<grid, 256, 0, s>>>(M, E, rowmax, N, LL, Lq);  // up to 101 labels
	else transftr_exp_kernel<<<(N + 7) / 8, 256, 0, s>>>(M, E, rowmax, N, LL, Lq);
}

size_t transftr_smem_bytes(uint32_t L) { return sizeof(float) * ((size_t)2 * ((L * L + 3) / 4 * 4) + 3 * (size_t)L + 32); }      // two dense matrices of Lq floats + the vectors

cudaError_t launch_transftr_dp(bool backward, const TransFtrParams& p, cudaStream_t s) {
	if (!p.n_utt) return cudaSuccess;
	const size_t smem = transftr_smem_bytes(p.L);
	if (p.L > (uint32_t)TF_MAX_THR) return cudaErrorInvalidValue;
	auto go = [&](auto kern, int thr) -> cudaError_t {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return e;
		kern<<<p.n_utt, thr, smem, s>>>(p);
		return cudaGetLastError();
	};
	// up to 128 labels: 256 threads share the L x L loops of the backward step (2.11 -> 1.68 ms at 64 utterances, 2.94 -> 2.77 ms at 462);
	// the forward step has no such loop beyond the matrix prefetch: 256 threads only while every CTA has an SM's issue slots to itself
	if (backward) return p.L <= 128 ? go(transftr_backward_kernel<256>, 256) : go(transftr_backward_kernel<TF_MAX_THR>, TF_MAX_THR);
	transftr_numer_kernel<<<(p.n_utt + 3) / 4, 128, 0, s>>>(p);          // reads S, M, labs only: ahead of the recursion on the same stream
	if (p.L > 128) return go(transftr_forward_kernel<TF_MAX_THR>, TF_MAX_THR);
	return p.n_utt <= 2 * 148 ? go(transftr_forward_kernel<256>, 256) : go(transftr_forward_kernel<128>, 128);
}


// =================================================================================================
// Segmental model WITHOUT duration labels and WITH transition features (stdseg_no_dur_no_segtransftr + stdtrans: the production TIMIT
// recipe, demo/segmental-timit-demo.cfg.in:11-48).  Reference: CRF_StdSegStateNode_WithoutDurLab_WithoutSegTransFtr
// (CRF/src/nodes/CRF_StdSegStateNode_WithoutDurLab_WithoutSegTransFtr.cpp: computeTransMatrix :39-121 -- M_t[y'][y] from the duration-1
// window of frame t --, computeAlpha :123-333, computeAlphaPlusTrans :1077-1116, computeBeta :395-614, computeExpF :616-1066 -- the
// transition counts of node t use node t+1's features and the NEXT reference label) driven by CRF_NewGradBuilder_StdSeg_NoDur_NoTrans.
//   A_t[y]     = logsum_y' (alpha_t[y'] + M_{t+1}[y'][y])         alpha_t[d,y] = S_t[d,y] + A_{t-d}[y]   (= S_t[d,y] when d == t+1)
//   B_t[y]     = logsum_d (S_{t+d}[d,y] + beta_{t+d}[y])          beta_t[y'] = logsum_y (M_{t+1}[y'][y] + B_t[y])
//   gamma_t[d,y] = exp(alpha_t[d,y] + beta_t[y] - logZ)           xi_t[y'][y]  = exp(alpha_t[y'] + M_{t+1}[y'][y] + B_t[y] - logZ)
// One CTA per utterance, thread = phone (<= 128), M_{t+1} fetched ahead with one bulk copy, alpha_t normalised to sum 1 with a running
// log scale rho_t, the D-term sums formed as float log-sum-exps of differences of the (double) scales.
// =================================================================================================
namespace {

constexpr uint32_t ND_RING = 32;     // max_dur <= 31

template <int TF_THR, int DR>
__global__ void __launch_bounds__(TF_THR) nodur_tf_forward_kernel(NodurTfParams p) {
	extern __shared__ __align__(16) float sm[];
	const uint32_t P = p.P, D = p.D, Lq = p.Lq;
	double* rring = reinterpret_cast<double*>(sm);            // [ND_RING] rho_t (first: 8-byte aligned)
	float* Ms = sm + 2 * ND_RING;                // [2][Lq] dense (element (q, y) at q * P + y): a frame is one bulk copy, as in the frame-level kernels
	float* av = Ms + 2 * Lq;                     // [P] alpha_t (sum 1)
	__shared__ uint64_t mbar[2];                 // matrix of frame f >= 1 arrived in buffer f & 1 (use (f - 1) >> 1 of that buffer)
	float* lgh = av + P;                         // [ND_RING][P] log A_t[y] - rho_t of the last D frames
	float* scratch = lgh + ND_RING * P;          // [8]
	const uint32_t u = blockIdx.x, off = p.off[u], T = p.off[u + 1] - off, y = threadIdx.x;
	if (y == 0) { tc05::mbar_init(&mbar[0], 1); tc05::mbar_init(&mbar[1], 1); tc05::fence_mbar_init(); }
	__syncthreads();
	if (y == 0 && T > 1) bulk_matrix(Ms + Lq, p.E + (size_t)(off + 1) * Lq, Lq, &mbar[1]);       // exp(M_1 - max) -> buffer 1
	// the score terms of a frame do not depend on the recursion: those of frame t+1 are requested while frame t is processed
	float sv[ND_RING];
#pragma unroll
	for (uint32_t d = 1; d <= (uint32_t)DR; d++) sv[d] = (y < P && T > 0 && d <= min(1u, D)) ? p.S[(size_t)off * p.Lp + (size_t)(d - 1) * P + y] : 0.0f;
	for (uint32_t t = 0; t < T; t++) {
		const size_t n = (size_t)off + t;
		const uint32_t dmax = min(t + 1, D);
		float sn[ND_RING];
		{
			const uint32_t dn = (t + 1 < T && y < P) ? min(t + 2, D) : 0;
#pragma unroll
			for (uint32_t d = 1; d <= (uint32_t)DR; d++) sn[d] = d <= dn ? p.S[(n + 1) * p.Lp + (size_t)(d - 1) * P + y] : 0.0f;
		}
		const float mmax = t + 1 < T ? p.rowmax[n + 1] : 0.0f;      // (requested here: used behind the matrix-vector product)
		const double rref = t > 0 ? rring[(t - 1) & (ND_RING - 1)] : 0.0;
		// w[y] = log sum_d exp(S_t[d,y] + A_{t-d}[y] - rref)
		float w = -INFINITY;
		if (y < P) {
			float lt[ND_RING];
			float mx = -INFINITY;
#pragma unroll
			for (uint32_t d = 1; d <= (uint32_t)DR; d++) {
				if (d <= dmax) {
					float v = sv[d];
					if (d <= t) v += lgh[((t - d) & (ND_RING - 1)) * P + y] + (float)(rring[(t - d) & (ND_RING - 1)] - rref);
					else v += (float)(-rref);
					lt[d - 1] = v; mx = fmaxf(mx, v);
				}
			}
			float sacc = 0.0f;
#pragma unroll
			for (uint32_t d = 1; d <= (uint32_t)DR; d++) if (d <= dmax) sacc += __expf(lt[d - 1] - mx);
			w = mx + __logf(sacc);
		}
		const float wmax = block_max<TF_THR>(w, scratch);
		const float a = y < P ? __expf(w - wmax) : 0.0f;
		const float asum = block_sum<TF_THR>(a, scratch);
		const double rho = rref + (double)wmax + (double)__logf(asum);
		if (y < P) { av[y] = a / asum; p.A[n * p.Pp + y] = a / asum; }
		if (y == 0) { rring[t & (ND_RING - 1)] = rho; p.rho[n] = rho; }
		// (the numerator -- scores of the reference path -- is three dependent global loads per segment for one thread: a pass of its
		// own, nodur_tf_numer_kernel, instead of a stall of the whole CTA at the frame's barrier)
		__syncthreads();
		if (t + 1 < T) {
			// A_t[y] = rho_t + lgh_t[y],  lgh_t[y] = mmax + log sum_q alpha_t[q] exp(M_{t+1}[q][y] - mmax)
			float* Mn = Ms + ((t + 1) & 1) * Lq;
			// (buffer t & 1 held frame t's matrix, last read in step t-1, which every thread left through the step's closing barrier)
			if (y == 0 && t + 2 < T) bulk_matrix(Ms + (t & 1) * Lq, p.E + (n + 2) * Lq, Lq, &mbar[t & 1]);
			tc05::mbar_wait(&mbar[(t + 1) & 1], (t >> 1) & 1);
			if (y < P) {
				float v = 0.0f;
#pragma unroll 8
				for (uint32_t q = 0; q < P; q++) v = fmaf(av[q], Mn[q * P + y], v);           // Mn = exp(M_{t+1} - mmax) from the pre-pass
				const float l = mmax + __logf(v);
				lgh[(t & (ND_RING - 1)) * P + y] = l; p.LG[n * p.Pp + y] = l;
			}
			__syncthreads();
		}
#pragma unroll
		for (uint32_t d = 1; d <= (uint32_t)DR; d++) sv[d] = sn[d];
	}
	if (y == 0) p.logZ[u] = T ? rring[(T - 1) & (ND_RING - 1)] : 0.0;
}

// numerator of an utterance: state scores of the reference segments + transition scores between consecutive reference segments (taken
// from the frame the next segment starts in); one warp per utterance, lanes stride the frames, fixed summation order
__global__ void __launch_bounds__(128) nodur_tf_numer_kernel(NodurTfParams p) {
	const uint32_t u = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (u >= p.n_utt) return;
	const uint32_t off = p.off[u], T = p.off[u + 1] - off, P = p.P;
	double acc = 0.0;
	for (uint32_t t = lane; t < T; t += 32) {
		const size_t n = (size_t)off + t;
		const uint32_t lab = p.node_lab[n];
		if (lab == LAB_BAD) continue;
		acc += (double)p.S[n * p.Lp + lab];
		if (t + 1 < T) {
			const uint32_t nl = p.next_lab[n];
			if (nl != LAB_BAD && p.tidx[(lab % P) * P + nl] != 0xffffffffu) acc += (double)p.M[(n + 1) * p.Lq + (lab % P) * P + nl];
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if (lane == 0) p.numer[u] = acc;
}

template <int TF_THR, int DR>
__global__ void __launch_bounds__(TF_THR) nodur_tf_backward_kernel(NodurTfParams p) {
	extern __shared__ __align__(16) float sm[];
	const uint32_t P = p.P, D = p.D, Lq = p.Lq;
	double* kring = reinterpret_cast<double*>(sm);            // [ND_RING] kappa_t (first: 8-byte aligned)
	float* Ms = sm + 2 * ND_RING;                // [2][Lq] dense (element (q, y) at q * P + y): a frame is one bulk copy
	float* av = Ms + 2 * Lq;                     // [P] alpha_t
	__shared__ uint64_t mbar[2];                 // matrix of frame f arrived in buffer f & 1 (use (T - 1 - f) >> 1 of that buffer)
	float* lbh = av + P;                         // [ND_RING][P] beta_t[y] - kappa_t of the last D frames
	float* ev = lbh + ND_RING * P;               // [P] exp(B_t[y] - its maximum)
	float* scratch = ev + P;                     // [16]
	float* rsum = scratch + 16;                  // [P] row sums of the scaled matrix (row threads -> phone threads)
	const uint32_t u = blockIdx.x, off = p.off[u], T = p.off[u + 1] - off, y = threadIdx.x;
	const double lz = p.logZ[u];
	// (pair of element i = y + k * TF_THR of the P x P matrix: a constant step with a carry, no division per element)
	const uint32_t q_first = y / P, y_first = y - q_first * P, dq_ = TF_THR / P, dy_ = TF_THR - dq_ * P;
	if (y == 0) { tc05::mbar_init(&mbar[0], 1); tc05::mbar_init(&mbar[1], 1); tc05::fence_mbar_init(); }
	__syncthreads();
	if (y == 0 && T > 1) bulk_matrix(Ms + ((T - 1) & 1) * Lq, p.E + (size_t)(off + T - 1) * Lq, Lq, &mbar[(T - 1) & 1]);    // exp(M_{T-1} - max)
	// second identity of a thread, as in the frame-level kernel: TPR threads share row rq of the matrix, whose sum over the scaled row IS
	// beta_t[rq] before its normalisation; row rq of an even phone count starts rq columns further on (wrapping round): the rows of a
	// dense matrix would share their banks
	const uint32_t TPR = TF_THR / P >= 4 ? 4u : TF_THR / P >= 2 ? 2u : 1u;
	const uint32_t rq = y / TPR, rpart = y - rq * TPR;
	const bool row_ok = rq < P;
	const uint32_t r_first = row_ok ? (rpart + ((P & 1u) ? 0u : rq)) % P : 0u, r_cnt = (row_ok && rpart < P) ? (P - rpart + TPR - 1) / TPR : 0u;
	// the score terms S_{t+d}[d,y] of a frame do not depend on the recursion: those of frame t-1 are requested while frame t is processed
	float sv[ND_RING], sn[ND_RING];
#pragma unroll
	for (uint32_t d = 1; d <= (uint32_t)DR; d++) sv[d] = 0.0f;      // the tail frame has no successor terms
	// ... and so are the frame's labels, forward scale, matrix maximum and alpha entry (none is needed at the tail frame)
	uint32_t lab = LAB_BAD, c_nl = LAB_BAD; double c_rho = 0.0; float c_mmax = 0.0f, c_av = 0.0f;
	for (uint32_t t = T; t-- > 0;) {
		const size_t n = (size_t)off + t;
		const uint32_t nn = min(T - 1 - t, D);
		uint32_t n_lab = LAB_BAD, n_nl = LAB_BAD; double n_rho = 0.0; float n_mmax = 0.0f, n_av = 0.0f;
		if (t > 0) {
			n_lab = p.node_lab[n - 1]; n_nl = p.next_lab[n - 1]; n_rho = p.rho[n - 1]; n_mmax = p.rowmax[n];
			n_av = y < P ? p.A[(n - 1) * p.Pp + y] : 0.0f;
		}
		{
			const uint32_t dn = (t > 0 && y < P) ? min(T - t, D) : 0;      // frame t-1: d <= T-1-(t-1)
#pragma unroll
			for (uint32_t d = 1; d <= (uint32_t)DR; d++) sn[d] = d <= dn ? p.S[(n - 1 + d) * p.Lp + (size_t)(d - 1) * P + y] : 0.0f;
		}
		float lb = 0.0f;                         // beta_t[y] - kappa_t; tail: beta = 0
		double kappa = 0.0;
		if (nn > 0) {
			// w[y] = B_t[y] - kref = log sum_d exp(S_{t+d}[d,y] + beta_{t+d}[y] - kref)
			const double kref = kring[(t + 1) & (ND_RING - 1)];
			float w = -INFINITY;
			if (y < P) {
				float lt[ND_RING];
				float mx = -INFINITY;
#pragma unroll
				for (uint32_t d = 1; d <= (uint32_t)DR; d++) {
					if (d <= nn) {
						const float v = sv[d] + lbh[((t + d) & (ND_RING - 1)) * P + y] + (float)(kring[(t + d) & (ND_RING - 1)] - kref);
						lt[d - 1] = v; mx = fmaxf(mx, v);
					}
				}
				float sacc = 0.0f;
#pragma unroll
				for (uint32_t d = 1; d <= (uint32_t)DR; d++) if (d <= nn) sacc += __expf(lt[d - 1] - mx);
				w = mx + __logf(sacc);
			}
			const float wmax = block_max<TF_THR>(w, scratch);
			float* Mn = Ms + ((t + 1) & 1) * Lq;     // M_{t+1}
			// exp(M_t - max) for the next step (buffer t & 1 was last touched in step t+1, left through its closing barrier behind the proxy fence)
			if (y == 0 && t > 0) bulk_matrix(Ms + (t & 1) * Lq, p.E + n * Lq, Lq, &mbar[t & 1]);
			tc05::mbar_wait(&mbar[(t + 1) & 1], ((T - 2 - t) >> 1) & 1);
			const float mmax = c_mmax;
			if (y < P) { ev[y] = __expf(w - wmax); av[y] = c_av; }
			__syncthreads();
			// E[q][yy] = exp(M_{t+1}[q][yy] - mmax) * ev[yy] in place;
			// xi_t[q][yy] = exp(alpha_t[q] + M_{t+1}[q][yy] + B_t[yy] - logZ) = alpha^_t[q] E[q][yy] exp(rho_t + mmax + kref + wmax - logZ)
			// (a segment boundary after frame t has probability <= 1: the posteriors are NOT renormalised per frame)
			float rs = 0.0f;
			{
				float* mrow = Mn + rq * P;
				uint32_t cc = r_first;
#pragma unroll 4
				for (uint32_t k = 0; k < r_cnt; k++) {
					const float e = mrow[cc] * ev[cc]; mrow[cc] = e; rs += e;
					cc += TPR; if (cc >= P) cc -= P;
				}
			}
			for (uint32_t o = 1; o < TPR; o <<= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
			if (row_ok && rpart == 0) rsum[rq] = rs;
			// maximum of the row sums: one barrier, which also publishes the scaled matrix and the row sums (scratch[8..15] is only
			// rewritten behind the step's closing barrier)
			float bmax = row_ok ? rs : 0.0f;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
			if ((threadIdx.x & 31) == 0) scratch[8 + (threadIdx.x >> 5)] = bmax;
			__syncthreads();
			bmax = scratch[8];
#pragma unroll
			for (int w = 1; w < TF_THR / 32; w++) bmax = fmaxf(bmax, scratch[8 + w]);
			uint32_t q = q_first, yy = y_first;
			const float xscale = __expf((float)(c_rho + (double)mmax + kref + (double)wmax - lz));
			uint32_t nl = lab != LAB_BAD ? c_nl : LAB_BAD;
			const uint32_t lq = lab != LAB_BAD ? lab % P : LAB_BAD;
			if (nl != LAB_BAD && p.tidx[lq * P + nl] == 0xffffffffu) nl = LAB_BAD;       // a reference pair the N-state map does not have
			float* xrow = p.Xd + (n + 1) * p.Lq;         // stored with the frame whose duration-1 window carries the transition features
			q = q_first; yy = y_first;
			for (uint32_t i = y; i < P * P; i += TF_THR) {
				xrow[i] = ((q == lq && yy == nl) ? 1.0f : 0.0f) - av[q] * Mn[i] * xscale;
				yy += dy_; q += dq_;
				if (yy >= P) { yy -= P; q++; }
			}
			lb = y < P ? __logf(rsum[y] / bmax) : 0.0f;
			kappa = kref + (double)mmax + (double)wmax + (double)__logf(bmax);
		} else if (T > 1) {
			// tail frame of a multi-frame utterance: nothing to wait for, M_{T-1} stays in flight for the next step
		}
		if (y < P) lbh[(t & (ND_RING - 1)) * P + y] = lb;
		if (y == 0) kring[t & (ND_RING - 1)] = kappa;
		if (t == 0) for (uint32_t i = y; i < P * P; i += TF_THR) p.Xd[n * p.Lq + i] = 0.0f;      // no transition enters the first frame
		// gamma_t[d,y] = exp(S_t[d,y] + A_{t-d}[y] + beta_t[y] - logZ) feeds nothing of this chain: beta_t[y] - kappa_t and kappa_t are
		// stored, and Dm = [ref] - gamma is the posterior pass of the native no_dur path behind this kernel (launch_nodur_post, Mmax = 0)
		if (y < P) p.LB[n * p.Pp + y] = lb;
		if (y == 0) p.kappa[n] = kappa;
#pragma unroll
		for (uint32_t d = 1; d <= (uint32_t)DR; d++) sv[d] = sn[d];
		lab = n_lab; c_nl = n_nl; c_rho = n_rho; c_mmax = n_mmax; c_av = n_av;
		tc05::fence_proxy_async_smem();      // this step's in-place scaling (generic stores) before the bulk copy that reuses the buffer
		__syncthreads();
	}
}

}  // namespace

// (two dense matrices of Lq floats; the admission rule keeps the odd-stride size of the cp.async version for even phone counts, so the
// documented limit -- 161 labels -- stays what the tests pin)
size_t nodur_tf_smem_bytes(uint32_t P) {
	const size_t Lq = ((size_t)P * P + 3) / 4 * 4, old = (size_t)P * (P | 1u);
	return sizeof(float) * (2 * (Lq > old ? Lq : old) + (size_t)(3 + ND_RING) * P + 16) + sizeof(double) * ND_RING + 16;
}

cudaError_t launch_nodur_tf_dp(bool backward, const NodurTfParams& p, cudaStream_t s) {
	if (!p.n_utt) return cudaSuccess;
	const size_t smem = nodur_tf_smem_bytes(p.P);
	if (p.P > (uint32_t)TF_MAX_THR) return cudaErrorInvalidValue;
	auto go = [&](auto kern, int thr) -> cudaError_t {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return e;
		kern<<<p.n_utt, thr, smem, s>>>(p);
		return cudaGetLastError();
	};
	if (!backward) nodur_tf_numer_kernel<<<(p.n_utt + 3) / 4, 128, 0, s>>>(p);      // the numerators: a pass of their own beside the forward chain
	// instantiation whose unrolled duration loops cover max_dur (the D score terms of a frame wait in registers)
	auto pick = [&](auto dr) -> cudaError_t {
		constexpr int DR = decltype(dr)::value;
		// (256 threads measured slower here: recursion 1.47 -> 1.79 ms forward, 2.78 -> 3.03 ms backward at the recipe's shape)
		if (p.P <= 128) return backward ? go(nodur_tf_backward_kernel<128, DR>, 128) : go(nodur_tf_forward_kernel<128, DR>, 128);
		return backward ? go(nodur_tf_backward_kernel<TF_MAX_THR, DR>, TF_MAX_THR) : go(nodur_tf_forward_kernel<TF_MAX_THR, DR>, TF_MAX_THR);
	};
	if (p.D <= 4) return pick(std::integral_constant<int, 4>{});
	if (p.D <= 10) return pick(std::integral_constant<int, 10>{});
	if (p.D <= 16) return pick(std::integral_constant<int, 16>{});
	return pick(std::integral_constant<int, 31>{});
}

}  // namespace crfgpu
