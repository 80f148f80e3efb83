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
// next frame's matrix prefetched with cp.async while the current one is used; the posteriors leave as Dm = [ref] - gamma and
// Xd = [ref pair] - xi, so that the state and transition gradients (and their empirical counts) are two more tensor-core GEMMs
// (launch_reduce_gemm_tc).  gamma and xi are normalised by their own per-frame sums, which are 1 in exact arithmetic.
#include <cfloat>

#include "crf_kernels.cuh"

namespace crfgpu {

namespace {

constexpr int TF_THR = 128;

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float block_max(float v, float* scratch) {
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	__syncthreads();
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	float r = scratch[0];
	for (int w = 1; w < TF_THR / 32; w++) r = fmaxf(r, scratch[w]);
	return r;
}
__device__ __forceinline__ float block_sum(float v, float* scratch) {
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	float r = 0.0f;
	for (int w = 0; w < TF_THR / 32; w++) r += scratch[w];
	return r;
}

// the L x L scores of frame n into a shared-memory matrix with odd row stride Ls (columns AND rows conflict-free)
__device__ __forceinline__ void prefetch_matrix(float* dst, const float* src, uint32_t L, uint32_t Ls) {
	for (uint32_t i = threadIdx.x; i < L * L; i += TF_THR) { const uint32_t p = i / L, c = i - p * L; cp_async4(dst + p * Ls + c, src + i); }
	cp_async_commit();
}

__global__ void __launch_bounds__(TF_THR) transftr_forward_kernel(TransFtrParams p) {
	extern __shared__ __align__(16) float sm[];
	const uint32_t L = p.L, Ls = L | 1u;
	float* Ms = sm;                              // [2][L][Ls]
	float* a_prev = sm + 2 * L * Ls;             // [L]
	float* scratch = a_prev + L;                 // [8]
	const uint32_t u = blockIdx.x, off = p.off[u], T = p.off[u + 1] - off, c = threadIdx.x;
	double rho = 0.0, num = 0.0;
	if (T > 1) prefetch_matrix(Ms + L * Ls, p.M + (size_t)(off + 1) * p.Lq, L, Ls);       // frame 1 -> buffer 1
	for (uint32_t t = 0; t < T; t++) {
		const size_t n = (size_t)off + t;
		float* Mt = Ms + (t & 1) * L * Ls;
		float w = -INFINITY;
		const float s = c < L ? p.S[n * p.Lp + c] : 0.0f;
		float mmax = 0.0f;
		if (t == 0) { if (c < L) w = s; }
		else {
			cp_async_wait_all();
			__syncthreads();
			if (t + 1 < T) prefetch_matrix(Ms + ((t + 1) & 1) * L * Ls, p.M + (n + 1) * p.Lq, L, Ls);
			float m = -INFINITY;
			for (uint32_t i = c; i < L * L; i += TF_THR) m = fmaxf(m, Mt[(i / L) * Ls + i % L]);
			mmax = block_max(m, scratch);
			if (c < L) {
				float v = 0.0f;
				for (uint32_t q = 0; q < L; q++) v = fmaf(a_prev[q], __expf(Mt[q * Ls + c] - mmax), v);
				w = __logf(v) + s;
			}
		}
		const float wmax = block_max(w, scratch);
		const float a = c < L ? __expf(w - wmax) : 0.0f;
		const float asum = block_sum(a, scratch);
		rho += (double)mmax + (double)wmax + (double)__logf(asum);
		if (c < L) { a_prev[c] = a / asum; p.A[n * p.Lp + c] = a / asum; }
		if (c == 0) {
			p.rho[n] = rho;
			// numerator on the reference path: state score of the frame's label + transition score from the previous frame's label
			const uint32_t y = p.labs[n];
			if (y < L) {
				num += (double)p.S[n * p.Lp + y];
				if (t > 0) { const uint32_t yp = p.labs[n - 1]; if (yp < L) num += (double)Mt[yp * Ls + y]; }
			}
		}
		__syncthreads();
	}
	if (c == 0) { p.logZ[u] = rho; p.numer[u] = num; }       // sum_c a_{T-1}[c] = 1: alpha_{T-1} sums to exp(rho)
}

__global__ void __launch_bounds__(TF_THR) transftr_backward_kernel(TransFtrParams p) {
	extern __shared__ __align__(16) float sm[];
	const uint32_t L = p.L, Ls = L | 1u;
	float* Ms = sm;                              // [2][L][Ls]
	float* b = sm + 2 * L * Ls;                  // [L] beta_t, max-normalised
	float* wv = b + L;                           // [L] exp(S_t - smax) * b_t
	float* av = wv + L;                          // [L] alpha_{t-1}
	float* scratch = av + L;                     // [8]
	const uint32_t u = blockIdx.x, off = p.off[u], T = p.off[u + 1] - off, c = threadIdx.x;
	if (c < L) b[c] = 1.0f;                      // setTailBeta
	if (T > 1) prefetch_matrix(Ms + ((T - 1) & 1) * L * Ls, p.M + (size_t)(off + T - 1) * p.Lq, L, Ls);
	__syncthreads();
	for (uint32_t t = T; t-- > 0;) {
		const size_t n = (size_t)off + t;
		const uint32_t y = p.labs[n];
		// gamma_t = A_t * b_t / sum
		const float a = c < L ? p.A[n * p.Lp + c] : 0.0f;
		const float g = c < L ? a * b[c] : 0.0f;
		const float gsum = block_sum(g, scratch);
		if (c < L) p.Dm[n * p.Lp + c] = ((y == c) ? 1.0f : 0.0f) - g / gsum;
		float* xrow = p.Xd + n * p.Lq;
		if (t == 0) {
			for (uint32_t i = c; i < L * L; i += TF_THR) xrow[i] = 0.0f;      // no transition enters the first frame
			break;
		}
		float* Mt = Ms + (t & 1) * L * Ls;
		cp_async_wait_all();
		__syncthreads();
		if (t > 1) prefetch_matrix(Ms + ((t - 1) & 1) * L * Ls, p.M + (n - 1) * p.Lq, L, Ls);
		// E_t = exp(M_t - mmax) in place
		float m = -INFINITY;
		for (uint32_t i = c; i < L * L; i += TF_THR) m = fmaxf(m, Mt[(i / L) * Ls + i % L]);
		const float mmax = block_max(m, scratch);
		const float s = c < L ? p.S[n * p.Lp + c] : -INFINITY;
		const float smax = block_max(s, scratch);
		if (c < L) { wv[c] = __expf(s - smax) * b[c]; av[c] = p.A[(n - 1) * p.Lp + c]; }
		__syncthreads();
		float part = 0.0f;
		for (uint32_t i = c; i < L * L; i += TF_THR) {
			const uint32_t q = i / L, cc = i - q * L;
			const float e = __expf(Mt[q * Ls + cc] - mmax) * wv[cc];          // E_t[q][cc] * w_t[cc]
			Mt[q * Ls + cc] = e;
			part += av[q] * e;
		}
		const float xsum = block_sum(part, scratch);                            // sum_{q,cc} alpha_{t-1}[q] E_t[q][cc] w_t[cc]
		const uint32_t yp = p.labs[n - 1];
		const float inv = 1.0f / xsum;
		for (uint32_t i = c; i < L * L; i += TF_THR) {
			const uint32_t q = i / L, cc = i - q * L;
			xrow[i] = ((q == yp && cc == y) ? 1.0f : 0.0f) - av[q] * Mt[q * Ls + cc] * inv;
		}
		// beta_{t-1}[q] = sum_cc E_t[q][cc] w_t[cc], max-normalised
		float bn = 0.0f;
		if (c < L) for (uint32_t cc = 0; cc < L; cc++) bn += Mt[c * Ls + cc];
		const float bmax = block_max(c < L ? bn : 0.0f, scratch);
		if (c < L) b[c] = bn / bmax;
		__syncthreads();
	}
}

}  // namespace

size_t transftr_smem_bytes(uint32_t L) { return sizeof(float) * ((size_t)2 * L * (L | 1u) + 3 * (size_t)L + 16); }

cudaError_t launch_transftr_dp(bool backward, const TransFtrParams& p, cudaStream_t s) {
	if (!p.n_utt) return cudaSuccess;
	const size_t smem = transftr_smem_bytes(p.L);
	cudaError_t e = cudaFuncSetAttribute(backward ? (const void*)transftr_backward_kernel : (const void*)transftr_forward_kernel,
	                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	if (backward) transftr_backward_kernel<<<p.n_utt, TF_THR, smem, s>>>(p);
	else transftr_forward_kernel<<<p.n_utt, TF_THR, smem, s>>>(p);
	return cudaGetLastError();
}

}  // namespace crfgpu
