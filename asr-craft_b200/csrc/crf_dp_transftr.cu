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

// CTA sizes: thread = label.  128 threads up to 128 labels, 192 beyond (the double-buffered L x L tile bounds the label count by shared memory:
// 169 labels frame-level, 160 segmental)
constexpr int TF_MAX_THR = 192;

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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

// the L x L scores of frame n into a shared-memory matrix with odd row stride Ls (columns AND rows conflict-free)
template <int TF_THR>
__device__ __forceinline__ void prefetch_matrix(float* dst, const float* src, uint32_t L, uint32_t Ls) {
	for (uint32_t i = threadIdx.x; i < L * L; i += TF_THR) { const uint32_t p = i / L, c = i - p * L; cp_async4(dst + p * Ls + c, src + i); }
	cp_async_commit();
}

template <int TF_THR>
__global__ void __launch_bounds__(TF_THR) transftr_forward_kernel(TransFtrParams p) {
	extern __shared__ __align__(16) float sm[];
	const uint32_t L = p.L, Ls = L | 1u;
	float* Ms = sm;                              // [2][L][Ls]
	float* a_prev = sm + 2 * L * Ls;             // [L]
	float* scratch = a_prev + L;                 // [8]
	const uint32_t u = blockIdx.x, off = p.off[u], T = p.off[u + 1] - off, c = threadIdx.x;
	double rho = 0.0, num = 0.0;
	if (T > 1) prefetch_matrix<TF_THR>(Ms + L * Ls, p.M + (size_t)(off + 1) * p.Lq, L, Ls);       // frame 1 -> buffer 1
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
			if (t + 1 < T) prefetch_matrix<TF_THR>(Ms + ((t + 1) & 1) * L * Ls, p.M + (n + 1) * p.Lq, L, Ls);
			float m = -INFINITY;
			for (uint32_t i = c; i < L * L; i += TF_THR) m = fmaxf(m, Mt[(i / L) * Ls + i % L]);
			mmax = block_max<TF_THR>(m, scratch);
			if (c < L) {
				float v = 0.0f;
				for (uint32_t q = 0; q < L; q++) v = fmaf(a_prev[q], __expf(Mt[q * Ls + c] - mmax), v);
				w = __logf(v) + s;
			}
		}
		const float wmax = block_max<TF_THR>(w, scratch);
		const float a = c < L ? __expf(w - wmax) : 0.0f;
		const float asum = block_sum<TF_THR>(a, scratch);
		rho += (double)mmax + (double)wmax + (double)__logf(asum);
		if (c < L) { a_prev[c] = a / asum; p.A[n * p.Lp + c] = a / asum; }
		if (c == 0) {
			p.rho[n] = rho;
			// numerator on the reference path: state score of the frame's label + transition score from the previous frame's label
			const uint32_t y = p.labs[n];
			if (y < L) {
				num += (double)p.S[n * p.Lp + y];
				if (t > 0) { const uint32_t yp = p.labs[n - 1]; if (yp < L && p.tidx[yp * L + y] != 0xffffffffu) num += (double)Mt[yp * Ls + y]; }
			}
		}
		__syncthreads();
	}
	if (c == 0) { p.logZ[u] = rho; p.numer[u] = num; }       // sum_c a_{T-1}[c] = 1: alpha_{T-1} sums to exp(rho)
}

template <int TF_THR>
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
	if (T > 1) prefetch_matrix<TF_THR>(Ms + ((T - 1) & 1) * L * Ls, p.M + (size_t)(off + T - 1) * p.Lq, L, Ls);
	__syncthreads();
	for (uint32_t t = T; t-- > 0;) {
		const size_t n = (size_t)off + t;
		const uint32_t y = p.labs[n];
		// gamma_t = A_t * b_t / sum
		const float a = c < L ? p.A[n * p.Lp + c] : 0.0f;
		const float g = c < L ? a * b[c] : 0.0f;
		const float gsum = block_sum<TF_THR>(g, scratch);
		if (c < L) p.Dm[n * p.Lp + c] = ((y == c) ? 1.0f : 0.0f) - g / gsum;
		float* xrow = p.Xd + n * p.Lq;
		if (t == 0) {
			for (uint32_t i = c; i < L * L; i += TF_THR) xrow[i] = 0.0f;      // no transition enters the first frame
			break;
		}
		float* Mt = Ms + (t & 1) * L * Ls;
		cp_async_wait_all();
		__syncthreads();
		if (t > 1) prefetch_matrix<TF_THR>(Ms + ((t - 1) & 1) * L * Ls, p.M + (n - 1) * p.Lq, L, Ls);
		// E_t = exp(M_t - mmax) in place
		float m = -INFINITY;
		for (uint32_t i = c; i < L * L; i += TF_THR) m = fmaxf(m, Mt[(i / L) * Ls + i % L]);
		const float mmax = block_max<TF_THR>(m, scratch);
		const float s = c < L ? p.S[n * p.Lp + c] : -INFINITY;
		const float smax = block_max<TF_THR>(s, scratch);
		if (c < L) { wv[c] = __expf(s - smax) * b[c]; av[c] = p.A[(n - 1) * p.Lp + c]; }
		__syncthreads();
		float part = 0.0f;
		for (uint32_t i = c; i < L * L; i += TF_THR) {
			const uint32_t q = i / L, cc = i - q * L;
			const float e = __expf(Mt[q * Ls + cc] - mmax) * wv[cc];          // E_t[q][cc] * w_t[cc]
			Mt[q * Ls + cc] = e;
			part += av[q] * e;
		}
		const float xsum = block_sum<TF_THR>(part, scratch);                            // sum_{q,cc} alpha_{t-1}[q] E_t[q][cc] w_t[cc]
		uint32_t yp = p.labs[n - 1];
		if (yp < L && y < L && p.tidx[yp * L + y] == 0xffffffffu) yp = LAB_BAD;      // a reference pair the N-state map does not have
		const float inv = 1.0f / xsum;
		for (uint32_t i = c; i < L * L; i += TF_THR) {
			const uint32_t q = i / L, cc = i - q * L;
			xrow[i] = ((q == yp && cc == y) ? 1.0f : 0.0f) - av[q] * Mt[q * Ls + cc] * inv;
		}
		// beta_{t-1}[q] = sum_cc E_t[q][cc] w_t[cc], max-normalised
		float bn = 0.0f;
		if (c < L) for (uint32_t cc = 0; cc < L; cc++) bn += Mt[c * Ls + cc];
		const float bmax = block_max<TF_THR>(c < L ? bn : 0.0f, scratch);
		if (c < L) b[c] = bn / bmax;
		__syncthreads();
	}
}

}  // namespace

size_t transftr_smem_bytes(uint32_t L) { return sizeof(float) * ((size_t)2 * L * (L | 1u) + 3 * (size_t)L + 16); }

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
	if (p.L <= 128) return backward ? go(transftr_backward_kernel<128>, 128) : go(transftr_forward_kernel<128>, 128);
	return backward ? go(transftr_backward_kernel<TF_MAX_THR>, TF_MAX_THR) : go(transftr_forward_kernel<TF_MAX_THR>, TF_MAX_THR);
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
// One CTA per utterance, thread = phone (<= 128), M_{t+1} prefetched with cp.async, alpha_t normalised to sum 1 with a running
// log scale rho_t, the D-term sums formed as float log-sum-exps of differences of the (double) scales.
// =================================================================================================
namespace {

constexpr uint32_t ND_RING = 32;     // max_dur <= 31

template <int TF_THR>
__global__ void __launch_bounds__(TF_THR) nodur_tf_forward_kernel(NodurTfParams p) {
	extern __shared__ __align__(16) float sm[];
	const uint32_t P = p.P, D = p.D, Ps = P | 1u;
	double* rring = reinterpret_cast<double*>(sm);            // [ND_RING] rho_t (first: 8-byte aligned)
	float* Ms = sm + 2 * ND_RING;                // [2][P][Ps]
	float* av = Ms + 2 * P * Ps;                 // [P] alpha_t (sum 1)
	float* lgh = av + P;                         // [ND_RING][P] log A_t[y] - rho_t of the last D frames
	float* scratch = lgh + ND_RING * P;          // [8]
	const uint32_t u = blockIdx.x, off = p.off[u], T = p.off[u + 1] - off, y = threadIdx.x;
	double num = 0.0;
	if (T > 1) prefetch_matrix<TF_THR>(Ms + P * Ps, p.M + (size_t)(off + 1) * p.Lq, P, Ps);       // M_1 -> buffer 1
	for (uint32_t t = 0; t < T; t++) {
		const size_t n = (size_t)off + t;
		const uint32_t dmax = min(t + 1, D);
		const double rref = t > 0 ? rring[(t - 1) & (ND_RING - 1)] : 0.0;
		// w[y] = log sum_d exp(S_t[d,y] + A_{t-d}[y] - rref)
		float w = -INFINITY;
		if (y < P) {
			float lt[ND_RING];
			float mx = -INFINITY;
			for (uint32_t d = 1; d <= dmax; d++) {
				float v = p.S[n * p.Lp + (size_t)(d - 1) * P + y];
				if (d <= t) v += lgh[((t - d) & (ND_RING - 1)) * P + y] + (float)(rring[(t - d) & (ND_RING - 1)] - rref);
				else v += (float)(-rref);
				lt[d - 1] = v; mx = fmaxf(mx, v);
			}
			float sacc = 0.0f;
			for (uint32_t d = 1; d <= dmax; d++) sacc += __expf(lt[d - 1] - mx);
			w = mx + __logf(sacc);
		}
		const float wmax = block_max<TF_THR>(w, scratch);
		const float a = y < P ? __expf(w - wmax) : 0.0f;
		const float asum = block_sum<TF_THR>(a, scratch);
		const double rho = rref + (double)wmax + (double)__logf(asum);
		if (y < P) { av[y] = a / asum; p.A[n * p.Pp + y] = a / asum; }
		if (y == 0) { rring[t & (ND_RING - 1)] = rho; p.rho[n] = rho; }
		// numerator: state score of the reference segment ending here
		const uint32_t lab = p.node_lab[n];
		if (y == 0 && lab != LAB_BAD) num += (double)p.S[n * p.Lp + lab];
		__syncthreads();
		if (t + 1 < T) {
			// A_t[y] = rho_t + lgh_t[y],  lgh_t[y] = mmax + log sum_q alpha_t[q] exp(M_{t+1}[q][y] - mmax)
			float* Mn = Ms + ((t + 1) & 1) * P * Ps;
			cp_async_wait_all();
			__syncthreads();
			if (t + 2 < T) prefetch_matrix<TF_THR>(Ms + (t & 1) * P * Ps, p.M + (n + 2) * p.Lq, P, Ps);
			float m = -INFINITY;
			for (uint32_t i = y; i < P * P; i += TF_THR) m = fmaxf(m, Mn[(i / P) * Ps + i % P]);
			const float mmax = block_max<TF_THR>(m, scratch);
			if (y < P) {
				float v = 0.0f;
				for (uint32_t q = 0; q < P; q++) v = fmaf(av[q], __expf(Mn[q * Ps + y] - mmax), v);
				const float l = mmax + __logf(v);
				lgh[(t & (ND_RING - 1)) * P + y] = l; p.LG[n * p.Pp + y] = l;
			}
			if (y == 0 && lab != LAB_BAD) { const uint32_t nl = p.next_lab[n]; if (nl != LAB_BAD && p.tidx[(lab % P) * P + nl] != 0xffffffffu) num += (double)Mn[(lab % P) * Ps + nl]; }
			__syncthreads();
		}
	}
	if (y == 0) { p.logZ[u] = T ? rring[(T - 1) & (ND_RING - 1)] : 0.0; p.numer[u] = num; }
}

template <int TF_THR>
__global__ void __launch_bounds__(TF_THR) nodur_tf_backward_kernel(NodurTfParams p) {
	extern __shared__ __align__(16) float sm[];
	const uint32_t P = p.P, D = p.D, Ps = P | 1u;
	double* kring = reinterpret_cast<double*>(sm);            // [ND_RING] kappa_t (first: 8-byte aligned)
	float* Ms = sm + 2 * ND_RING;                // [2][P][Ps]
	float* av = Ms + 2 * P * Ps;                 // [P] alpha_t
	float* lbh = av + P;                         // [ND_RING][P] beta_t[y] - kappa_t of the last D frames
	float* ev = lbh + ND_RING * P;               // [P] exp(B_t[y] - its maximum)
	float* scratch = ev + P;                     // [8]
	const uint32_t u = blockIdx.x, off = p.off[u], T = p.off[u + 1] - off, y = threadIdx.x;
	const double lz = p.logZ[u];
	if (T > 1) prefetch_matrix<TF_THR>(Ms + ((T - 1) & 1) * P * Ps, p.M + (size_t)(off + T - 1) * p.Lq, P, Ps);    // M_{T-1}
	for (uint32_t t = T; t-- > 0;) {
		const size_t n = (size_t)off + t;
		const uint32_t nn = min(T - 1 - t, D), dmax = min(t + 1, D);
		const uint32_t lab = p.node_lab[n];
		float lb = 0.0f;                         // beta_t[y] - kappa_t; tail: beta = 0
		double kappa = 0.0;
		if (nn > 0) {
			// w[y] = B_t[y] - kref = log sum_d exp(S_{t+d}[d,y] + beta_{t+d}[y] - kref)
			const double kref = kring[(t + 1) & (ND_RING - 1)];
			float w = -INFINITY;
			if (y < P) {
				float lt[ND_RING];
				float mx = -INFINITY;
				for (uint32_t d = 1; d <= nn; d++) {
					const float v = p.S[(n + d) * p.Lp + (size_t)(d - 1) * P + y] + lbh[((t + d) & (ND_RING - 1)) * P + y] + (float)(kring[(t + d) & (ND_RING - 1)] - kref);
					lt[d - 1] = v; mx = fmaxf(mx, v);
				}
				float sacc = 0.0f;
				for (uint32_t d = 1; d <= nn; d++) sacc += __expf(lt[d - 1] - mx);
				w = mx + __logf(sacc);
			}
			const float wmax = block_max<TF_THR>(w, scratch);
			float* Mn = Ms + ((t + 1) & 1) * P * Ps;     // M_{t+1}
			cp_async_wait_all();
			__syncthreads();
			if (t > 0) prefetch_matrix<TF_THR>(Ms + (t & 1) * P * Ps, p.M + n * p.Lq, P, Ps);      // M_t for the next step
			float m = -INFINITY;
			for (uint32_t i = y; i < P * P; i += TF_THR) m = fmaxf(m, Mn[(i / P) * Ps + i % P]);
			const float mmax = block_max<TF_THR>(m, scratch);
			if (y < P) { ev[y] = __expf(w - wmax); av[y] = p.A[n * p.Pp + y]; }
			__syncthreads();
			// E[q][yy] = exp(M_{t+1}[q][yy] - mmax) * ev[yy] in place;
			// xi_t[q][yy] = exp(alpha_t[q] + M_{t+1}[q][yy] + B_t[yy] - logZ) = alpha^_t[q] E[q][yy] exp(rho_t + mmax + kref + wmax - logZ)
			// (a segment boundary after frame t has probability <= 1: the posteriors are NOT renormalised per frame)
			for (uint32_t i = y; i < P * P; i += TF_THR) {
				const uint32_t q = i / P, yy = i - q * P;
				Mn[q * Ps + yy] = __expf(Mn[q * Ps + yy] - mmax) * ev[yy];
			}
			__syncthreads();
			const float xscale = __expf((float)(p.rho[n] + (double)mmax + kref + (double)wmax - lz));
			uint32_t nl = lab != LAB_BAD ? p.next_lab[n] : LAB_BAD;
			const uint32_t lq = lab != LAB_BAD ? lab % P : LAB_BAD;
			if (nl != LAB_BAD && p.tidx[lq * P + nl] == 0xffffffffu) nl = LAB_BAD;       // a reference pair the N-state map does not have
			float* xrow = p.Xd + (n + 1) * p.Lq;         // stored with the frame whose duration-1 window carries the transition features
			for (uint32_t i = y; i < P * P; i += TF_THR) {
				const uint32_t q = i / P, yy = i - q * P;
				xrow[i] = ((q == lq && yy == nl) ? 1.0f : 0.0f) - av[q] * Mn[q * Ps + yy] * xscale;
			}
			float bn = 0.0f;
			if (y < P) for (uint32_t yy = 0; yy < P; yy++) bn += Mn[y * Ps + yy];
			const float bmax = block_max<TF_THR>(y < P ? bn : 0.0f, scratch);
			lb = y < P ? __logf(bn / bmax) : 0.0f;
			kappa = kref + (double)mmax + (double)wmax + (double)__logf(bmax);
		} else if (T > 1) {
			// tail frame of a multi-frame utterance: nothing to wait for, M_{T-1} stays in flight for the next step
		}
		if (y < P) lbh[(t & (ND_RING - 1)) * P + y] = lb;
		if (y == 0) kring[t & (ND_RING - 1)] = kappa;
		if (t == 0) for (uint32_t i = y; i < P * P; i += TF_THR) p.Xd[n * p.Lq + i] = 0.0f;      // no transition enters the first frame
		// gamma_t[d,y] = exp(S_t[d,y] + A_{t-d}[y] + beta_t[y] - logZ): the scalar part (rho_{t-d} + kappa_t - logZ) is formed in double
		if (y < P) {
			for (uint32_t d = 1; d <= D; d++) {
				const uint32_t col = (d - 1) * P + y;
				float dm = 0.0f;
				if (d <= dmax) {
					float v = p.S[n * p.Lp + col] + lb;
					if (d <= t) v += p.LG[(n - d) * p.Pp + y] + (float)(p.rho[n - d] + kappa - lz);
					else v += (float)(kappa - lz);
					dm = ((lab == col) ? 1.0f : 0.0f) - __expf(v);
				}
				p.Dm[n * p.Lp + col] = dm;
			}
		}
		__syncthreads();
	}
}

}  // namespace

size_t nodur_tf_smem_bytes(uint32_t P) { return sizeof(float) * ((size_t)2 * P * (P | 1u) + (size_t)(2 + ND_RING) * P + 16) + sizeof(double) * ND_RING + 16; }

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
	if (p.P <= 128) return backward ? go(nodur_tf_backward_kernel<128>, 128) : go(nodur_tf_forward_kernel<128>, 128);
	return backward ? go(nodur_tf_backward_kernel<TF_MAX_THR>, TF_MAX_THR) : go(nodur_tf_forward_kernel<TF_MAX_THR>, TF_MAX_THR);
}

}  // namespace crfgpu
