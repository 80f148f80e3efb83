// Frame-level CRFs with at most 64 labels (cfg2: 61 phones, one state): forward and backward recursions with ONE WARP PER UTTERANCE.
//
// The lattice kernels built for the segmental models (crf_dp_tc.cu) advance 16 utterances per cluster through tensor-core steps of
// about 3.5 k cycles; for a 61 x 61 transition matrix that is all latency.  Here the matrix lives in REGISTERS -- lane l owns labels
// l and l + 32 and keeps their two columns of E = exp(M - Mmax) (forward) or rows (backward; the columns of E^T) -- the vector of
// the previous frame is broadcast through 256 bytes of shared memory per warp, and a frame is 128 FMAs per lane plus one exp:
// a few hundred cycles.  The recursions are those of CRF_StdStateNode::computeAlpha / computeBeta / computeExpF
// (CRF/src/nodes/CRF_StdStateNode.cpp:81-110,140-173,221-277; N-state models: CRF_StdNStateNode with E = 0 on the illegal pairs) in
// the probability domain, and the outputs are the ones the dense path's GEMMs and the alpha/beta dump consume (DpParams):
//   A_t[c], m_t         alpha_t[c] = m_t + log A_t[c]
//   logZ                m_{T-1} + log sum_c A_{T-1}[c]
//   Dm_t[c]             [c == reference label] - gamma_t[c],  gamma_t[c] = A_t[c] u_t[c] exp(m_t + bbase_t - logZ)
//   R_t[c]              exp(S_t[c] + beta_t[c] + m_{t-1} + Mmax - logZ), 0 on the first frame: xi_t(q,c) = A_{t-1}[q] E[q][c] R_t[c]
//   bbase_t, Uvec       beta_t[q] = bbase_t + log u_t[q]
// Scaling: every frame's vector is multiplied by a power of two taken from the maximum of the PREVIOUS frame's vector (its warp
// reduction runs beside the matrix product instead of behind it), so the log scales are exact sums of Mmax, the frame's score
// maximum and multiples of ln 2 in double.
#include "crf_kernels.cuh"

#include <cuda_runtime.h>
#include <cfloat>

namespace crfgpu {

namespace {

constexpr int FW = 4;                    // warps (utterances) per CTA
constexpr double LN2 = 0.693147180559945309417232121458;

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
// 2^-e with e = the binary exponent of mx (mx > 0, finite); e itself through *e_out
__device__ __forceinline__ float pow2_scale(float mx, int* e_out) {
	int e = ((__float_as_int(mx) >> 23) & 0xff) - 127;
	if (!(mx > 0.0f) || e < -100 || e > 100) e = 0;      // degenerate frame (no path) or denormal: leave the scale alone
	*e_out = e;
	return __int_as_float((127 - e) << 23);
}

// out[j] = sum_k vec[k] * Mreg[j][k]: the vector is broadcast from shared memory four entries per load
__device__ __forceinline__ void matvec(const float (&Mreg)[2][64], const float* vec, float (&out)[2]) {
	float acc[2][4] = {{0.0f, 0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 0.0f}};
#pragma unroll
	for (int k4 = 0; k4 < 16; k4++) {
		const float4 v = reinterpret_cast<const float4*>(vec)[k4];
		const int k = 4 * k4;
#pragma unroll
		for (int j = 0; j < 2; j++) {
			acc[j][0] = fmaf(v.x, Mreg[j][k + 0], acc[j][0]);
			acc[j][1] = fmaf(v.y, Mreg[j][k + 1], acc[j][1]);
			acc[j][2] = fmaf(v.z, Mreg[j][k + 2], acc[j][2]);
			acc[j][3] = fmaf(v.w, Mreg[j][k + 3], acc[j][3]);
		}
	}
	out[0] = (acc[0][0] + acc[0][1]) + (acc[0][2] + acc[0][3]);
	out[1] = (acc[1][0] + acc[1][1]) + (acc[1][2] + acc[1][3]);
}

}  // namespace

// BWD = false: forward.  BWD = true: backward; FUSED = true also forms the posteriors Dm / R of every frame (needs the forward pass
// to have finished), FUSED = false only runs the beta chain and stores u_t, bbase_t for frame_post_kernel -- the two chains of an
// utterance are independent given the scores, so they can then run side by side in one launch.
template <bool BWD, bool FUSED>
__device__ __forceinline__ void frame_dp_body(const DpParams& p, const uint32_t* __restrict__ utt_list, uint32_t n_utt, uint32_t cta, float (*vec_s)[64]) {
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t slot = cta * FW + warp;
	if (slot >= n_utt) return;                               // whole warps leave: no CTA-wide barrier below
	const uint32_t utt = utt_list[slot];
	const uint32_t L = p.L, Lp = p.Lp;
	const uint32_t off = p.off[utt], T = p.off[utt + 1] - off;
	if (T == 0) { if (!BWD && lane == 0) p.logZ[utt] = -DBL_MAX; return; }
	const uint32_t c0 = lane, c1 = lane + 32;
	const bool ok0 = c0 < L, ok1 = c1 < L, pad0 = c0 < Lp, pad1 = c1 < Lp;
	float* vec = vec_s[warp];
	// my two columns of E (forward: out[c] = sum_q a[q] E[q][c]) or of E^T (backward: out[q] = sum_c E[q][c] w[c])
	float Mreg[2][64];
	{
		const float* Msrc = BWD ? p.ET : p.E;
#pragma unroll
		for (int k = 0; k < 64; k++) {
			Mreg[0][k] = ((uint32_t)k < L && ok0) ? __ldg(Msrc + (size_t)k * Lp + c0) : 0.0f;
			Mreg[1][k] = ((uint32_t)k < L && ok1) ? __ldg(Msrc + (size_t)k * Lp + c1) : 0.0f;
		}
	}
	auto load_s = [&](uint32_t t, float (&sv)[2]) {
		const float* row = p.S + (size_t)(off + t) * Lp;
		sv[0] = ok0 ? __ldg(row + c0) : -INFINITY;
		sv[1] = ok1 ? __ldg(row + c1) : -INFINITY;
	};

	if (!BWD) {
		// scores two frames ahead: one frame (~700 cycles) does not cover a miss to HBM
		float sv[2], sn[2] = {-INFINITY, -INFINITY}, sn2[2] = {-INFINITY, -INFINITY};
		load_s(0, sv);
		if (T > 1) load_s(1, sn);
		if (T > 2) load_s(2, sn2);
		// frame 0: alpha = S (computeFirstAlpha)
		float sref = wmax(fmaxf(sv[0], sv[1]));
		float a[2] = {__expf(sv[0] - sref), __expf(sv[1] - sref)};
		double m = (double)sref;
		{
			float* Ar = p.A + (size_t)off * Lp;
			if (pad0) Ar[c0] = a[0];
			if (pad1) Ar[c1] = a[1];
			if (lane == 0) p.m[off] = m;
		}
		for (uint32_t t = 1; t < T; t++) {
			sv[0] = sn[0]; sv[1] = sn[1]; sn[0] = sn2[0]; sn[1] = sn2[1];
			if (t + 2 < T) load_s(t + 2, sn2);
			vec[c0] = a[0]; vec[c1] = a[1];
			__syncwarp();
			const float mxprev = wmax(fmaxf(a[0], a[1]));          // beside the product, not behind it
			sref = wmax(fmaxf(sv[0], sv[1]));
			const float es0 = __expf(sv[0] - sref), es1 = __expf(sv[1] - sref);
			float g[2];
			matvec(Mreg, vec, g);
			__syncwarp();                                            // everybody has read vec before the next frame overwrites it
			int e; const float sc = pow2_scale(mxprev, &e);
			a[0] = g[0] * es0 * sc; a[1] = g[1] * es1 * sc;
			m += p.Mmax + (double)sref + (double)e * LN2;
			float* Ar = p.A + (size_t)(off + t) * Lp;
			if (pad0) Ar[c0] = a[0];
			if (pad1) Ar[c1] = a[1];
			if (lane == 0) p.m[off + t] = m;
		}
		const float zs = wsum(a[0] + a[1]);
		if (lane == 0) p.logZ[utt] = m + log((double)zs);
	} else if (!FUSED) {
		// the beta chain alone: w~_t = exp(S_t - sref) u_t 2^-e;  u_t and bbase_t are what the posteriors need
		float sv[2], sn[2] = {-INFINITY, -INFINITY}, sn2[2] = {-INFINITY, -INFINITY};
		load_s(T - 1, sv);
		if (T > 1) load_s(T - 2, sn);
		if (T > 2) load_s(T - 3, sn2);
		// last frame: beta = 0 (setTailBeta): u = 1, bbase = 0
		float sref = wmax(fmaxf(sv[0], sv[1]));
		float w[2] = {__expf(sv[0] - sref), __expf(sv[1] - sref)};
		double kap = (double)sref;
		{
			const size_t n = (size_t)off + T - 1;
			float* Ur = p.Uvec + n * Lp;
			if (pad0) Ur[c0] = ok0 ? 1.0f : 0.0f;
			if (pad1) Ur[c1] = ok1 ? 1.0f : 0.0f;
			if (lane == 0) { p.bbase[n] = 0.0; p.kappa[n] = kap; }
		}
		for (uint32_t t = T - 1; t-- > 0;) {
			sv[0] = sn[0]; sv[1] = sn[1]; sn[0] = sn2[0]; sn[1] = sn2[1];
			if (t >= 2) load_s(t - 2, sn2);
			vec[c0] = w[0]; vec[c1] = w[1];
			__syncwarp();
			const float mxprev = wmax(fmaxf(w[0], w[1]));
			sref = wmax(fmaxf(sv[0], sv[1]));
			const float es0 = __expf(sv[0] - sref), es1 = __expf(sv[1] - sref);
			float u[2];
			matvec(Mreg, vec, u);
			__syncwarp();
			int e; const float sc = pow2_scale(mxprev, &e);
			w[0] = u[0] * es0 * sc; w[1] = u[1] * es1 * sc;      // es = 0 on the padding labels
			const double bbase = kap + p.Mmax;
			kap = bbase + (double)sref + (double)e * LN2;
			const size_t n = (size_t)off + t;
			float* Ur = p.Uvec + n * Lp;
			if (pad0) Ur[c0] = u[0];                               // rows / columns of E beyond L are zero: u = 0 there
			if (pad1) Ur[c1] = u[1];
			if (lane == 0) { p.bbase[n] = bbase; p.kappa[n] = kap; }
		}
	} else {
		const double logZ = p.logZ[utt];
		// frame state prefetched one frame ahead: scores, alpha, its scales, the reference label
		float sv[2], av[2], sn[2] = {-INFINITY, -INFINITY}, an[2] = {0.0f, 0.0f};
		double mt, mp, mtn = 0.0, mpn = 0.0; uint32_t lab, labn = LAB_BAD;
		auto load_frame = [&](uint32_t t, float (&s2)[2], float (&a2)[2], double& m_t, double& m_prev, uint32_t& lb) {
			load_s(t, s2);
			const float* Ar = p.A + (size_t)(off + t) * Lp;
			a2[0] = ok0 ? Ar[c0] : 0.0f; a2[1] = ok1 ? Ar[c1] : 0.0f;
			m_t = p.m[off + t]; m_prev = t ? p.m[off + t - 1] : 0.0;
			lb = p.node_lab[off + t];
		};
		load_frame(T - 1, sv, av, mt, mp, lab);
		if (T > 1) load_frame(T - 2, sn, an, mtn, mpn, labn);
		float w[2] = {0.0f, 0.0f};      // w~_{t+1}
		double kap = 0.0;               // kappa_{t+1}
		for (uint32_t t = T; t-- > 0;) {
			const bool tail = t + 1 == T;
			const size_t n = (size_t)off + t;
			float u[2] = {1.0f, 1.0f};
			double bbase = 0.0;           // tail: beta = 0 (setTailBeta)
			float sc = 1.0f; int e = 0;
			const float sref = wmax(fmaxf(sv[0], sv[1]));
			const float es0 = __expf(sv[0] - sref), es1 = __expf(sv[1] - sref);
			if (!tail) {
				vec[c0] = w[0]; vec[c1] = w[1];
				__syncwarp();
				const float mxprev = wmax(fmaxf(w[0], w[1]));
				matvec(Mreg, vec, u);
				__syncwarp();
				sc = pow2_scale(mxprev, &e);
				bbase = kap + p.Mmax;
			}
			// the chain: w~_t = exp(S_t - sref) u_t 2^-e,  kappa_t = bbase_t + sref + e ln 2
			w[0] = ok0 ? u[0] * es0 * sc : 0.0f; w[1] = ok1 ? u[1] * es1 * sc : 0.0f;
			kap = bbase + (double)sref + (double)e * LN2;
			// posteriors of the frame (off the chain)
			const float gsc = expf((float)(mt + bbase - logZ));
			const float rsc = t ? expf((float)(kap + mp + p.Mmax - logZ)) : 0.0f;
			const float g0 = av[0] * u[0] * gsc, g1 = av[1] * u[1] * gsc;
			float* Dr = p.Dm + n * Lp; float* Rr = p.R + n * Lp;
			if (pad0) { Dr[c0] = ok0 ? ((lab == c0) ? 1.0f : 0.0f) - g0 : 0.0f; Rr[c0] = w[0] * rsc; }
			if (pad1) { Dr[c1] = ok1 ? ((lab == c1) ? 1.0f : 0.0f) - g1 : 0.0f; Rr[c1] = w[1] * rsc; }
			if (p.Uvec) { float* Ur = p.Uvec + n * Lp; if (pad0) Ur[c0] = ok0 ? u[0] : 0.0f; if (pad1) Ur[c1] = ok1 ? u[1] : 0.0f; }
			if (lane == 0) { p.bbase[n] = bbase; p.kappa[n] = kap; }
			if (p.mass) { const float gs = wsum((ok0 ? g0 : 0.0f) + (ok1 ? g1 : 0.0f)); if (lane == 0) p.mass[n] = (double)gs; }
			// next frame (t - 1)
			sv[0] = sn[0]; sv[1] = sn[1]; av[0] = an[0]; av[1] = an[1]; mt = mtn; mp = mpn; lab = labn;
			if (t >= 2) load_frame(t - 2, sn, an, mtn, mpn, labn);
		}
	}
}

template <bool BWD>
__global__ void __launch_bounds__(FW * 32) frame_dp_kernel(DpParams p, const uint32_t* __restrict__ utt_list, uint32_t n_utt) {
	__shared__ __align__(16) float vec_s[FW][64];
	frame_dp_body<BWD, true>(p, utt_list, n_utt, blockIdx.x, vec_s);
}
// both chains in one launch: even CTAs run the forward recursion, odd CTAs the beta chain of the same four utterances.  The launch
// asks for enough dynamic shared memory that ONE CTA is resident per SM: each of its four warps then has a scheduler to itself (a
// frame is ~350 issue slots of a ~700-cycle chain -- a second warp on the scheduler lengthens the chain), the utterance list is
// longest first, so the longest chains start first and the short ones fill in behind them.
__global__ void __launch_bounds__(FW * 32) frame_dp_pair_kernel(DpParams p, const uint32_t* __restrict__ utt_list, uint32_t n_utt) {
	__shared__ __align__(16) float vec_s[FW][64];
	if ((blockIdx.x & 1u) == 0) frame_dp_body<false, true>(p, utt_list, n_utt, blockIdx.x >> 1, vec_s);
	else frame_dp_body<true, false>(p, utt_list, n_utt, blockIdx.x >> 1, vec_s);
}
// posteriors of every frame from the two finished chains (no recursion: one thread per frame and label)
//   gamma_t[c] = A_t[c] u_t[c] exp(m_t + bbase_t - logZ),   R_t[c] = u_t[c] exp(S_t[c] + bbase_t + m_{t-1} + Mmax - logZ), 0 on the first frame
__global__ void __launch_bounds__(256) frame_post_kernel(DpParams p, const uint32_t* __restrict__ frame_t, const uint32_t* __restrict__ frame_utt, uint32_t N) {
	// one warp per frame: the frame's scalars (two exponentials' arguments) once per lane, columns lane and lane + 32
	const uint64_t n = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
	const uint32_t lane = threadIdx.x & 31, Lp = p.Lp;
	if (n >= N) return;
	const double lz = p.logZ[frame_utt[n]], bb = p.bbase[n];
	const float gsc = expf((float)(p.m[n] + bb - lz));
	const bool first = frame_t[n] == 0;
	const float rarg = first ? 0.0f : (float)(bb + p.m[n - 1] + p.Mmax - lz);
	const uint32_t lab = p.node_lab[n];
	const size_t row = (size_t)n * Lp;
	for (uint32_t c = lane; c < Lp; c += 32) {
		float dm = 0.0f, r = 0.0f;
		if (c < p.L) {
			const float u = p.Uvec[row + c];
			dm = ((lab == c) ? 1.0f : 0.0f) - p.A[row + c] * u * gsc;
			if (!first) r = u * expf(__ldg(p.S + row + c) + rarg);
		}
		p.Dm[row + c] = dm; p.R[row + c] = r;
	}
}

cudaError_t launch_frame_dp(bool backward, const DpParams& p, const uint32_t* utt_list, uint32_t n_utt, cudaStream_t s) {
	if (!n_utt) return cudaSuccess;
	const unsigned grid = (n_utt + FW - 1) / FW;
	if (backward) frame_dp_kernel<true><<<grid, FW * 32, 0, s>>>(p, utt_list, n_utt);
	else frame_dp_kernel<false><<<grid, FW * 32, 0, s>>>(p, utt_list, n_utt);
	return cudaGetLastError();
}
cudaError_t launch_frame_dp_pair(const DpParams& p, const uint32_t* utt_list, uint32_t n_utt, cudaStream_t s) {
	if (!n_utt) return cudaSuccess;
	const unsigned G = (n_utt + FW - 1) / FW;
	// up to two CTAs per SM the launch is as long as its longest chain: ask for more than half of an SM's shared memory so that every
	// CTA has its SM to itself; larger batches are a matter of throughput and share the SMs
	int dev = 0, sms = 0;
	cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int excl = (sms > 0 && 2 * G <= 2u * (unsigned)sms) ? 120 * 1024 : 0;
	cudaError_t e = cudaFuncSetAttribute(frame_dp_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
	if (e != cudaSuccess) return e;
	frame_dp_pair_kernel<<<2 * G, FW * 32, excl, s>>>(p, utt_list, n_utt);
	return cudaGetLastError();
}
cudaError_t launch_frame_post(const DpParams& p, const uint32_t* frame_t, const uint32_t* frame_utt, uint32_t N, cudaStream_t s) {
	if (!N) return cudaSuccess;
	frame_post_kernel<<<(unsigned)(((uint64_t)N * 32 + 255) / 256), 256, 0, s>>>(p, frame_t, frame_utt, N);
	return cudaGetLastError();
}

}  // namespace crfgpu
