// Native lattice recursions of the segmental models WITHOUT duration labels (stdseg_no_dur, _no_transftr, _no_segtransftr; one
// state per phone, transition bias only) for sm_100a -- the O(P^2 + D*P) form that scales to P >= 1000 phones, D = 30.
//
// Reference semantics (log domain, fp64): CRF_StdSegStateNode_WithoutDurLab_WithoutSegTransFtr::computeAlpha / computeAlphaPlusTrans /
// computeBeta / computeExpF (CRF/src/nodes/CRF_StdSegStateNode_WithoutDurLab_WithoutSegTransFtr.cpp:123-333, 1077-1116, 395-614,
// 616-1066) driven by CRF_NewGradBuilder_StdSeg_NoDur_NoTrans::buildGradient
// (CRF/src/trainers/gradbuilders/CRF_NewGradBuilder_StdSeg_NoDur_NoTrans.cpp:65-492):
//   A_t[y]     = logsum_y' (alpha_t[y'] + M[y',y])              alpha_t[d,y] = S_t[d,y] + A_{t-d}[y]   (= S_t[d,y] when d == t+1)
//   alpha_t[y] = logsum_d alpha_t[d,y]                           logZ = logsum_y alpha_{T-1}[y]
//   B_t[y]     = logsum_d (S_{t+d}[d,y] + beta_{t+d}[y])         beta_t[y'] = logsum_y (M[y',y] + B_t[y]),  beta_{T-1} = 0
//   gamma_t[d,y] = exp(alpha_t[d,y] + beta_t[y] - logZ)          xi_t[y',y]  = exp(alpha_t[y'] + M[y',y] + B_t[y] - logZ)
//
// Device form (probability domain, one log scale per frame):  alpha_t[y] = rho_t + log a_t[y],  A_t[y] = rho_t + Mmax + LG_t[y] with
// exp(LG_t[y]) = sum_y' a_t[y'] E[y'][y], E = exp(M - Mmax);  B_t[y] = sigma_t + log bh_t[y],  beta_t[y'] = kappa_t + LB_t[y'] with
// exp(LB_t[y']) = sum_y E[y'][y] bh_t[y], kappa_t = sigma_t + Mmax.  The scales are upper bounds built from per-duration score maxima
// and the EXACT log-sums of earlier frames, so every stored value is <= D and nothing overflows.
//
// Work split: the P x P matrix is sliced over the CTAs of a GROUP -- CTA pt keeps the 32 columns (forward) / rows (backward) of its
// phone tile in shared memory for the whole launch, pre-split into bf16 hi / lo halves in mma.sync fragment order -- and UT = 16
// utterances advance in lock-step per group.  Per lock-step:
//   phase A  (element-wise, D terms per (utterance, phone)): the new vector a_t / bh_t of the CTA's own phones from score terms that
//            wait in registers (requested a step ahead, their lines pulled into L2 a phase before that), history terms in a
//            shared-memory ring and float scale terms formed once per (utterance, duration); written to the lattice array and, in the
//            A-fragment order of the product, to the group's exchange buffer together with its partial sums;
//   ONE barrier among the CTAs of the group (a monotonic counter in global memory; the launch is cooperative, so they are co-resident);
//   phase B  the CTA's 16 x 32 block of the matrix product from the exchanged vector on the tensor cores (mma.sync m16n8k16, bf16
//            hi/lo splits, 8 warps = 8 interleaved slices of the contraction index, one shared-memory reduction), LG / LB, and the
//            scale of the next frame -- every CTA of the group derives the same scale from the same exchanged partial sums, so no
//            second barrier is needed; the per-utterance bookkeeping is done by the warp that owns the utterance, lane = duration.
//   The posteriors Dm[n][(d,y)] = [reference segment] - gamma_t[d,y] for the state-gradient GEMM are a pass of their own behind the
//   backward recursion (nodur_post_kernel); R[n+1][y] (phase A) feeds the Xi GEMM; both GEMMs run on the TMA-fed kernels of
//   crf_tma_gemm.cu.
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include <cfloat>

#include "crf_kernels.cuh"

namespace crfgpu {

namespace {

constexpr int UT = NODUR_UT, PT = 32, NTHR = 256, NW = NTHR / 32, RING = 32;
constexpr uint32_t DC = 10;         // durations whose loads are in flight together
static_assert(UT == 16 && NW * 2 == UT, "thread mapping: warp w owns utterances 2w, 2w+1 of the batch");

struct Shared {
	float red[NW][UT][PT];        // partial products of the 8 contraction slices
	float tileF[UT * PT];         // the CTA's slice of the new vector in the A-fragment order of the product (frag_index)
	double g_ring[UT][RING];      // forward: exact log-sum of alpha_t | backward: upper bound of beta_t
	double s_ring[UT][RING];      // forward: rho_t                    | backward: kappa_t
	double scale[UT];             // scale of the frame being produced (rho_t | sigma_t)
	double lz[UT];                // backward: logZ of the utterance
	float scf[UT][RING];          // per-step float scale of duration d: the fp64 differences are formed once per (utterance, duration), not per entry
	uint32_t utt[UT], off[UT], len[UT];
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

// (x, y) -> packed bf16 pairs: hi = round-to-nearest halves, lo = the halves of the remainders (x in the low 16 bits)
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
	const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
	const __nv_bfloat162 l = __floats2bfloat162_rn(x - __low2float(h), y - __high2float(h));
	hi = *reinterpret_cast<const uint32_t*>(&h); lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
	asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
	             : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Exchange layout = A-fragment order of mma.sync m16n8k16 (rows = utterances, k = phone): element (u, k) of k-step ks = k / 16 sits at
// float ((ks * 4 + e) * 32 + lane) * 2 + (k & 1) with e = u / 8 + 2 * ((k % 16) / 8), lane = (u % 8) * 4 + (k % 8) / 2 -- so the four
// float2 operand loads of a k-step are 256 contiguous bytes per warp, and a CTA's 32-phone tile (two k-steps) is one 2 KB block.
__device__ __forceinline__ uint32_t frag_index(uint32_t u, uint32_t j) {      // j = phone inside the 32-phone tile
	const uint32_t ks = j >> 4, kk = j & 15;
	return (((ks * 4 + (u >> 3) + 2 * (kk >> 3)) * 32 + (u & 7) * 4 + ((kk & 7) >> 1)) << 1) + (kk & 1);
}

// all CTAs of the group have published their slice of step `gstep`
__device__ __forceinline__ void group_barrier(uint32_t* ctr, uint32_t target) {
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		atomicAdd(ctr, 1u);
		while (ld_acquire(ctr) < target) {}
	}
	__syncthreads();
}

}  // namespace

size_t nodur_smem_bytes(uint32_t P) { return sizeof(Shared) + (size_t)((P + 31) / 32 * 32) * PT * sizeof(float) + sizeof(float) * RING * UT * PT + 32; }

#define NTICK(i) do { if (timing) { const long long now_ = clock64(); tacc[i] += (unsigned long long)(now_ - tlast); tlast = now_; } } while (0)

template <bool BWD, int DR>
__global__ void __launch_bounds__(NTHR, 1) nodur_dp_kernel(NodurParams p) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
	// B operand of the product, E[k][y0+j] (fwd) | E[y0+j][k] (bwd), as bf16 hi / lo halves in mma.sync fragment order:
	// Ef[(ks * 4 + nt) * 32 + lane] = {hi(k 2t,2t+1), hi(k 2t+8,2t+9), lo(..), lo(..)} of column nt * 8 + g, k relative to 16 ks
	uint4* Ef = reinterpret_cast<uint4*>(smem_raw + (sizeof(Shared) + 15) / 16 * 16);
	const uint32_t P = p.P, Pp = p.Pp, D = p.D, Pk = (P + 31) / 32 * 32, n_ks = Pk / 16;
	// the CTA's own slice of log(a_t E) (forward) / log(E bh_t) (backward) of the last RING frames: phase A reads its D history terms
	// here instead of the lattice arrays in global memory
	float* hist = reinterpret_cast<float*>(Ef + (size_t)n_ks * 128);                     // [RING][UT][PT]
	const size_t Lp = p.Lp;
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t g = blockIdx.x / p.npt, pt = blockIdx.x % p.npt, y0 = pt * PT, y = y0 + lane;
	const bool y_ok = y < P;
	const float* Msrc = BWD ? p.ET : p.E;
	const float* const Sg = p.S;
	for (uint32_t i = tid; i < n_ks * 4 * 32; i += NTHR) {
		const uint32_t ks = i >> 7, nt = (i >> 5) & 3, ln = i & 31, fg = ln >> 2, ft = ln & 3, j = nt * 8 + fg;
		float v[4];
#pragma unroll
		for (int e = 0; e < 4; e++) {
			const uint32_t k = ks * 16 + 2 * ft + (e & 1) + (e >> 1) * 8;
			v[e] = (k < P && y0 + j < P) ? Msrc[(size_t)k * Pp + y0 + j] : 0.0f;
		}
		uint4 w;
		split2(v[0], v[1], w.x, w.z); split2(v[2], v[3], w.y, w.w);
		Ef[i] = w;
	}
	float* xbase = p.xch + (size_t)g * 2 * ((size_t)Pk * UT + (size_t)p.npt * UT);
	const size_t xstride = (size_t)Pk * UT + (size_t)p.npt * UT;
	uint32_t* ctr = p.ctr + g;
	uint32_t gstep = 0;
	const bool timing = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
	unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	long long tlast = timing ? clock64() : 0;
	// Everything that belongs to ONE utterance of the batch -- scales, rings, the float scale terms of phase A -- is kept by the warp that
	// owns the utterance in phases A and B (warp w: utterances 2w, 2w + 1; lane = duration - 1 for the D-term maxima), so the
	// bookkeeping needs no block-wide barrier and no single-thread loops over D.
	const uint32_t u0 = warp * 2;
	const uint32_t s_step = BWD ? (uint32_t)Lp + P : P;      // forward: S[n][(d-1)P + y]; backward: S[n+d][(d-1)P + y]
	// values that do not depend on the recursion are requested one step ahead and wait in registers: the D score terms of both
	// entries of the thread, the score maxima that bound the next scale, the forward scale of the frame (backward: Xi factor)
	float sn[2][DR]; float sm[2]; double rh[2];
	auto prefetch = [&](uint32_t tt, uint32_t maxlen) {
#pragma unroll
		for (int i = 0; i < 2; i++) {
			const uint32_t u = u0 + i, len = sh.len[u];
			const bool in = tt < maxlen;                          // (tt wraps to 0xffffffff behind the last backward step)
			bool a; uint32_t lm;
			if (!BWD) { a = in && tt < len; lm = a ? min(tt + 1, D) : 0; }
			else { a = in && tt + 1 < len; lm = a ? min(len - 1 - tt, D) : 0; }
			const size_t n = (size_t)sh.off[u] + (a ? tt : 0);
			const float* q = Sg + (n + (BWD ? 1 : 0)) * Lp + (y_ok ? y : 0);     // the d = 1 term; one pointer increment per duration
			const uint32_t lmy = y_ok ? lm : 0;
#pragma unroll
			for (uint32_t d = 1; d <= (uint32_t)DR; d++) {
				sn[i][d - 1] = d <= lmy ? __ldg(q) : -INFINITY;      // a term that does not exist contributes exp(-inf) = 0
				q += s_step;
			}
			// lane = d - 1: forward smaxd[n][d-1] (d <= min(tt + 1, D)); backward smaxd[n + d][d - 1] (d <= min(len - 1 - tt, D))
			sm[i] = lane < lm ? __ldg(p.smaxd + (BWD ? (n + lane + 1) * D + lane : n * D + lane)) : 0.0f;
			rh[i] = (BWD && a) ? p.rho[n] : 0.0;
		}
	};
	// ... and one step before that their lines are pulled from DRAM into L2 (lane = d - 1 asks for the 128-byte line of duration d): all
	// CTAs request their 60 KB of scores at the same point of the lock-step, 8 MB in a burst that ran at the HBM rate for 4 k cycles
	// when the register loads met DRAM themselves; requested here, the transfer hides behind the barrier and the product
	auto prefetch_l2 = [&](uint32_t tt, uint32_t maxlen) {
#pragma unroll
		for (int i = 0; i < 2; i++) {
			const uint32_t u = u0 + i, len = sh.len[u];
			const bool in = tt < maxlen;
			uint32_t lm;
			if (!BWD) lm = (in && tt < len) ? min(tt + 1, D) : 0;
			else lm = (in && tt + 1 < len) ? min(len - 1 - tt, D) : 0;
			if (lane < lm) {
				const float* q = Sg + ((size_t)sh.off[u] + tt + (BWD ? 1 : 0)) * Lp + y0 + (size_t)lane * s_step;
				asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
			}
		}
	};
	// scale of step tt and the float scale terms of its phase A, for utterance u = u0 + i (whole warp; lane = d - 1 / d)
	auto next_scale = [&](uint32_t tt, int i) {
		const uint32_t u = u0 + i, len = sh.len[u];
		__syncwarp();
		if (!BWD) {
			// rho_tt = max_d (smaxd_tt[d] + Mmax + exact log-sum of alpha_{tt-d}), the segment that starts the utterance without history
			const bool a = tt < len;
			const uint32_t d = lane + 1;
			double c = -DBL_MAX;
			if (a && d <= min(tt, D)) c = (double)sm[i] + p.Mmax + sh.g_ring[u][(tt - d) & (RING - 1)];
			else if (a && d == tt + 1 && tt < D) c = (double)sm[i];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) c = fmax(c, __shfl_xor_sync(0xffffffffu, c, o));
			if (a && lane == 0) { sh.scale[u] = c; sh.s_ring[u][tt & (RING - 1)] = c; if (pt == 0) p.rho[(size_t)sh.off[u] + tt] = c; }
			float v = 0.0f;      // scf[u][lane]: rho_{tt-d} + Mmax - rho_tt (d <= tt) or -rho_tt (d == tt+1), d = lane
			if (a && lane >= 1 && lane <= D && lane <= tt + 1) v = (lane <= tt) ? (float)(sh.s_ring[u][(tt - lane) & (RING - 1)] + p.Mmax - c) : (float)(-c);
			sh.scf[u][lane] = v;
		} else {
			const bool a = tt + 1 < len && tt < len;              // (tt may have wrapped)
			const uint32_t d = lane + 1, nn = a ? min(len - 1 - tt, D) : 0;
			double c = -DBL_MAX;
			if (d <= nn) c = (double)sm[i] + sh.g_ring[u][(tt + d) & (RING - 1)];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) c = fmax(c, __shfl_xor_sync(0xffffffffu, c, o));
			if (a && lane == 0) sh.scale[u] = c;
			float v = 0.0f;      // scf[u][lane]: kappa_{tt+d} - sigma_tt, d = lane
			if (a && lane >= 1 && lane <= nn) v = (float)(sh.s_ring[u][(tt + lane) & (RING - 1)] - c);
			sh.scf[u][lane] = v;
		}
		__syncwarp();
	};

	for (uint32_t b = p.grp_off[g]; b < p.grp_off[g + 1]; b++) {
		__syncthreads();
		if (tid < UT) {
			const uint32_t u = p.batch_utt[(size_t)b * UT + tid];
			sh.utt[tid] = u;
			sh.off[tid] = u != LAB_BAD ? p.off[u] : 0;
			sh.len[tid] = u != LAB_BAD ? p.off[u + 1] - p.off[u] : 0;
			if (BWD) sh.lz[tid] = u != LAB_BAD ? p.logZ[u] : 0.0;
		}
		__syncthreads();
		uint32_t maxlen = 0;
		for (int u = 0; u < UT; u++) maxlen = max(maxlen, sh.len[u]);
		if (maxlen == 0) continue;
		{
			// first step of the batch: its score terms, and (forward) alpha_0[1,y] = S_0[1,y] scaled by the score maximum of frame 0
			const uint32_t t_first = BWD ? maxlen - 1 : 0;
			prefetch(t_first, maxlen);
			next_scale(t_first, 0); next_scale(t_first, 1);
		}

		for (uint32_t step = 0; step < maxlen; step++, gstep++) {
			const uint32_t t = BWD ? maxlen - 1 - step : step;
			const uint32_t t_next = BWD ? t - 1 : t + 1;          // (wraps behind the last backward step: prefetch / next_scale switch off)
			float* xch = xbase + (gstep & 1) * xstride;
			NTICK(7);
			// ---------------------------------------------------------------- phase A: my slice of the new vector
			//   forward  a_t[y]  = sum_d exp(S_t[d,y] + LG_{t-d}[y] + (rho_{t-d} + Mmax - rho_t))   (d == t+1: exp(S_t[d,y] - rho_t))
			//   backward bh_t[y] = sum_d exp(S_{t+d}[d,y] + LB_{t+d}[y] + (kappa_{t+d} - sigma_t))
			{
				float acc[2] = {0.0f, 0.0f};
				size_t nf[2]; bool act[2];
#pragma unroll
				for (int i = 0; i < 2; i++) {
					const uint32_t u = u0 + i, len = sh.len[u];
					nf[i] = (size_t)sh.off[u] + t;
					uint32_t limL;
					if (!BWD) { act[i] = t < len && y_ok; limL = act[i] ? min(t, D) : 0; }
					else { act[i] = t + 1 < len && y_ok; limL = act[i] ? min(len - 1 - t, D) : 0; }
					const float* hq = hist + (size_t)u * PT + lane;
#pragma unroll
					for (uint32_t d = 1; d <= (uint32_t)DR; d++) {
						const float lx = d <= limL ? hq[(size_t)((BWD ? t + d : t - d) & (RING - 1)) * (UT * PT)] : 0.0f;   // forward, d == t+1: the segment starts the utterance
						acc[i] += __expf(sn[i][d - 1] + lx + sh.scf[u][d & (RING - 1)]);
					}
				}
#pragma unroll
				for (int i = 0; i < 2; i++) {
					const uint32_t u = u0 + i, len = sh.len[u];
					if (!BWD) { if (act[i]) p.A[nf[i] * Pp + y] = acc[i]; }
					else {
						// xi_t[y',y] = a_t[y'] E[y'][y] R_{t+1}[y]: stored with the frame the new segment starts in (row shift 1 of the Xi GEMM)
						if (act[i]) p.R[(nf[i] + 1) * Pp + y] = acc[i] * expf((float)(rh[i] + p.Mmax + sh.scale[u] - sh.lz[u]));
						if (t == 0 && len > 0 && y_ok) p.R[nf[i] * Pp + y] = 0.0f;      // no transition enters the first frame
					}
					sh.tileF[frag_index(u, lane)] = acc[i];
					const float ps = warp_sum(acc[i]);
					if (lane == 0) __stcg(xch + (size_t)Pk * UT + (size_t)pt * UT + u, ps);
				}
			}
			prefetch_l2(t_next, maxlen);
			__syncthreads();
			NTICK(1);   // phase A
			// the tile in fragment order: one contiguous 2 KB block of the exchange buffer
			if (tid < PT * UT / 4) __stcg(reinterpret_cast<float4*>(xch + (size_t)y0 * UT) + tid, reinterpret_cast<const float4*>(&sh.tileF[0])[tid]);
			group_barrier(ctr, (gstep + 1) * p.npt);
			NTICK(2);   // barrier
			// ---------------------------------------------------------------- phase B: my block of the matrix product
			// the partial sums of the exchanged vector (one per CTA of the group and utterance), requested ahead of the product
			float vpart[2][2];
#pragma unroll
			for (int i = 0; i < 2; i++)
#pragma unroll
				for (int h2 = 0; h2 < 2; h2++) {
					const uint32_t q = lane + 32 * h2;
					vpart[i][h2] = q < p.npt ? __ldcg(xch + (size_t)Pk * UT + (size_t)q * UT + u0 + i) : 0.0f;
				}
			NTICK(3);
			{
				// the CTA's 16 x 32 block of the product on the tensor cores: mma.sync m16n8k16, rows = utterances, columns = the tile's
				// phones, bf16 hi/lo splits of both operands (hi*hi + hi*lo + lo*hi, fp32 accumulation: ~16 mantissa bits like every other
				// GEMM-shaped product here); warp w takes the k-steps w, w + 8, ... and the partial blocks are summed through shared memory.
				// All A fragments of a warp's k-steps (the exchanged vector, written by the other CTAs: L1 bypassed) are requested before
				// the first is used: one round trip to L2 per step.
				float acc[4][4];
#pragma unroll
				for (int nt = 0; nt < 4; nt++)
#pragma unroll
					for (int e = 0; e < 4; e++) acc[nt][e] = 0.0f;
				const uint32_t fg = lane >> 2, ft = lane & 3;
				constexpr uint32_t KB = 9;      // k-steps per warp and pass (P <= 1152 in one pass)
				for (uint32_t ks0 = warp; ks0 < n_ks; ks0 += NW * KB) {
					float2 av[KB][4];
#pragma unroll
					for (uint32_t j = 0; j < KB; j++) {
						const uint32_t ks = ks0 + j * NW;
						if (ks < n_ks) {
							const float2* r0 = reinterpret_cast<const float2*>(xch) + (size_t)ks * 128 + lane;
#pragma unroll
							for (int e = 0; e < 4; e++) av[j][e] = __ldcg(r0 + e * 32);
						}
					}
					NTICK(0);   // operand requests issued
#pragma unroll
					for (uint32_t j = 0; j < KB; j++) {
						const uint32_t ks = ks0 + j * NW;
						if (ks < n_ks) {      // (warp-uniform)
							uint32_t ah[4], al[4];
#pragma unroll
							for (int e = 0; e < 4; e++) split2(av[j][e].x, av[j][e].y, ah[e], al[e]);
							const uint4* bf = Ef + (size_t)ks * 128 + lane;
#pragma unroll
							for (int nt = 0; nt < 4; nt++) {
								const uint4 b = bf[nt * 32];
								mma_bf16(acc[nt], ah, b.x, b.y);
								mma_bf16(acc[nt], ah, b.z, b.w);
								mma_bf16(acc[nt], al, b.x, b.y);
							}
						}
					}
				}
				NTICK(6);   // operands arrived, split, multiplied
#pragma unroll
				for (int nt = 0; nt < 4; nt++) {
					*reinterpret_cast<float2*>(&sh.red[warp][fg][nt * 8 + 2 * ft]) = make_float2(acc[nt][0], acc[nt][1]);
					*reinterpret_cast<float2*>(&sh.red[warp][fg + 8][nt * 8 + 2 * ft]) = make_float2(acc[nt][2], acc[nt][3]);
				}
			}
			__syncthreads();
			NTICK(4);   // product
			prefetch(t_next, maxlen);      // in flight across the reduction and the scale bookkeeping (requested earlier, the 62 values would
			                               // sit in registers through the product and serialise its operand loads)
#pragma unroll
			for (int i = 0; i < 2; i++) {
				const uint32_t u = u0 + i, len = sh.len[u];
				const size_t n = (size_t)sh.off[u] + t;
				float s = 0.0f;
#pragma unroll
				for (int w = 0; w < NW; w++) s += sh.red[w][u][lane];
				if (t < len) {
					const bool tail = BWD && t + 1 == len;                          // setTailBeta: beta_{T-1} = 0
					const float l = tail ? 0.0f : logf(s);
					hist[(size_t)(t & (RING - 1)) * (UT * PT) + (size_t)u * PT + lane] = l;
					if (y_ok) (BWD ? p.LB : p.LG)[n * Pp + y] = l;
					// per-utterance bookkeeping by the whole warp (every lane holds the same values; lane 0 stores)
					// log of the vector's sum in float: a scale is only the reference its own frame is stored against (the entries are formed
					// from the scales exactly as stored), so this rounding does not accumulate over the frames -- the fp64 log cost more than
					// the whole reduction
					const double lvs = (double)logf(warp_sum(vpart[i][0] + vpart[i][1]));
					if (!BWD) {
						const double gh = sh.scale[u] + lvs;                          // log-sum of alpha_t
						if (lane == 0) { sh.g_ring[u][t & (RING - 1)] = gh; if (t + 1 == len && pt == 0) p.logZ[sh.utt[u]] = gh; }   // computeAlphaSum
					} else {
						const double kp = tail ? 0.0 : sh.scale[u] + p.Mmax;
						if (lane == 0) {
							sh.s_ring[u][t & (RING - 1)] = kp;
							if (pt == 0) p.kappa[n] = kp;                               // for the posterior pass
							sh.g_ring[u][t & (RING - 1)] = tail ? 0.0 : kp + lvs;
						}
					}
				}
				next_scale(t_next, i);
			}
			NTICK(5);   // reduction + scales
		}
	}
	if (timing) { for (int i = 0; i < 8; i++) p.dbg[i] = tacc[i]; p.dbg[8] = gstep; }
}

// Posteriors of the native no_dur recursion, a pass of its own behind the backward recursion (they feed nothing of the chain: inside the
// lock-step loop they were 21 k of the 56 k cycles per step at cfg5):
//   Dm[n][(d,y)] = [reference segment (d,y) ends at n] - gamma_t[d,y],
//   gamma_t[d,y] = exp(S_t[d,y] + A_{t-d}[y] + beta_t[y] - logZ) = exp(S + LB_t[y] + LG_{t-d}[y] + (rho_{t-d} + Mmax + kappa_t - logZ))
// (d <= t; d == t+1: the segment starts the utterance, exp(S + LB_t[y] + (kappa_t - logZ))); durations that do not exist get 0.
// One CTA per (frame, 256-phone tile): the scalar part per duration in fp64 by the first warp, then D coalesced rows.
__global__ void __launch_bounds__(256) nodur_post_kernel(NodurParams p, const uint32_t* __restrict__ frame_t, const uint32_t* __restrict__ frame_utt) {
	__shared__ float rcf[RING];
	const uint32_t n = blockIdx.x, y = blockIdx.y * 256 + threadIdx.x;
	const uint32_t t = frame_t[n], P = p.P, D = p.D, dmax = min(t + 1, D);
	if (threadIdx.x < RING) {
		const uint32_t d = threadIdx.x;
		float v = 0.0f;
		if (d >= 1 && d <= dmax) {
			const double kzu = p.kappa[n] - p.logZ[frame_utt[n]];
			v = (d <= t) ? (float)(p.rho[(size_t)n - d] + p.Mmax + kzu) : (float)kzu;
		}
		rcf[d] = v;
	}
	__syncthreads();
	if (y >= P) return;
	const uint32_t lab = p.node_lab[n];
	const float lcur = p.LB[(size_t)n * p.Pp + y];
	const float* Sq = p.S + (size_t)n * p.Lp + y;
	float* Dq = p.Dm + (size_t)n * p.Lp + y;
	const float* Lq = p.LG + (size_t)n * p.Pp + y;
	for (uint32_t d0 = 1; d0 <= D; d0 += DC) {
		float sv[DC], lv[DC];
#pragma unroll
		for (uint32_t j = 0; j < DC; j++) {
			const uint32_t d = d0 + j;
			const bool on = d <= dmax, onL = d <= min(t, D);
			sv[j] = on ? __ldg(Sq + (size_t)(d - 1) * P) : 0.0f;
			lv[j] = onL ? *(Lq - (size_t)d * p.Pp) : 0.0f;
		}
#pragma unroll
		for (uint32_t j = 0; j < DC; j++) {
			const uint32_t d = d0 + j, o_s = (d - 1) * P;
			if (d > D) continue;
			float dm = 0.0f;
			if (d <= dmax) dm = ((lab == o_s + y) ? 1.0f : 0.0f) - __expf(sv[j] + lcur + lv[j] + rcf[d & (RING - 1)]);
			Dq[o_s] = dm;
		}
	}
}

void launch_nodur_post(const NodurParams& p, const uint32_t* frame_t, const uint32_t* frame_utt, uint32_t N, cudaStream_t s) {
	if (!N) return;
	nodur_post_kernel<<<dim3(N, (p.P + 255) / 256), 256, 0, s>>>(p, frame_t, frame_utt);
}

namespace {
// the instantiation whose unrolled duration loops cover max_dur (the D terms of an entry wait in registers)
template <bool BWD>
const void* nodur_fn(uint32_t D) {
	if (D <= 4) return (const void*)nodur_dp_kernel<BWD, 4>;
	if (D <= 10) return (const void*)nodur_dp_kernel<BWD, 10>;
	if (D <= 16) return (const void*)nodur_dp_kernel<BWD, 16>;
	return (const void*)nodur_dp_kernel<BWD, 31>;
}
}  // namespace

int nodur_max_groups(uint32_t P) {
	int dev = 0, sms = 0, per_sm = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const size_t smem = nodur_smem_bytes(P);
	for (uint32_t D : {4u, 10u, 16u, 31u}) {
		if (cudaFuncSetAttribute(nodur_fn<false>(D), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
		if (cudaFuncSetAttribute(nodur_fn<true>(D), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
	}
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nodur_fn<true>(31), NTHR, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return 0; }
	const uint32_t npt = (P + PT - 1) / PT;
	return (int)((uint32_t)(sms * per_sm) / npt);
}

cudaError_t launch_nodur_dp(bool backward, const NodurParams& p, cudaStream_t s) {
	if (!p.n_groups) return cudaSuccess;
	if (p.D > 31) return cudaErrorInvalidValue;
	const size_t smem = nodur_smem_bytes(p.P);
	cudaError_t e = cudaMemsetAsync(p.ctr, 0, sizeof(uint32_t) * p.n_groups, s);
	if (e != cudaSuccess) return e;
	NodurParams q = p;
	void* args[] = {&q};
	const void* fn = backward ? nodur_fn<true>(p.D) : nodur_fn<false>(p.D);
	return cudaLaunchCooperativeKernel(fn, dim3(p.n_groups * p.npt), dim3(NTHR), args, smem, s);
}

}  // namespace crfgpu
