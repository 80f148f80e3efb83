// Tensor-core, cluster-resident lattice recursions (K2/K3/K4) for sm_100a.
//
// Reference semantics: CRF_StdSegStateNode::computeAlpha / computeBeta / computeExpF / computeAlphaSum
// (CRF/src/nodes/CRF_StdSegStateNode.cpp:135-186, 219-308, 343-438, 447-462); with max_dur == 1 these are
// CRF_StdStateNode's (CRF/src/nodes/CRF_StdStateNode.cpp:81-299) and the N-state topology
// (CRF_StdNStateNode.cpp:110-365) is the same recursion with E == 0 on illegal pairs.
//
// Per frame the recursions need one product with the L x L matrix E = exp(M - Mmax).  A thread-block CLUSTER of
// CS CTAs keeps E on chip for the whole launch and advances UB = 16 utterances ("slots") in lock-step:
//   * CTA r owns the label slice [r*CW, (r+1)*CW): its rows of E (backward) / E^T (forward), split into bf16
//     hi + lo.  The hi half lives in TENSOR MEMORY (tcgen05.mma A operand from TMEM, K/2 columns), the lo half in
//     shared memory, so that the 16-slot frame vectors fit beside it.
//   * per step:  D[128 x 16] = E_slice[128 x K] * V[K x 16]  as K/16 * 3 tcgen05.mma (hi*hi + hi*lo + lo*hi),
//     fp32 accumulators in TMEM, ~16 mantissa bits (tools/tc_probe.cu: 3e-6 worst relative error).
//   * the only exchange is the all-gather of the new frame vector: every CTA writes its slice (already in the
//     MMA's shared-memory layout, bf16 hi/lo) and pushes it to its CS-1 peers with ONE bulk DSMEM copy each
//     (cp.async.bulk.shared::cluster, completion counted on the receiver's mbarrier); no cluster barrier and
//     no max-reduction per frame (tools/dsmem_probe.cu measured 60 cycles/KB for this exchange).
//   * scales: every frame vector is stored relative to a log scale that is an UPPER BOUND computable one step
//     ahead from per-duration score maxima and the previous frames' sums, so that entries are <= 1 and the
//     largest is >= exp(Mmin - Mmax): alpha_t[c] = rho_t + log A_t[c];  S_t[c] + beta_t[c] = base_t + lw_t[c].
//   * slots are refilled from the cluster's utterance list as soon as an utterance ends (continuous batching).
//
// Forward:   G_t[c] = sum_q A_t[q] E[q][c]   (pushed once per frame, block d of G_t is consumed at t+d)
//            alpha_{t+1}[(d,y)] = S_{t+1}[(d,y)] + Mmax + rho_{t+1-d} + log G_{t+1-d}[(d,y)]     (d <= t+1)
//                               = S_{t+1}[(d,y)]                                                 (d == t+2, segment starts the utterance)
// Backward:  u_t[q] = sum_c E[q][c] v_t[c],  v_t[(d,y)] = exp(S+beta of frame t+d, label (d,y), - sigma_t)
//            beta_t[q] = Mmax + sigma_t + log u_t[q] =: base_t + log u_t[q]
//            gamma_t[q] = A_t[q] u_t[q] exp(rho_t + base_t - logZ)                                (:369-372)
//            xi_t(q,c)  = A_{t-d}[q] E[q][c] R_t[c],  R_t[c] = exp(S_t[c]+beta_t[c]+rho_{t-d}+Mmax-logZ)   (:389-397)
#include "crf_kernels.cuh"
#include "tc05.cuh"

#include <cfloat>
#include <cstdio>

namespace crfgpu {

using namespace tc05;

namespace {

constexpr int UB = 16;             // slots per cluster == MMA N
constexpr int DMAX = 32;
constexpr int N_LANE_THREADS = 256;
constexpr int N_THREADS = 288;     // 8 lane warps + 1 control warp
constexpr uint32_t D_COL = 0, EHI_COL = 32;

struct Slot {
	uint32_t utt, off, len, t;     // utt == LAB_BAD: idle
	double lz;                     // logZ of the utterance (backward)
};

struct Ctl {
	uint64_t gather[2], mma_bar;
	uint32_t tmem, any[2], pad;
	Slot cur[2][UB], nxt[2][UB];   // [step parity][slot]: frame held by the gathered vector / frame being produced
	double ring_a[UB][DMAX];       // forward: ghat_t = log sum_c alpha_t   | backward: bl_t = base_t + log sum_c v_t
	double ring_b[UB][DMAX];       // forward: rho_t                        | backward: base_t
	float delta[UB][DMAX + 1];     // log-scale correction of duration block d for the vector being produced
	float rsc[UB][DMAX + 1];       // backward: scale of R for duration d
	float sg[UB];                  // backward: rho_t + base_t - logZ
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
	uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(N_THREADS) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(N_THREADS) : "memory"); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ uint16_t bf16_bits(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }

}  // namespace

size_t tc_dp_ctl_bytes() { return sizeof(Ctl); }

template <bool BWD>
__global__ void __launch_bounds__(N_THREADS, 1) dp_tc_kernel(TcDpParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	const uint32_t L = p.L, Lp = p.Lp, P = p.P, D = p.D, CS = p.CS, CW = p.CW, K = p.K;
	const uint32_t CHUNK = CW * 64 + 256;                  // hi/lo tile of one slice for 16 slots + [4 quadrants][16] partial sums
	const uint32_t ELO_BYTES = CW * K * 2, SBO_E = K * 16;
	unsigned char* elo = smem;
	unsigned char* bbuf = smem + ELO_BYTES;                // [2][CS][CHUNK]
	Ctl* ctl = reinterpret_cast<Ctl*>(smem + p.ctl_off);

	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t rank = CS > 1 ? cluster_ctarank() : 0, cl = blockIdx.x / CS;
	const uint32_t c0 = rank * CW;
	const uint32_t q4 = warp & 3, half = (warp >> 2) & 1;
	const uint32_t row = q4 * 32 + lane, c = c0 + row;
	const bool lane_thread = warp < 8;
	const bool in_tile = lane_thread && row < CW;          // this thread owns a row of the slice tile
	const bool row_valid = in_tile && c < L;               // ... that is a real label
	const uint32_t my_d = row_valid ? c / P + 1 : 0xffffu;
	const uint32_t list_begin = p.cl_off[cl], list_end = p.cl_off[cl + 1];
	const float* Msrc = BWD ? p.E : p.ET;                  // rows = my labels, columns = the contracted label

	// ---------------------------------------------------------------- setup
	if (tid == 0) {
		mbar_init(&ctl->gather[0], 1); mbar_init(&ctl->gather[1], 1); mbar_init(&ctl->mma_bar, 1);
		fence_mbar_init();
	}
	if (warp == 8) tmem_alloc(&ctl->tmem, p.tmem_cols);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = ctl->tmem;
	if (lane_thread && q4 * 32 < CW) {
		// my row of the E slice: hi -> TMEM (8 packed columns per 16-wide k-step), lo -> shared memory (K-major, LBO 128, SBO K*16)
		const uint32_t KS = K / 16, ks_lo = half ? KS / 2 : 0, ks_hi = half ? KS : KS / 2;
		for (uint32_t ks = ks_lo; ks < ks_hi; ks++) {
			float x[16];
			const float* src = Msrc + (size_t)c * Lp + ks * 16;
#pragma unroll
			for (int j = 0; j < 16; j++) x[j] = (row_valid && ks * 16 + j < L) ? __ldg(src + j) : 0.0f;
			const float (&x0)[8] = *reinterpret_cast<const float (*)[8]>(&x[0]);
			const float (&x1)[8] = *reinterpret_cast<const float (*)[8]>(&x[8]);
			uint4 h0, l0, h1, l1;
			split8(x0, h0, l0); split8(x1, h1, l1);
			if (in_tile) {
				unsigned char* dst = elo + (row / 8) * SBO_E + (2 * ks) * 128 + (row % 8) * 16;
				*reinterpret_cast<uint4*>(dst) = l0; *reinterpret_cast<uint4*>(dst + 128) = l1;
			}
			const uint32_t r[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
			tmem_st8(tmem + ((q4 * 32u) << 16) + EHI_COL + ks * 8, r);
		}
		tmem_st_wait();
	}
	if (warp == 8) {
		// schedule of step 0: nothing gathered yet, the first UB utterances of the list are produced at their first frame
		if (lane < UB) {
			Slot idle{LAB_BAD, 0, 0, 0, 0.0}, n = idle;
			const uint32_t idx = list_begin + lane;
			if (idx < list_end) {
				const uint32_t utt = p.cl_list[idx];
				n.utt = utt; n.off = p.off[utt]; n.len = p.off[utt + 1] - p.off[utt]; n.t = BWD ? n.len - 1 : 0;
				if (BWD) n.lz = p.logZ[utt];
			}
			ctl->cur[0][lane] = idle; ctl->nxt[0][lane] = n;
		}
		if (lane == 0) ctl->any[0] = list_begin < list_end ? 1u : 0u;
	}
	fence_proxy_async_smem();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	if (CS > 1) cluster_sync_all();      // every CTA's mbarriers exist before the first remote completion can arrive
	uint32_t list_next = min(list_begin + (uint32_t)UB, list_end);

	// ---------------------------------------------------------------- steps
	uint32_t it = 0;
	for (;; it++) {
		const uint32_t pb = it & 1, nb = pb ^ 1;
		if (!ctl->any[pb]) break;
		unsigned char* mychunk = bbuf + ((size_t)nb * CS + rank) * CHUNK;
		if (warp == 8) {
			// ===================== control warp =====================
			if (it > 0) {
				mbar_wait_cluster(&ctl->gather[pb], ((it - 1) >> 1) & 1);
				if (lane == 0) {
					tc_fence_after();
					const uint32_t idesc = idesc_bf16_f32(128, UB, false, false);
					const uint32_t bb = smem_u32(bbuf + (size_t)pb * CS * CHUNK), eb = smem_u32(elo);
					const uint32_t ks_per_chunk = CW / 16;
					uint32_t ks = 0;
					for (uint32_t j = 0; j < CS; j++) {
						for (uint32_t k2 = 0; k2 < ks_per_chunk; k2++, ks++) {
							const uint32_t ba = bb + j * CHUNK + k2 * 1024;
							const uint64_t vhi = smem_desc(ba, 512, 128), vlo = smem_desc(ba + 256, 512, 128);
							const uint64_t el = smem_desc(eb + ks * 256, 128, SBO_E);
							mma_ts(tmem + D_COL, tmem + EHI_COL + ks * 8, vhi, idesc, ks > 0);
							mma_ts(tmem + D_COL, tmem + EHI_COL + ks * 8, vlo, idesc, true);
							mma_ss(tmem + D_COL, el, vhi, idesc, true);
						}
					}
					mma_commit(&ctl->mma_bar);
				}
				__syncwarp();
			}
			Slot c2{LAB_BAD, 0, 0, 0, 0.0}, n2 = c2;
			bool want_refill = false;
			if (lane < UB) {
				const Slot cur = ctl->cur[pb][lane], nxt = ctl->nxt[pb][lane];
				// ---- scale bookkeeping ----
				float vsum = 0.0f;
				if (cur.utt != LAB_BAD) {
					const unsigned char* base = bbuf + (size_t)pb * CS * CHUNK + CW * 64;
					for (uint32_t j = 0; j < CS; j++)
#pragma unroll
						for (int q = 0; q < 4; q++) vsum += *reinterpret_cast<const float*>(base + j * CHUNK + (q * 16 + lane) * 4);
				}
				if (!BWD) {
					if (cur.utt != LAB_BAD) {
						const double ghat = ctl->ring_b[lane][cur.t % D] + log((double)vsum);
						ctl->ring_a[lane][cur.t % D] = ghat;
						if (cur.t + 1 == cur.len && rank == 0) p.logZ[cur.utt] = ghat;   // computeAlphaSum (:447-462)
					}
					if (nxt.utt != LAB_BAD) {
						const uint32_t t1 = nxt.t;
						const float* sm = p.smaxd + ((size_t)nxt.off + t1) * D;
						double rho = -DBL_MAX;
						for (uint32_t d = 1; d <= min(t1, D); d++) rho = fmax(rho, (double)sm[d - 1] + p.Mmax + ctl->ring_a[lane][(t1 - d) % D]);
						if (t1 < D) rho = fmax(rho, (double)sm[t1]);            // d == t1+1: the segment starts the utterance, alpha = S
						for (uint32_t d = 1; d <= D; d++) {
							float dl = -INFINITY;
							if (d <= t1) dl = (float)(p.Mmax + ctl->ring_b[lane][(t1 - d) % D] - rho);
							else if (d == t1 + 1) dl = (float)(-rho);
							ctl->delta[lane][d] = dl;
						}
						ctl->ring_b[lane][t1 % D] = rho;
						if (rank == 0) p.m[(size_t)nxt.off + t1] = rho;
					}
				} else {
					if (cur.utt != LAB_BAD) {
						const uint32_t t = cur.t; const size_t n = (size_t)cur.off + t;
						const bool tail = t + 1 == cur.len;
						const double base = ctl->ring_b[lane][t % D];
						ctl->ring_a[lane][t % D] = tail ? 0.0 : base + log((double)vsum);
						const double rho = p.m[n];
						ctl->sg[lane] = (float)(rho + base - cur.lz);
						for (uint32_t d = 1; d <= D; d++)
							ctl->rsc[lane][d] = (d <= t) ? (float)(base + p.m[n - d] + p.Mmax - cur.lz) : -INFINITY;
						if (rank == 0) p.bbase[n] = base;
					}
					if (nxt.utt != LAB_BAD) {
						const uint32_t t1 = nxt.t; const size_t n1 = (size_t)nxt.off + t1;
						const uint32_t numNext = min(nxt.len - 1 - t1, D);
						double sigma = -DBL_MAX;
						for (uint32_t d = 1; d <= numNext; d++)
							sigma = fmax(sigma, (double)p.smaxd[(n1 + d) * D + d - 1] + ctl->ring_a[lane][(t1 + d) % D]);
						for (uint32_t d = 1; d <= D; d++)
							ctl->delta[lane][d] = (d <= numNext) ? (float)(ctl->ring_b[lane][(t1 + d) % D] - sigma) : -INFINITY;
						ctl->ring_b[lane][t1 % D] = numNext ? p.Mmax + sigma : 0.0;   // tail: beta = 0 (setTailBeta)
					}
				}
				// ---- schedule of the next step ----
				c2 = nxt;
				if (c2.utt != LAB_BAD && (BWD ? c2.t > 0 : c2.t + 1 < c2.len)) { n2 = c2; n2.t = BWD ? c2.t - 1 : c2.t + 1; }
				else want_refill = true;
			}
			const uint32_t mask = __ballot_sync(0xffffffffu, want_refill);
			if (lane < UB) {
				if (want_refill) {
					const uint32_t idx = list_next + __popc(mask & ((1u << lane) - 1u));
					if (idx < list_end) {
						const uint32_t utt = p.cl_list[idx];
						n2.utt = utt; n2.off = p.off[utt]; n2.len = p.off[utt + 1] - p.off[utt]; n2.t = BWD ? n2.len - 1 : 0;
						if (BWD) n2.lz = p.logZ[utt];
					}
				}
				ctl->cur[nb][lane] = c2; ctl->nxt[nb][lane] = n2;
			}
			list_next = min(list_next + (uint32_t)__popc(mask), list_end);
			const uint32_t any = __ballot_sync(0xffffffffu, lane < UB && (c2.utt != LAB_BAD || n2.utt != LAB_BAD));
			if (lane == 0) ctl->any[nb] = any ? 1u : 0u;
			__syncwarp();
			bar_arrive(1);          // scales + next schedule are published
			bar_sync(2);            // my slice of the new vector is complete in shared memory
			if (lane == 0) mbar_arrive_expect_tx(&ctl->gather[nb], (CS - 1) * CHUNK);
			if (lane < CS && lane != rank) {
				const uint32_t dst = mapa(smem_u32(mychunk), lane), rbar = mapa(smem_u32(&ctl->gather[nb]), lane);
				asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				             ::"r"(dst), "r"(smem_u32(mychunk)), "r"(CHUNK), "r"(rbar) : "memory");
			}
		} else {
			// ===================== lane threads: row `row` of the slice, slots half*8 .. half*8+7 =====================
			float sv[8], aux[8], old[8];
			uint32_t lab[8];
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const uint32_t s = half * 8 + j;
				sv[j] = 0.0f; aux[j] = 0.0f; old[j] = 0.0f; lab[j] = LAB_BAD;
				if (!row_valid) continue;
				const Slot nxt = ctl->nxt[pb][s];
				if (!BWD) {
					if (nxt.utt != LAB_BAD) {
						const size_t n1 = (size_t)nxt.off + nxt.t;
						sv[j] = __ldg(p.S + n1 * Lp + c);
						if (my_d >= 2 && my_d <= nxt.t) old[j] = p.G[(n1 - my_d) * Lp + c];
					}
				} else {
					const Slot cur = ctl->cur[pb][s];
					if (cur.utt != LAB_BAD) {
						const size_t n = (size_t)cur.off + cur.t;
						sv[j] = __ldg(p.S + n * Lp + c); aux[j] = __ldg(p.A + n * Lp + c); lab[j] = __ldg(p.node_lab + n);
					}
					if (nxt.utt != LAB_BAD && my_d >= 2 && my_d <= min(nxt.len - 1 - nxt.t, D))
						old[j] = p.G[((size_t)nxt.off + nxt.t + my_d) * Lp + c];
				}
			}
			bar_sync(1);
			float g[8];
#pragma unroll
			for (int j = 0; j < 8; j++) g[j] = 0.0f;
			if (it > 0) {
				mbar_wait(&ctl->mma_bar, (it - 1) & 1);
				tc_fence_after();
				tmem_ld8(tmem + ((q4 * 32u) << 16) + D_COL + half * 8, g);
				tmem_ld_wait();
				tc_fence_before();
			}
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const uint32_t s = half * 8 + j;
				const Slot cur = ctl->cur[pb][s], nxt = ctl->nxt[pb][s];
				float val = 0.0f;
				if (!BWD) {
					if (row_valid) {
						if (cur.utt != LAB_BAD && cur.t + 1 < cur.len) p.G[((size_t)cur.off + cur.t) * Lp + c] = g[j];
						if (nxt.utt != LAB_BAD) {
							const uint32_t t1 = nxt.t;
							if (c < P * min(t1 + 1, D)) {
								float lr = sv[j] + ctl->delta[s][my_d];
								if (my_d <= t1) lr += __logf(my_d == 1 ? g[j] : old[j]);
								val = __expf(lr);
							}
							p.A[((size_t)nxt.off + t1) * Lp + c] = val;
						}
					}
				} else {
					float lw = -INFINITY;
					if (row_valid && cur.utt != LAB_BAD) {
						const uint32_t t = cur.t; const size_t n = (size_t)cur.off + t;
						float dm = 0.0f, r = 0.0f, uu = 0.0f;
						if (c < P * min(t + 1, D)) {
							uu = (t + 1 == cur.len) ? 1.0f : g[j];
							const float lu = __logf(uu);
							lw = sv[j] + lu;
							const float gamma = aux[j] * __expf(lu + ctl->sg[s]);
							dm = ((lab[j] == c) ? 1.0f : 0.0f) - gamma;
							if (my_d <= t) r = __expf(lw + ctl->rsc[s][my_d]);
						}
						p.Dm[n * Lp + c] = dm; p.R[n * Lp + c] = r;
						if (p.Uvec) p.Uvec[n * Lp + c] = uu;
						p.G[n * Lp + c] = lw;          // log-domain S+beta relative to base_t, read back by this thread d frames earlier
					}
					if (row_valid && nxt.utt != LAB_BAD && my_d <= min(nxt.len - 1 - nxt.t, D))
						val = __expf((my_d == 1 ? lw : old[j]) + ctl->delta[s][my_d]);
				}
				if (in_tile) {
					const __nv_bfloat16 hi = __float2bfloat16_rn(val);
					const __nv_bfloat16 lo = __float2bfloat16_rn(val - __bfloat162float(hi));
					unsigned char* dst = mychunk + (row / 8) * 512 + half * 128 + j * 16 + (row % 8) * 2;
					*reinterpret_cast<__nv_bfloat16*>(dst) = hi; *reinterpret_cast<__nv_bfloat16*>(dst + 256) = lo;
				}
				const float ws = warp_sum(val);
				if (lane == 0) *reinterpret_cast<float*>(mychunk + CW * 64 + (q4 * 16 + s) * 4) = ws;
			}
			fence_proxy_async_smem();
			bar_arrive(2);
		}
	}
	// the vectors pushed during the last executed step are never consumed: wait for them so that no bulk copy is in
	// flight (into this CTA or out of it) when the cluster retires
	if (warp == 8 && it > 0) mbar_wait_cluster(&ctl->gather[it & 1], ((it - 1) >> 1) & 1);
	tc_fence_before();
	__syncthreads();
	if (CS > 1) cluster_sync_all();
	if (warp == 8) tmem_dealloc(tmem, p.tmem_cols);
}

// per-duration maxima of the state scores: smaxd[n][d-1] = max_y S_n[(d,y)]  (-inf when the block does not exist, d > t+1)
__global__ void __launch_bounds__(256) block_max_kernel(const float* S, const uint32_t* frame_t, float* smaxd, uint32_t N, uint32_t Lp, uint32_t P, uint32_t D) {
	const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t n = w; n < N; n += nw) {
		const uint32_t t = frame_t[n];
		for (uint32_t d = 0; d < D; d++) {
			float m = -INFINITY;
			if (d <= t) for (uint32_t y = lane; y < P; y += 32) m = fmaxf(m, S[(size_t)n * Lp + d * P + y]);
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
			if (lane == 0) smaxd[(size_t)n * D + d] = m;
		}
	}
}
void launch_block_max(const float* S, const uint32_t* frame_t, float* smaxd, uint32_t N, uint32_t Lp, uint32_t P, uint32_t D, cudaStream_t s) {
	if (!N) return;
	unsigned blocks = (N + 7) / 8; if (blocks > 148 * 8) blocks = 148 * 8;
	block_max_kernel<<<blocks, 256, 0, s>>>(S, frame_t, smaxd, N, Lp, P, D);
}

// ------------------------------------------------------------------------------------------------
bool plan_tc_dp(uint32_t L, uint32_t D, int max_smem_optin, TcDpPlan* plan) {
	if (D > DMAX) return false;
	for (uint32_t CS = 1; CS <= 8; CS *= 2) {
		const uint32_t CW = ((L + CS - 1) / CS + 15) / 16 * 16;
		if (CW > 128) continue;
		if (CS > 1 && (CS - 1) * CW >= L) continue;              // an empty slice
		const uint32_t K = CS * CW;
		if (EHI_COL + K / 2 > 512) continue;
		uint32_t cols = 32; while (cols < EHI_COL + K / 2) cols *= 2;
		const size_t chunk = (size_t)CW * 64 + 256;
		size_t body = (size_t)CW * K * 2 + 2 * CS * chunk;
		const size_t reach = (size_t)16 * K * 16 + 256;           // the MMA reads 128 rows of the lo tile: rows >= CW alias what follows it
		if (body < reach) body = reach;
		const size_t ctl_off = (body + 127) / 128 * 128;
		size_t total = ctl_off + tc_dp_ctl_bytes();
		if (total + 1024 > (size_t)max_smem_optin) continue;
		// resident CTAs own their TMEM columns until they exit: size the shared-memory request so that no more CTAs
		// fit on an SM than its 512 TMEM columns can serve (an over-subscribed tcgen05.alloc would stall a whole cluster)
		const size_t min_smem = (size_t)233472 / (512 / cols + 1) + 1;
		if (total < min_smem) total = min_smem;
		plan->CS = CS; plan->CW = CW; plan->K = K; plan->tmem_cols = cols; plan->ctl_off = (uint32_t)ctl_off; plan->smem = total;
		return true;
	}
	return false;
}

static cudaError_t configure(void* kern, const TcDpPlan& plan, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, uint32_t n_clusters, cudaStream_t s) {
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
	if (e != cudaSuccess) return e;
	cfg->gridDim = dim3(n_clusters * plan.CS); cfg->blockDim = dim3(N_THREADS);
	cfg->dynamicSmemBytes = plan.smem; cfg->stream = s;
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = plan.CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg->attrs = attr; cfg->numAttrs = 1;
	return cudaSuccess;
}

int max_active_tc_clusters(const TcDpPlan& plan) {
	cudaLaunchConfig_t cfg{}; cudaLaunchAttribute attr[1];
	if (configure((void*)dp_tc_kernel<false>, plan, &cfg, attr, 1, nullptr) != cudaSuccess) { cudaGetLastError(); return 0; }
	int n = 0;
	if (cudaOccupancyMaxActiveClusters(&n, dp_tc_kernel<false>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
	// the kernels own the SM's tensor memory while resident: at most 512 / tmem_cols CTAs per SM
	return n;
}

cudaError_t launch_tc_dp(bool backward, const TcDpParams& p, const TcDpPlan& plan, cudaStream_t s) {
	if (!p.n_clusters) return cudaSuccess;
	cudaLaunchConfig_t cfg{}; cudaLaunchAttribute attr[1];
	cudaError_t e = configure(backward ? (void*)dp_tc_kernel<true> : (void*)dp_tc_kernel<false>, plan, &cfg, attr, p.n_clusters, s);
	if (e != cudaSuccess) return e;
	return backward ? cudaLaunchKernelEx(&cfg, dp_tc_kernel<true>, p) : cudaLaunchKernelEx(&cfg, dp_tc_kernel<false>, p);
}

}  // namespace crfgpu
