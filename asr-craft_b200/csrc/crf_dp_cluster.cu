// Cluster-resident lattice recursions (K2/K3/K4) for sm_100a.
//
// The dense recursions need, per frame, one mat-vec with the L x L transition matrix E = exp(M - Mmax).
// For cfg4 (L = 610) E is 1.49 MB: streaming it from L2 for every frame made the first version of these
// kernels load-latency bound (profiles/r1a_backward_kernel_full.md).  Here a thread-block CLUSTER of CS CTAs
// keeps E on chip for the whole launch: CTA r holds the column slice [r*CW, (r+1)*CW) in shared memory
// (186 kB for L=610, CS=8), the cluster advances UB utterances ("slots") in lock-step and the only per-frame
// exchange is the UB x L log-domain frame vector, published through L2 with ONE cluster barrier per frame.
// Slots are refilled from the cluster's utterance list as soon as an utterance ends (continuous batching),
// so ragged lengths do not idle the FMA pipes.
//
// Math and reference citations are those of crf_kernels.cu (forward_kernel / backward_kernel):
// CRF_StdSegStateNode::computeAlpha / computeBeta / computeExpF / computeAlphaSum
// (CRF/src/nodes/CRF_StdSegStateNode.cpp:135-186, 219-308, 343-438, 447-462).
#include "crf_kernels.cuh"

#include <cfloat>
#include <cstdio>

namespace crfgpu {

namespace {

__device__ __forceinline__ int float_key(float f) {
	int i = __float_as_int(f);
	return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
constexpr int KEY_NEG_INF = (int)0x807fffff;

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
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
// all threads of all CTAs of the cluster; release/acquire orders the global-memory exchange
__device__ __forceinline__ void cluster_barrier(bool clustered) {
	if (clustered) {
		asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
		asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	} else {
		__syncthreads();
	}
}

struct SmemLayout {
	size_t e_bytes, a_bytes, total;
};
__host__ __device__ inline SmemLayout smem_layout(uint32_t L, uint32_t CWp, int UB, uint32_t D) {
	SmemLayout s;
	s.e_bytes = ((size_t)L * CWp * sizeof(float) + 15) / 16 * 16;
	s.a_bytes = ((size_t)L * UB * sizeof(float) + 15) / 16 * 16;
	size_t misc = sizeof(double) * ((size_t)UB * D + 4 * UB)   // ring + mu/base, zsum/sg, gmax-as-double spare
	            + sizeof(float) * (2 * (size_t)UB * (D + 1) + UB) // delta, rsc, gmax
	            + sizeof(int) * (2 * UB)                          // keys
	            + sizeof(uint32_t) * (4 * UB + 8);                // slot state
	s.total = s.e_bytes + s.a_bytes + (misc + 15) / 16 * 16;
	return s;
}

}  // namespace

// =================================================================================================
// forward
// =================================================================================================
template <int UB>
__global__ void __launch_bounds__(512, 1) forward_cluster_kernel(ClusterDpParams p) {
	constexpr int UQ = UB / 4;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const SmemLayout lay = smem_layout(p.L, p.CWp, UB, p.D);
	float* E_s = reinterpret_cast<float*>(smem_raw);                         // [L][CWp]  E[q][c0+j]
	float4* A_s = reinterpret_cast<float4*>(smem_raw + lay.e_bytes);          // [UQ][L] of float4 (4 slots)
	double* m_ring = reinterpret_cast<double*>(smem_raw + lay.e_bytes + lay.a_bytes);   // [UB][D]
	double* mu_s = m_ring + (size_t)UB * p.D;                                 // [UB]
	double* zsum_s = mu_s + UB;                                               // [UB]
	float* delta_s = reinterpret_cast<float*>(zsum_s + 3 * UB);               // [UB][D+1]
	float* gmax_s = delta_s + 2 * (size_t)UB * (p.D + 1);                     // [UB]
	int* key_s = reinterpret_cast<int*>(gmax_s + UB);                         // [UB]
	uint32_t* s_utt = reinterpret_cast<uint32_t*>(key_s + 2 * UB);            // [UB]
	uint32_t* s_off = s_utt + UB;                                             // [UB]
	uint32_t* s_len = s_off + UB;                                             // [UB]  0 = idle
	uint32_t* s_t = s_len + UB;                                               // [UB]
	uint32_t* s_ctl = s_t + UB;                                               // [0] next list index, [1] any active

	const uint32_t L = p.L, Lp = p.Lp, P = p.P, D = p.D, CWp = p.CWp;
	const bool clustered = p.CS > 1;
	const uint32_t rank = clustered ? cluster_ctarank() : 0;
	const uint32_t cl = blockIdx.x / p.CS;
	const uint32_t c0 = rank * p.CW;
	const uint32_t ncol = c0 < L ? min(p.CW, L - c0) : 0;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
	const uint32_t cj = tid % p.CWt, uq = tid / p.CWt;       // mat-vec mapping: one column x 4 slots
	const bool mv_thread = (uq < UQ) && (cj < ncol);
	const uint32_t my_c = c0 + cj;
	const uint32_t my_d = mv_thread ? my_c / P + 1 : 0xffffu;
	float* xch = p.xch + (size_t)cl * 2 * UB * Lp;           // [2][UB][Lp] log-domain frame vectors
	float* xmax = p.xmax + (size_t)cl * 2 * p.CS * UB;       // [2][CS][UB]
	const uint32_t list_begin = p.cl_off[cl], list_end = p.cl_off[cl + 1];

	for (uint32_t i = tid; i < L * CWp; i += blockDim.x) {
		const uint32_t q = i / CWp, j = i % CWp;
		E_s[i] = (j < ncol) ? p.E[(size_t)q * Lp + c0 + j] : 0.0f;
	}
	if (tid < UB) { s_len[tid] = 0; s_t[tid] = 0; s_utt[tid] = LAB_BAD; key_s[tid] = KEY_NEG_INF; zsum_s[tid] = 0.0; }
	if (tid == 0) s_ctl[0] = list_begin;
	__syncthreads();

	for (uint32_t step = 0;; step++) {
		const uint32_t buf = step & 1;
		// ---- slot management: every CTA of the cluster replays the same deterministic schedule ----
		if (tid == 0) {
			uint32_t next = s_ctl[0], any = 0;
			for (int s = 0; s < UB; s++) {
				if (s_len[s] && s_t[s] + 1 < s_len[s]) { s_t[s]++; any = 1; continue; }   // advance to the next frame
				if (next < list_end) {
					const uint32_t utt = p.cl_list[next++];
					s_utt[s] = utt; s_off[s] = p.off[utt]; s_len[s] = p.off[utt + 1] - p.off[utt]; s_t[s] = 0; any = 1;
				} else s_len[s] = 0;
			}
			s_ctl[0] = next; s_ctl[1] = any;
		}
		__syncthreads();
		if (!s_ctl[1]) break;
		// ---- [A] scales: warp serves slot, lane d-1 serves duration d ----
		for (uint32_t s = warp; s < UB; s += n_warps) {
			if (s_len[s]) {
				const uint32_t t = s_t[s], d = lane + 1;
				double val = -DBL_MAX;
				if (d <= D) {
					if (d <= t) val = m_ring[s * D + (t - d) % D] + p.Mmax;
					else if (d == t + 1) val = 0.0;
				}
				const double mu = warp_max_d(val);
				if (d <= D) delta_s[s * (D + 1) + d] = (val == -DBL_MAX) ? -INFINITY : (float)(val - mu);
				if (lane == 0) { mu_s[s] = mu; }
			}
		}
		if (tid < UB) key_s[tid] = KEY_NEG_INF;
		__syncthreads();
		// ---- [B] log-domain candidates of my column for my 4 slots; publish ----
		if (uq < UQ) {
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t s = uq * 4 + k;
				float lr = -INFINITY;
				if (mv_thread && s_len[s]) {
					const uint32_t t = s_t[s];
					const uint32_t avail = P * min(t + 1, D);
					const uint64_t n = (uint64_t)s_off[s] + t;
					if (my_c < avail) {
						float lg = 0.0f;
						if (my_d <= t) lg = logf(p.G[(n - my_d) * Lp + my_c]);
						lr = p.S[n * Lp + my_c] + lg + delta_s[s * (D + 1) + my_d];
					}
					xch[((size_t)buf * UB + s) * Lp + my_c] = lr;
				}
				const float wm = warp_max(lr);
				if (lane == 0 && wm > -INFINITY) atomicMax(&key_s[s], float_key(wm));
			}
		}
		__syncthreads();
		if (tid < UB) xmax[((size_t)buf * p.CS + rank) * UB + tid] = key_float(key_s[tid]);
		cluster_barrier(clustered);
		// ---- [C] global max per slot, scale bookkeeping ----
		if (tid < UB && s_len[tid]) {
			float g = -INFINITY;
			for (uint32_t r = 0; r < p.CS; r++) g = fmaxf(g, xmax[((size_t)buf * p.CS + r) * UB + tid]);
			gmax_s[tid] = g;
			const uint32_t t = s_t[tid];
			const double mt = mu_s[tid] + (double)g;
			m_ring[tid * D + t % D] = mt;
			if (rank == 0) p.m[(uint64_t)s_off[tid] + t] = mt;
			zsum_s[tid] = 0.0;
		}
		__syncthreads();
		// ---- [D] every CTA rebuilds the full probability-domain vector A_t of each slot ----
		for (uint32_t i = tid; i < (uint32_t)UQ * L; i += blockDim.x) {
			const uint32_t g4 = i / L, q = i % L;
			float v[4];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t s = g4 * 4 + k;
				v[k] = 0.0f;
				if (s_len[s]) {
					const float lr = xch[((size_t)buf * UB + s) * Lp + q];
					v[k] = expf(lr - gmax_s[s]);        // -inf (label not available yet) -> 0
					if (q >= c0 && q < c0 + ncol) p.A[((uint64_t)s_off[s] + s_t[s]) * Lp + q] = v[k];
				}
			}
			A_s[(size_t)g4 * L + q] = make_float4(v[0], v[1], v[2], v[3]);
			// logZ = m_{T-1} + log sum_q A_{T-1}[q]  (computeAlphaSum :447-462), once per utterance, rank 0 only
			if (rank == 0) {
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const uint32_t s = g4 * 4 + k;
					if (s_len[s] && s_t[s] + 1 == s_len[s] && v[k] != 0.0f) atomicAdd(&zsum_s[s], (double)v[k]);
				}
			}
		}
		__syncthreads();
		if (rank == 0 && tid < UB && s_len[tid] && s_t[tid] + 1 == s_len[tid])
			p.logZ[s_utt[tid]] = m_ring[tid * D + s_t[tid] % D] + log(zsum_s[tid]);
		// ---- [E] push: G_t[c] = sum_q A_t[q] * E[q][c] for my column, my 4 slots ----
		if (mv_thread) {
			uint32_t qmax = 0; bool need = false;
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t s = uq * 4 + k;
				if (s_len[s] && s_t[s] + 1 < s_len[s]) { need = true; }
			}
			for (int s = 0; s < UB; s++) if (s_len[s]) qmax = max(qmax, P * min(s_t[s] + 1, D));
			if (need) {
				float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
				const float* Ec = E_s + cj;
				const float4* Au = A_s + (size_t)uq * L;
				uint32_t q = 0;
				for (; q + 4 <= qmax; q += 4) {
					const float e0 = Ec[(q + 0) * CWp], e1 = Ec[(q + 1) * CWp], e2 = Ec[(q + 2) * CWp], e3 = Ec[(q + 3) * CWp];
					const float4 a0 = Au[q + 0], a1 = Au[q + 1], a2 = Au[q + 2], a3 = Au[q + 3];
					acc0 = fmaf(a0.x, e0, acc0); acc1 = fmaf(a0.y, e0, acc1); acc2 = fmaf(a0.z, e0, acc2); acc3 = fmaf(a0.w, e0, acc3);
					acc0 = fmaf(a1.x, e1, acc0); acc1 = fmaf(a1.y, e1, acc1); acc2 = fmaf(a1.z, e1, acc2); acc3 = fmaf(a1.w, e1, acc3);
					acc0 = fmaf(a2.x, e2, acc0); acc1 = fmaf(a2.y, e2, acc1); acc2 = fmaf(a2.z, e2, acc2); acc3 = fmaf(a2.w, e2, acc3);
					acc0 = fmaf(a3.x, e3, acc0); acc1 = fmaf(a3.y, e3, acc1); acc2 = fmaf(a3.z, e3, acc2); acc3 = fmaf(a3.w, e3, acc3);
				}
				for (; q < qmax; q++) {
					const float e0 = Ec[q * CWp];
					const float4 a0 = Au[q];
					acc0 = fmaf(a0.x, e0, acc0); acc1 = fmaf(a0.y, e0, acc1); acc2 = fmaf(a0.z, e0, acc2); acc3 = fmaf(a0.w, e0, acc3);
				}
				const float accv[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const uint32_t s = uq * 4 + k;
					if (s_len[s] && s_t[s] + 1 < s_len[s]) p.G[((uint64_t)s_off[s] + s_t[s]) * Lp + my_c] = accv[k];
				}
			}
		}
		__syncthreads();   // A_s and the slot state are rewritten at the top of the next step
	}
}

// =================================================================================================
// backward + posteriors
// =================================================================================================
template <int UB>
__global__ void __launch_bounds__(512, 1) backward_cluster_kernel(ClusterDpParams p) {
	constexpr int UQ = UB / 4;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const SmemLayout lay = smem_layout(p.L, p.CWp, UB, p.D);
	float* E_s = reinterpret_cast<float*>(smem_raw);                         // [L][CWp]  ET: E_s[c][j] = E[q0+j][c]
	float4* V_s = reinterpret_cast<float4*>(smem_raw + lay.e_bytes);          // [UQ][L] of float4
	double* k_ring = reinterpret_cast<double*>(smem_raw + lay.e_bytes + lay.a_bytes);   // [UB][D]
	double* base_s = k_ring + (size_t)UB * p.D;                               // [UB]
	double* sg_s = base_s + UB;                                               // [UB]
	double* lz_s = sg_s + UB;                                                 // [UB] logZ of the slot's utterance
	float* delta_s = reinterpret_cast<float*>(lz_s + 2 * UB);                 // [UB][D+1]  base_{t+d} - kappa*
	float* rsc_s = delta_s + (size_t)UB * (p.D + 1);                          // [UB][D+1]
	float* gmax_s = rsc_s + (size_t)UB * (p.D + 1);                           // [UB]
	int* key_s = reinterpret_cast<int*>(gmax_s + UB);                         // [UB]
	uint32_t* s_utt = reinterpret_cast<uint32_t*>(key_s + 2 * UB);
	uint32_t* s_off = s_utt + UB;
	uint32_t* s_len = s_off + UB;
	uint32_t* s_t = s_len + UB;
	uint32_t* s_ctl = s_t + UB;

	const uint32_t L = p.L, Lp = p.Lp, P = p.P, D = p.D, CWp = p.CWp;
	const bool clustered = p.CS > 1;
	const uint32_t rank = clustered ? cluster_ctarank() : 0;
	const uint32_t cl = blockIdx.x / p.CS;
	const uint32_t c0 = rank * p.CW;
	const uint32_t ncol = c0 < L ? min(p.CW, L - c0) : 0;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
	const uint32_t cj = tid % p.CWt, uq = tid / p.CWt;
	const bool mv_thread = (uq < UQ) && (cj < ncol);
	const uint32_t my_q = c0 + cj;                                            // the label this thread owns
	const uint32_t my_d = mv_thread ? my_q / P + 1 : 0xffffu;
	float* xmax = p.xmax + (size_t)cl * 2 * p.CS * UB;
	const uint32_t list_begin = p.cl_off[cl], list_end = p.cl_off[cl + 1];

	for (uint32_t i = tid; i < L * CWp; i += blockDim.x) {
		const uint32_t c = i / CWp, j = i % CWp;
		E_s[i] = (j < ncol) ? p.ET[(size_t)c * Lp + c0 + j] : 0.0f;
	}
	if (tid < UB) { s_len[tid] = 0; s_t[tid] = 0; s_utt[tid] = LAB_BAD; key_s[tid] = KEY_NEG_INF; }
	if (tid == 0) s_ctl[0] = list_begin;
	__syncthreads();

	for (uint32_t step = 0;; step++) {
		const uint32_t buf = step & 1;
		// ---- slot management (t runs from len-1 down to 0) ----
		if (tid == 0) {
			uint32_t next = s_ctl[0], any = 0;
			for (int s = 0; s < UB; s++) {
				if (s_len[s] && s_t[s] > 0) { s_t[s]--; any = 1; continue; }
				if (next < list_end) {
					const uint32_t utt = p.cl_list[next++];
					s_utt[s] = utt; s_off[s] = p.off[utt]; s_len[s] = p.off[utt + 1] - p.off[utt]; s_t[s] = s_len[s] - 1;
					lz_s[s] = p.logZ[utt]; any = 1;
				} else s_len[s] = 0;
			}
			s_ctl[0] = next; s_ctl[1] = any;
		}
		__syncthreads();
		if (!s_ctl[1]) break;
		// ---- [A] scales ----
		for (uint32_t s = warp; s < UB; s += n_warps) {
			if (s_len[s]) {
				const uint32_t t = s_t[s], d = lane + 1;
				const uint32_t numNext = min(s_len[s] - 1 - t, D);
				const uint64_t n = (uint64_t)s_off[s] + t;
				double kv = -DBL_MAX, bv = 0.0;
				if (d <= numNext) { kv = k_ring[s * D + (t + d) % D]; bv = p.bbase[n + d]; }
				const double kstar = warp_max_d(kv);
				const double base = numNext ? kstar + p.Mmax : 0.0;
				if (d <= D) {
					// v[(d,y)] = exp(lw_{t+d} + bbase_{t+d} - kappa*)
					delta_s[s * (D + 1) + d] = (d <= numNext) ? (float)(bv - kstar) : -INFINITY;
					rsc_s[s * (D + 1) + d] = (d <= t) ? (float)(base + p.m[n - d] + p.Mmax - lz_s[s]) : -INFINITY;
				}
				if (lane == 0) {
					base_s[s] = base;
					sg_s[s] = p.m[n] + base - lz_s[s];
					if (rank == 0) p.bbase[n] = base;
				}
			}
		}
		if (tid < UB) key_s[tid] = KEY_NEG_INF;
		__syncthreads();
		// ---- [B] every CTA gathers the full v vector of each slot from the next numNext frames ----
		uint32_t cmax = 0;
		for (int s = 0; s < UB; s++) if (s_len[s]) cmax = max(cmax, P * min(s_len[s] - 1 - s_t[s], D));
		for (uint32_t i = tid; i < (uint32_t)UQ * L; i += blockDim.x) {
			const uint32_t g4 = i / L, c = i % L;
			const uint32_t d = c / P + 1;
			float v[4];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t s = g4 * 4 + k;
				v[k] = 0.0f;
				if (s_len[s]) {
					const uint32_t numNext = min(s_len[s] - 1 - s_t[s], D);
					if (d <= numNext)
						v[k] = expf(p.G[((uint64_t)s_off[s] + s_t[s] + d) * Lp + c] + delta_s[s * (D + 1) + d]);
				}
			}
			V_s[(size_t)g4 * L + c] = make_float4(v[0], v[1], v[2], v[3]);
		}
		__syncthreads();
		// ---- [C] u[q] = sum_c E[q][c] v[c] for my label, my 4 slots; posteriors ----
		if (uq < UQ) {
			float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
			if (mv_thread && cmax) {
				const float* Ec = E_s + cj;
				const float4* Vu = V_s + (size_t)uq * L;
				uint32_t c = 0;
				for (; c + 4 <= cmax; c += 4) {
					const float e0 = Ec[(c + 0) * CWp], e1 = Ec[(c + 1) * CWp], e2 = Ec[(c + 2) * CWp], e3 = Ec[(c + 3) * CWp];
					const float4 a0 = Vu[c + 0], a1 = Vu[c + 1], a2 = Vu[c + 2], a3 = Vu[c + 3];
					acc0 = fmaf(a0.x, e0, acc0); acc1 = fmaf(a0.y, e0, acc1); acc2 = fmaf(a0.z, e0, acc2); acc3 = fmaf(a0.w, e0, acc3);
					acc0 = fmaf(a1.x, e1, acc0); acc1 = fmaf(a1.y, e1, acc1); acc2 = fmaf(a1.z, e1, acc2); acc3 = fmaf(a1.w, e1, acc3);
					acc0 = fmaf(a2.x, e2, acc0); acc1 = fmaf(a2.y, e2, acc1); acc2 = fmaf(a2.z, e2, acc2); acc3 = fmaf(a2.w, e2, acc3);
					acc0 = fmaf(a3.x, e3, acc0); acc1 = fmaf(a3.y, e3, acc1); acc2 = fmaf(a3.z, e3, acc2); acc3 = fmaf(a3.w, e3, acc3);
				}
				for (; c < cmax; c++) {
					const float e0 = Ec[c * CWp];
					const float4 a0 = Vu[c];
					acc0 = fmaf(a0.x, e0, acc0); acc1 = fmaf(a0.y, e0, acc1); acc2 = fmaf(a0.z, e0, acc2); acc3 = fmaf(a0.w, e0, acc3);
				}
			}
			const float accv[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t s = uq * 4 + k;
				float lw = -INFINITY;
				if (mv_thread && s_len[s]) {
					const uint32_t t = s_t[s];
					const uint32_t avail = P * min(t + 1, D);
					const uint64_t n = (uint64_t)s_off[s] + t;
					if (my_q < avail) {
						const bool tail = (t + 1 == s_len[s]);
						const float uu = tail ? 1.0f : accv[k];
						const float lu = logf(uu);
						lw = p.S[n * Lp + my_q] + lu;
						const float gamma = p.A[n * Lp + my_q] * expf(lu + (float)sg_s[s]);
						p.Dm[n * Lp + my_q] = ((p.node_lab[n] == my_q) ? 1.0f : 0.0f) - gamma;
						p.R[n * Lp + my_q] = (my_d <= t) ? expf(lw + rsc_s[s * (D + 1) + my_d]) : 0.0f;
						if (p.Uvec) p.Uvec[n * Lp + my_q] = uu;
					} else {
						p.Dm[n * Lp + my_q] = 0.0f;
						p.R[n * Lp + my_q] = 0.0f;
						if (p.Uvec) p.Uvec[n * Lp + my_q] = 0.0f;
					}
					p.G[n * Lp + my_q] = lw;     // log-domain w_t = S_t + log u_t, read back by all CTAs d frames earlier
				}
				const float wm = warp_max(lw);
				if (lane == 0 && wm > -INFINITY) atomicMax(&key_s[s], float_key(wm));
			}
		}
		__syncthreads();
		if (tid < UB) xmax[((size_t)buf * p.CS + rank) * UB + tid] = key_float(key_s[tid]);
		cluster_barrier(clustered);
		// ---- [D] scale of w_t: kappa_t = bbase_t + max_c lw_t[c] ----
		if (tid < UB && s_len[tid]) {
			float g = -INFINITY;
			for (uint32_t r = 0; r < p.CS; r++) g = fmaxf(g, xmax[((size_t)buf * p.CS + r) * UB + tid]);
			const double kap = base_s[tid] + (double)g;
			k_ring[tid * D + s_t[tid] % D] = kap;
			if (rank == 0) p.kappa[(uint64_t)s_off[tid] + s_t[tid]] = kap;
		}
		__syncthreads();
	}
}

// =================================================================================================
// host side: plan + launch
// =================================================================================================
bool plan_cluster_dp(uint32_t L, uint32_t D, int max_smem_optin, ClusterPlan* plan, int ub_cap) {
	// smallest cluster whose E slice + a 12..32-slot frame buffer fits the opt-in shared memory
	for (uint32_t CS = 1; CS <= 8; CS *= 2) {
		const uint32_t CW = (L + CS - 1) / CS;
		const uint32_t CWp = CW | 1;                              // odd row stride: conflict-free column reads
		if ((CS - 1) * CW >= L) continue;                         // an empty slice: L too small for this CS
		for (int UB : {32, 16, 12, 8, 4}) {
			if (UB > ub_cap) continue;
			const uint32_t CWt = (CW + 31) / 32 * 32;
			if (CWt * (UB / 4) > 512) continue;
			const SmemLayout s = smem_layout(L, CWp, UB, D);
			if (s.total + 1024 <= (size_t)max_smem_optin) {
				plan->CS = CS; plan->CW = CW; plan->CWp = CWp; plan->CWt = CWt; plan->UB = UB;
				plan->threads = max(CWt * (UB / 4), 64u); plan->smem = s.total;
				return true;
			}
		}
	}
	return false;
}

template <int UB>
static cudaError_t launch_one(bool backward, const ClusterDpParams& p, const ClusterPlan& plan, cudaStream_t s) {
	auto kern = backward ? backward_cluster_kernel<UB> : forward_cluster_kernel<UB>;
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
	if (e != cudaSuccess) return e;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(p.n_clusters * plan.CS); cfg.blockDim = dim3(plan.threads);
	cfg.dynamicSmemBytes = plan.smem; cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = plan.CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int UB>
static int max_clusters_one(const ClusterPlan& plan) {
	auto kern = forward_cluster_kernel<UB>;
	if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem) != cudaSuccess) return 0;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(plan.CS); cfg.blockDim = dim3(plan.threads); cfg.dynamicSmemBytes = plan.smem;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = plan.CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	int n = 0;
	if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int max_active_clusters(const ClusterPlan& plan) {
	switch (plan.UB) {
		case 4: return max_clusters_one<4>(plan);
		case 8: return max_clusters_one<8>(plan);
		case 12: return max_clusters_one<12>(plan);
		case 16: return max_clusters_one<16>(plan);
		default: return max_clusters_one<32>(plan);
	}
}

cudaError_t launch_cluster_dp(bool backward, const ClusterDpParams& p, const ClusterPlan& plan, cudaStream_t s) {
	if (!p.n_clusters) return cudaSuccess;
	switch (plan.UB) {
		case 4: return launch_one<4>(backward, p, plan, s);
		case 8: return launch_one<8>(backward, p, plan, s);
		case 12: return launch_one<12>(backward, p, plan, s);
		case 16: return launch_one<16>(backward, p, plan, s);
		default: return launch_one<32>(backward, p, plan, s);
	}
}

}  // namespace crfgpu
