// Viterbi recursion for LARGE phone sets with one state per phone (cfg5: 1024 phones, maxDur 30): the P x P table of cross-phone
// costs does not fit the shared memory of one CTA, and one CTA per utterance (viterbi_kernel) streams all of it from L2 for every
// frame of every utterance.  Here, as in the native no_dur training recursion (crf_dp_nodur.cu), the table is SLICED over a group of
// ceil(P / 32) CTAs -- CTA pt keeps the 32 target columns of its phone tile in shared memory for the whole launch -- and 16
// utterances advance in lock-step per group, so every table element read from shared memory serves two utterances per thread and
// the per-frame exchange is one all-gather of the kept costs (16 x P floats through L2) behind one group barrier.
//
// The arithmetic and every tie-break are those of viterbi_kernel (crf_kernels.cu) restricted to one state per phone -- the array
// restatement of CRF_ViterbiDecoder_StdSeg_NoSegTransFtr<CRF_ViterbiNode>::nStateDecode with the free-phone LM and no beam
// (.cpp:116-166 state update, :246-415 within-phone, :435-543 cross-phone, :976-1106 kept list, :2156-2171 final argmin,
// :2204-2349 traceback; DESIGN.md "Viterbi exactness"):
//   * cross-phone candidates are scanned in the kept-list order of the previous frame, strict '<' (first arrival wins);
//     the order is the identity or "increasing with phone g moved to the back", g in {0, 1};
//   * the arrival-order descriptors a[s] of viterbi_kernel's ring have the closed form a[0] = identity, a[s] = ((s-1)/D) & 1:
//     a[s] is the head of the list described by a[max(0, s-D)], which is 0 for the identity and for g = 1, and 1 for g = 0;
//   * the durations d >= 2 of node s only involve candidates of EARLIER start frames: their best is formed while the group barrier
//     of the frame is still collecting arrivals, so the loads of the candidate rings are off the chain.
#include "crf_kernels.cuh"

#include <cuda_runtime.h>
#include <math_constants.h>

namespace crfgpu {

namespace {

constexpr int UT = VITG_UT, PT = 32, NTHR = 256;
constexpr uint32_t DC = 8;
constexpr float VINF = 99999.0f;
static_assert(UT == 16 && (NTHR / 32) * 2 == UT, "warp w owns utterances 2w, 2w+1 of the batch");

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
	uint32_t v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
// kept-list descriptor helpers (closed form of viterbi_kernel's s_move ring, see the header)
__device__ __forceinline__ uint32_t arrival_desc(uint32_t s, uint32_t D) { return s == 0 ? 0xffu : (((s - 1) / D) & 1u); }
// order of the kept list of frame s (feeds the cross scan of frame s+1 and the final argmin)
__device__ __forceinline__ uint32_t kept_desc(uint32_t s, uint32_t D) { const uint32_t j = s + 1 >= D ? s + 1 - D : 0; return arrival_desc(j, D); }

}  // namespace

#define VTICK(i) do { if (timing) { const long long now_ = clock64(); tacc[i] += (unsigned long long)(now_ - tlast); tlast = now_; } } while (0)

size_t vitg_smem_bytes(uint32_t P) { return (size_t)((P + 31) / 32 * 32) * (PT + UT) * sizeof(float) + 16; }

__global__ void __launch_bounds__(NTHR, 1) viterbi_group_kernel(VitGroupParams p) {
	extern __shared__ __align__(16) float vg_smem[];
	__shared__ __align__(16) float tileT[PT][UT];
	__shared__ uint32_t s_utt[UT], s_off[UT], s_len[UT];
	const uint32_t P = p.P, D = p.D, Pk = (P + 31) / 32 * 32;
	float* crossS = vg_smem;                 // [Pk][32]: crossS[pp][l] = crossT[pp][y0 + l]
	float* xs = crossS + (size_t)Pk * PT;    // [Pk][UT]: kept costs of the previous frame, all phones of the 16 utterances
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t grp = blockIdx.x / p.npt, pt = blockIdx.x % p.npt, y0 = pt * PT, y = y0 + lane;
	const bool y_ok = y < P;
	for (uint32_t i = tid; i < Pk * PT; i += NTHR) {
		const uint32_t pp = i / PT, yy = y0 + (i % PT);
		crossS[i] = (pp < P && yy < P) ? p.crossT[(size_t)pp * P + yy] : 0.0f;
	}
	const float my_diag = y_ok ? p.negDiag[y] : 0.0f;
	uint32_t* ctr = p.ctr + grp;
	float* xch0 = p.xch + (size_t)grp * 2 * Pk * UT;
	uint32_t gstep = 0;      // arrivals of this CTA so far
	__syncthreads();
	const bool timing = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
	unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	long long tlast = timing ? clock64() : 0;

	for (uint32_t batch = grp; batch < p.n_batches; batch += p.n_groups) {
		if (gstep) {    // every CTA of the group has left the previous batch (its exchange buffers are free again)
			if (tid == 0) while (ld_acquire_u32(ctr) < gstep * p.npt) {}
			__syncthreads();
		}
		if (tid < UT) {
			const uint32_t u = p.slot_utt[(size_t)batch * UT + tid];
			s_utt[tid] = u;
			s_off[tid] = u != 0xffffffffu ? p.off[u] : 0;
			s_len[tid] = u != 0xffffffffu ? p.off[u + 1] - p.off[u] : 0;
		}
		__syncthreads();
		uint32_t Tmax = 0;
		for (int u = 0; u < UT; u++) Tmax = max(Tmax, s_len[u]);
		const uint32_t u0 = warp * 2;
		float wprev[2] = {0.0f, 0.0f};
		uint32_t slot = 0;      // s % D
		for (uint32_t s = 0; s < Tmax; s++) {
			// ---- durations >= 2 of node s (candidates of earlier start frames), longest first; and the duration-1 score ----
			float best[2], ns1[2]; int32_t bptr[2]; uint32_t bdur[2]; bool act[2];
			const uint32_t dmax = min(s + 1, D);
			// base pointers once per frame, 32-bit offsets inside the frame's block of scores / the utterance's ring; entries that are
			// not active read entry 0 of the arrays (harmless) and write nothing
			const float* sp[2]; const float2* ring[2];
#pragma unroll
			for (int i = 0; i < 2; i++) {
				const uint32_t u = u0 + i;
				act[i] = s < s_len[u] && y_ok;
				best[i] = CUDART_INF_F; bptr[i] = -1; bdur[i] = 0;      // the first (longest) duration always wins against +inf: "d == dmax ||"
				sp[i] = p.negS + (act[i] ? (size_t)(s_off[u] + s) * D * P + y : 0);
				ring[i] = p.cand + (act[i] ? (size_t)s_utt[u] * D * P + y : 0);
				ns1[i] = __ldg(sp[i]);
			}
			if (dmax >= 2 && (act[0] || act[1])) {
				// k = d - 1 runs from dmax - 1 down to 1; the candidates of start frame s - k sit in ring slot (slot - k) mod D.
				// The loads of DC durations of BOTH entries are issued before any is used.
				uint32_t k = dmax - 1;
				int32_t sl = (int32_t)slot - (int32_t)k; if (sl < 0) sl += (int32_t)D;
				for (; k >= DC; k -= DC) {
					float2 cv[2][DC]; float sv[2][DC];
#pragma unroll
					for (uint32_t j = 0; j < DC; j++) {
						int32_t slj = sl + (int32_t)j; if (slj >= (int32_t)D) slj -= (int32_t)D;
						const uint32_t o_r = (uint32_t)slj * P, o_s = (k - j) * P;
#pragma unroll
						for (int i = 0; i < 2; i++) { cv[i][j] = ring[i][o_r]; sv[i][j] = __ldg(sp[i] + o_s); }
					}
#pragma unroll
					for (int i = 0; i < 2; i++)
#pragma unroll
						for (uint32_t j = 0; j < DC; j++) {
							float ww = cv[i][j].x;
							if (ww < VINF) ww = ww + sv[i][j];
							if (ww < best[i]) { best[i] = ww; bptr[i] = __float_as_int(cv[i][j].y); bdur[i] = k - j + 1; }
						}
					sl += (int32_t)DC; if (sl >= (int32_t)D) sl -= (int32_t)D;
				}
				if (k >= 1) {      // k, ..., 1
					float2 cv[2][DC]; float sv[2][DC];
#pragma unroll
					for (uint32_t j = 0; j < DC; j++) {
						const bool on = j < k;
						int32_t slj = sl + (int32_t)j; if (slj >= (int32_t)D) slj -= (int32_t)D;
						const uint32_t o_r = on ? (uint32_t)slj * P : 0, o_s = on ? (k - j) * P : 0;
#pragma unroll
						for (int i = 0; i < 2; i++) { cv[i][j] = ring[i][o_r]; sv[i][j] = __ldg(sp[i] + o_s); }
					}
#pragma unroll
					for (int i = 0; i < 2; i++)
#pragma unroll
						for (uint32_t j = 0; j < DC; j++) {
							if (j < k) {
								float ww = cv[i][j].x;
								if (ww < VINF) ww = ww + sv[i][j];
								if (ww < best[i]) { best[i] = ww; bptr[i] = __float_as_int(cv[i][j].y); bdur[i] = k - j + 1; }
							}
						}
				}
			}
			VTICK(0);   // durations >= 2
			// ---- cross-phone candidates for segments starting at frame s ----
			float cw[2] = {0.0f, 0.0f}; int32_t cp[2] = {-1, -1};     // s == 0: lm_start weight 0 + arc weight 0 (:444-447)
			if (s > 0) {
				if (tid == 0) while (ld_acquire_u32(ctr) < gstep * p.npt) {}
				__syncthreads();
				VTICK(1);   // barrier wait
				{
					const float4* xv = reinterpret_cast<const float4*>(xch0 + (size_t)((s - 1) & 1) * Pk * UT);
					float4* xd = reinterpret_cast<float4*>(xs);
#pragma unroll 8
					for (uint32_t i = tid; i < Pk * (UT / 4); i += NTHR) xd[i] = __ldcg(xv + i);
				}
				__syncthreads();
				VTICK(2);   // staging
				const uint32_t g = kept_desc(s - 1, D);
				float pw0 = CUDART_INF_F, pw1 = CUDART_INF_F; int32_t q0 = -1, q1 = -1;
				const float* xw = xs + u0;          // the published costs already carry the reference's "+ 0.0f" (see the publish step)
				const float* cs = crossS + lane;
				constexpr int32_t TAG = 0x40000000;   // q = TAG | first phone of an 8-phone block: the index inside the block is resolved after the scan
				auto checked = [&](uint32_t pp) {
					if (pp != y) {     // free-phone LM, one state per phone: no arc to the same phone (:1332-1346)
						const float c = cs[(size_t)pp * PT];
						const float2 w = *reinterpret_cast<const float2*>(xw + (size_t)pp * UT);
						const float c0 = w.x + c, c1 = w.y + c;
						if (c0 < pw0) { pw0 = c0; q0 = (int32_t)pp; }
						if (c1 < pw1) { pw1 = c1; q1 = (int32_t)pp; }
					}
				};
				// blocks of 8 phones: the block minimum (a tree of FMNMX, no index bookkeeping) replaces the running minimum only if
				// strictly smaller, so the FIRST block that attains the final minimum is remembered and the first phone inside it that
				// attains it is found afterwards -- the same winner as the element-by-element scan with strict '<'
				auto range = [&](uint32_t a, uint32_t b) {
					uint32_t pp = a;
					for (; pp + 8 <= b; pp += 8) {
						float c0[8], c1[8];
#pragma unroll
						for (int j = 0; j < 8; j++) {
							const float c = cs[(size_t)(pp + j) * PT];
							const float2 w = *reinterpret_cast<const float2*>(xw + (size_t)(pp + j) * UT);
							c0[j] = w.x + c; c1[j] = w.y + c;
						}
						const float m0 = fminf(fminf(fminf(c0[0], c0[1]), fminf(c0[2], c0[3])), fminf(fminf(c0[4], c0[5]), fminf(c0[6], c0[7])));
						const float m1 = fminf(fminf(fminf(c1[0], c1[1]), fminf(c1[2], c1[3])), fminf(fminf(c1[4], c1[5]), fminf(c1[6], c1[7])));
						if (m0 < pw0) { pw0 = m0; q0 = TAG | (int32_t)pp; }
						if (m1 < pw1) { pw1 = m1; q1 = TAG | (int32_t)pp; }
					}
					for (; pp < b; pp++) {
						const float c = cs[(size_t)pp * PT];
						const float2 w = *reinterpret_cast<const float2*>(xw + (size_t)pp * UT);
						const float c0 = w.x + c, c1 = w.y + c;
						if (c0 < pw0) { pw0 = c0; q0 = (int32_t)pp; }
						if (c1 < pw1) { pw1 = c1; q1 = (int32_t)pp; }
					}
				};
				auto resolve = [&](float& pw, int32_t q, int which) -> int32_t {
					if (q < TAG) return q;
					const uint32_t b0 = (uint32_t)(q & ~TAG);
					int32_t r = -1; float v = pw;
#pragma unroll
					for (int j = 7; j >= 0; j--) {
						const float c = cs[(size_t)(b0 + j) * PT];
						const float w = xw[(size_t)(b0 + j) * UT + which];
						if (w + c == pw) { r = (int32_t)(b0 + j); v = w + c; }
					}
					pw = v;      // the winner's own value (fminf may have picked the other sign of a zero)
					return r;
				};
				// increasing phone order with g moved to the back; phones 0, 1 (possible g) and the CTA's own tile (possible target) checked
				if (g != 0u) checked(0);
				if (g != 1u && P > 1) checked(1);
				const uint32_t t_lo = max(2u, y0), t_hi = min(P, y0 + PT);
				if (t_lo > 2) range(2, min(t_lo, P));
				for (uint32_t pp = t_lo; pp < t_hi; pp++) checked(pp);
				if (t_hi < P) range(max(t_hi, 2u), P);
				if (g != 0xffu) checked(g);
				q0 = resolve(pw0, q0, 0); q1 = resolve(pw1, q1, 1);
				// within-phone: the self transition, taken only if strictly smaller than the cross candidate (first arrival wins)
				cw[0] = pw0; cp[0] = q0; cw[1] = pw1; cp[1] = q1;
#pragma unroll
				for (int i = 0; i < 2; i++) {
					const float n1 = wprev[i] + my_diag;
					if (cp[i] < 0 || n1 < cw[i]) { cw[i] = n1; cp[i] = (int32_t)y; }
				}
			}
			VTICK(3);   // scan
			// ---- node s: duration 1 joins the durations formed above; back pointers; the kept cost of the frame ----
#pragma unroll
			for (int i = 0; i < 2; i++) {
				const uint32_t u = u0 + i;
				if (act[i]) {
					if (D > 1) p.cand[((size_t)s_utt[u] * D + slot) * P + y] = make_float2(cw[i], __int_as_float(cp[i]));
					float w1 = cw[i];
					if (w1 < VINF) w1 = w1 + ns1[i];
					if (dmax == 1 || w1 < best[i]) { best[i] = w1; bptr[i] = cp[i]; bdur[i] = 1; }
					wprev[i] = best[i];
					const size_t n = (size_t)(s_off[u] + s) * P + y;
					p.bp[n] = bptr[i] < 0 ? (uint16_t)0xffff : (uint16_t)bptr[i];
					p.bd[n] = (uint8_t)bdur[i];
					if (s + 1 == s_len[u]) p.finalW[(size_t)s_utt[u] * P + y] = best[i];
				}
				tileT[lane][u] = wprev[i] + 0.0f;     // the cross-phone scan reads (kept cost + 0.0f) (the LM arc weight of the free-phone loop)
			}
			__syncthreads();
			// ---- publish the CTA's slice of the kept costs and arrive ----
			if (tid < PT * UT / 4)
				__stcg(reinterpret_cast<float4*>(xch0 + (size_t)(s & 1) * Pk * UT + (size_t)y0 * UT) + tid, reinterpret_cast<const float4*>(&tileT[0][0])[tid]);
			__syncthreads();
			if (tid == 0) { __threadfence(); atomicAdd(ctr, 1u); }
			VTICK(4);   // node update + publish
			gstep++;
			slot = slot + 1 == D ? 0 : slot + 1;
		}
	}
	if (timing) { for (int i = 0; i < 8; i++) p.dbg[i] = tacc[i]; p.dbg[8] = gstep; }
}

// final argmin over the kept list of the last frame (first in list order wins) and traceback, one warp per utterance
__global__ void __launch_bounds__(128) viterbi_group_traceback_kernel(VitGroupParams p) {
	const uint32_t u = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (u >= p.n_utt) return;
	const uint32_t P = p.P, D = p.D, off = p.off[u], T = p.off[u + 1] - off;
	if (T == 0) { if (lane == 0) { p.n_seg[u] = 0; p.cost[u] = VINF; } return; }
	const uint32_t g = kept_desc(T - 1, D);
	const float* fw = p.finalW + (size_t)u * P;
	float bw = VINF; uint32_t bi = 0xffffffffu;     // (cost, list position); the reference starts from 99999.0 with strict '<'
	for (uint32_t i = lane; i < P; i += 32) {
		const uint32_t ph = g == 0xffu ? i : (i + 1 == P ? g : (i < g ? i : i + 1));
		const float w = fw[ph];
		if (w < bw) { bw = w; bi = i; }
	}
	for (int o = 16; o > 0; o >>= 1) {
		const float ow = __shfl_xor_sync(0xffffffffu, bw, o);
		const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
		if (ow < bw || (ow == bw && oi < bi)) { bw = ow; bi = oi; }
	}
	if (lane != 0) return;
	uint32_t nseg = 0;
	if (bi != 0xffffffffu) {
		int cur = (int)(g == 0xffu ? bi : (bi + 1 == P ? g : (bi < g ? bi : bi + 1)));
		int end = (int)T - 1;
		uint32_t* ol = p.out_lab + off; uint32_t* od = p.out_dur + off; uint32_t* op = p.out_phn + off;
		while (end >= 0) {
			const uint32_t d = p.bd[(size_t)(off + end) * P + cur];
			const uint16_t prev = p.bp[(size_t)(off + end) * P + cur];
			const int start = end + 1 - (int)d;
			ol[nseg] = (uint32_t)cur; od[nseg] = d;
			op[nseg] = (start == 0 || (int)prev != cur) ? (uint32_t)cur : LAB_BAD;
			nseg++;
			if (start == 0) break;
			cur = (int)prev; end = start - 1;
		}
		for (uint32_t i = 0; i < nseg / 2; i++) {
			uint32_t a;
			a = ol[i]; ol[i] = ol[nseg - 1 - i]; ol[nseg - 1 - i] = a;
			a = od[i]; od[i] = od[nseg - 1 - i]; od[nseg - 1 - i] = a;
			a = op[i]; op[i] = op[nseg - 1 - i]; op[nseg - 1 - i] = a;
		}
	}
	p.n_seg[u] = nseg; p.cost[u] = bw;
}

int vitg_max_groups(uint32_t P) {
	int dev = 0, sms = 0, per_sm = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const size_t smem = vitg_smem_bytes(P);
	if (cudaFuncSetAttribute(viterbi_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, viterbi_group_kernel, NTHR, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return 0; }
	const uint32_t npt = (P + PT - 1) / PT;
	return (int)((uint32_t)(sms * per_sm) / npt);
}

cudaError_t launch_viterbi_group(const VitGroupParams& p, cudaStream_t s) {
	if (!p.n_utt) return cudaSuccess;
	cudaError_t e = cudaMemsetAsync(p.ctr, 0, sizeof(uint32_t) * p.n_groups, s);
	if (e != cudaSuccess) return e;
	VitGroupParams q = p;
	void* args[] = {&q};
	e = cudaLaunchCooperativeKernel((const void*)viterbi_group_kernel, dim3(p.n_groups * p.npt), dim3(NTHR), args, vitg_smem_bytes(p.P), s);
	if (e != cudaSuccess) return e;
	viterbi_group_traceback_kernel<<<(p.n_utt + 3) / 4, 128, 0, s>>>(q);
	return cudaGetLastError();
}

}  // namespace crfgpu
