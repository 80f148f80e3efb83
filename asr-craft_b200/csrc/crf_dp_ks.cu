// Tensor-core, cluster-resident lattice recursions (K2/K3/K4) for sm_100a -- CONTRACTION-SLICED variant (default where it fits).
//
// Reference semantics as crf_dp_tc.cu: CRF_StdSegStateNode::computeAlpha / computeBeta / computeExpF / computeAlphaSum
// (CRF/src/nodes/CRF_StdSegStateNode.cpp:135-186, 219-308, 343-438, 447-462), CRF_StdStateNode / CRF_StdNStateNode for max_dur == 1.
//
// crf_dp_tc.cu slices the OUTPUT labels of the per-frame product  D[L x 16] = E[L x L] * V[L x 16]  over the CTAs of a cluster: every
// CTA needs the whole vector V, so the step is  MMA (120 narrow MMAs, a third of them with the A operand streamed from shared memory
// because only the hi half of the slice fits tensor memory: 3.0 k cycles) -> epilogue -> ALL-GATHER of V.  Here the CONTRACTION index
// is sliced instead: CTA r keeps the columns E[:, slice r] of ALL L rows -- as ceil(L/128) M-tiles whose hi AND lo halves both fit
// tensor memory (cfg4: 5 tiles x 80 columns = 400 of 512 TMEM columns, 80 more for the accumulators) -- and multiplies them with ITS OWN
// slice of V, which its own lane threads produced in the previous step:
//   * per step 5 tiles x 5 k-steps x 3 = 75 tcgen05.mma, every A operand from TMEM (20 instead of 53 cycles): 1.5 k cycles;
//   * the partial products are REDUCE-SCATTERED tile by tile while the later tiles are still being multiplied: the lane threads read
//     their accumulator rows (tcgen05.ld) as soon as a tile's commit arrives and send each row's 4 slots straight to the owner of the
//     row with one remote store that counts its bytes on the owner's mbarrier (st.async, 16 B per thread and tile); the owner sums the
//     8 partials of its 80 rows and runs the usual epilogue (log, + score, + scale, exp, bf16 hi/lo split) straight into its local MMA
//     operand buffer -- no vector ever has to be gathered, nothing is staged, no copy-issuing warp sits in the chain;
//   * only the 64 partial sums per CTA that the scale bookkeeping needs travel to every peer (256 B, off the critical path).
// No shared-memory E half (100 KB freed), no operand fetch from shared memory in the product.  Scales, slot refill, masking and the
// lane-thread arithmetic are those of crf_dp_tc.cu.  Geometries whose tiles do not fit tensor memory (more than 640 padded labels)
// stay on crf_dp_tc.cu.
#include "crf_kernels.cuh"
#include "tc05.cuh"

#include <cfloat>
#include <cstdio>
#include <cstdlib>

namespace crfgpu {

using namespace tc05;

namespace {

constexpr int UB = TC_DP_SLOTS;        // slots per cluster == MMA N
constexpr int DMAX = 32;               // scale rings (power of two); max_dur < DMAX
constexpr int LANE_WARPS = 16;         // 4 per TMEM lane quadrant, each owning SPT slots of its 32 rows
constexpr int SPT = UB / (LANE_WARPS / 4);
constexpr int NBK = 4;                 // bookkeeping warps, each owning SPW slots with LPS lanes per slot
constexpr int SPW = UB / NBK, LPS = 32 / SPW;
constexpr int MMA_WARP = LANE_WARPS, COMM_WARP = LANE_WARPS + 1, BK_WARP0 = LANE_WARPS + 2;
constexpr int N_THREADS = (LANE_WARPS + 2 + NBK) * 32;
constexpr int BAR_ALL = N_THREADS, BAR_LANES_1 = (LANE_WARPS + 1) * 32, BAR_LANES_2 = (LANE_WARPS + 2) * 32;   // lane threads + one / both service warps
constexpr int MT_MAX = 5;                // M-tiles of 128 label rows (plan: MT * (CW + 16) <= 512 TMEM columns)
constexpr uint32_t PS_BYTES = 256;       // [4 lane quadrants][16 slots] partial sums of one CTA
static_assert(SPT == 4, "the partial-sum butterfly below is written for 4 slots per thread");

struct Slot {
	uint32_t utt, off, len, t;     // utt == LAB_BAD: idle
	double lz;                     // logZ of the utterance (backward)
	double sc;                     // log scale of THIS frame (forward rho_t, backward base_t), stashed with the schedule entry when it is
	                               // derived: step [1] must not read it back from ring_b, because a slot that was refilled in between has
	                               // already stored the NEW utterance's scale of frame 0 there -- the same ring index whenever
	                               // (len - 1) % DMAX == 0 (that aliasing made logZ of such utterances wrong: 1.5 % on the cfg4 bench shard)
};

// what the lane threads need to know about one slot in one step (published by the bookkeeping warp)
struct LaneCtl {
	uint32_t flags;    // 1: gathered frame active, 2: produced frame active, 4: forward: keep G of the gathered frame | backward: gathered frame is the tail
	uint32_t crow;     // element index (frame * Lp) of the gathered frame
	uint32_t nrow;     // ... of the produced frame
	uint32_t cn;       // frame index of the gathered frame
	uint32_t ct;       // t of the gathered frame
	uint32_t nt;       // forward: t of the produced frame | backward: numNext of the produced frame
	uint32_t navail;   // forward: labels available in the produced frame | backward: in the gathered frame
	uint32_t pad;
};

struct Ctl {
	uint64_t gather[2], psg[2], mma_bar[5];   // gather: partial products of my slice arrived | psg: partial sums of every CTA arrived | one commit per M-tile
	uint32_t tmem, pad;
	uint32_t any[2][NBK];
	// published per step, double-buffered by step parity (the bookkeeping warp runs ahead of the lane threads)
	LaneCtl lc[2][UB];
	float delta[2][UB][DMAX + 1];  // log-scale correction of duration block d for the vector being produced
	float rsc[2][UB][DMAX + 1];    // backward: scale of R for duration d
	float sg[2][UB];               // backward: rho_t + base_t - logZ
	// private to the bookkeeping warp
	Slot fr[4][UB];                // fr[j & 3] = F(j), the frame produced in step j (gathered in step j+1)
	double ring_a[UB][DMAX];       // forward: ghat_t = log sum_c exp(alpha_t[c]) | backward: bl_t = base_t + log sum_c v_t[c]
	double ring_b[UB][DMAX];       // forward: rho_t                              | backward: base_t
	float ring_af[UB][DMAX];       // ring_a narrowed to float: the scale bounds are searched in fp32 (a bound may be off by an ulp of its
	                               // magnitude -- entries <= 1.001 instead of <= 1 --, what matters is that the chosen scale is then used exactly)
	float pre_sm[2][UB][DMAX];     // [j & 1]: forward smaxd[frame of F(j)][d-1] | backward smaxd[frame of F(j) + d][d-1]
	double pre_rho[UB][DMAX + 1];  // backward: rho_{t-d}, d = 0..D, of the frame about to be gathered
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
	uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int COUNT> __device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(COUNT) : "memory"); }
template <int COUNT> __device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(COUNT) : "memory"); }
// named barriers: 1 = scales/schedule of the step published (bookkeeping -> everyone), 2 = my slice of the new vector written
// (lanes -> MMA and communication warps), 3 = partial products of my slice arrived from every peer (communication warp -> lanes)
constexpr int BAR_SCALES = 1, BAR_TILE = 2;

// reductions over the LPS lanes that share a slot (lane = slot + SPW*h); executed by the whole warp
__device__ __forceinline__ double slot_max_unused(double v) {
#pragma unroll
	for (int o = SPW; o < 32; o <<= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ float slot_maxf(float v) {
#pragma unroll
	for (int o = SPW; o < 32; o <<= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ float slot_sum(float v) {
#pragma unroll
	for (int o = SPW; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

__device__ __forceinline__ Slot next_frame(const Slot& s, bool bwd) {
	Slot n = s;
	if (bwd) n.t = s.t - 1; else n.t = s.t + 1;
	return n;
}
__device__ __forceinline__ bool has_next(const Slot& s, bool bwd) { return s.utt != LAB_BAD && (bwd ? s.t > 0 : s.t + 1 < s.len); }

}  // namespace

size_t ks_dp_ctl_bytes() { return sizeof(Ctl); }

#define TICK(var) const long long var = timing ? clock64() : 0
#define TACC(slot, a, b) do { if (timing) tacc[slot] += (unsigned long long)((b) - (a)); } while (0)

template <bool BWD>
__global__ void __launch_bounds__(N_THREADS, 1) dp_ks_kernel(KsDpParams p) {
	extern __shared__ __align__(1024) unsigned char smem[];
	const uint32_t L = p.L, Lp = p.Lp, P = p.P, D = p.D, CS = p.CS, CW = p.CW, K = p.K;
	const uint32_t MT = p.MT, KS = CW / 16;
	const uint32_t CH = CW * 64;                           // one slice x 16 slots: fp32 partial products [4 slot quads][CW rows][4], or the bf16 hi/lo operand tile
	unsigned char* recvb = smem + p.recv_off;              // [2][CS][CH] partial products of MY slice by source CTA (double-buffered: a peer may be one step ahead)
	unsigned char* vbuf = smem + p.vbuf_off;               // [CH] my slice of the frame vector as the MMA's B operand (hi/lo, K-major)
	unsigned char* psb = smem + p.ps_off;                  // [2][CS][PS_BYTES] partial sums of every CTA's slice
	Ctl* ctl = reinterpret_cast<Ctl*>(smem + p.ctl_off);
	const uint32_t E_COL = MT * 16;                        // accumulators of tile i at columns [16 i, 16 i + 16); E tiles behind them

	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t rank = CS > 1 ? cluster_ctarank() : 0, cl = blockIdx.x / CS;
	const uint32_t c0 = rank * CW;
	// lane warps have TWO identities.  Reading accumulators (tcgen05.ld) ties a warp to the TMEM lane quadrant warp % 4: (q4, ssub) = the 32
	// accumulator rows and the 4 slot columns it scatters.  The epilogue reads its sums from shared memory, so its rows are free to choose:
	// the RW = ceil(CW / 32) row groups of the 4 slot quads are dealt to warps 0 .. 4 RW - 1 in order, which spreads the busy warps evenly
	// over the four SM sub-partitions (cfg4: 12 busy warps, 3 per sub-partition; tied to the quadrants it was 4 + 4 + 4 + 0 and the
	// epilogue is bound by instruction issue)
	const uint32_t q4 = warp & 3, ssub = (warp >> 2) & 3;
	const uint32_t RW = (CW + 31) / 32;
	const uint32_t sub = warp / RW, rgrp = warp - sub * RW;    // epilogue: slot quad, row group
	const uint32_t srow = q4 * 32 + lane;                      // accumulator row inside a tile
	const uint32_t row = rgrp * 32 + lane, c = c0 + row;       // epilogue: row of my slice, label
	const bool lane_thread = warp < LANE_WARPS;
	const bool epi_warp = lane_thread && sub < 4;              // this warp has epilogue work
	const bool in_tile = epi_warp && row < CW;                 // this thread owns a row of the slice tile
	const bool row_valid = in_tile && c < L;                   // ... that is a real label
	const uint32_t my_d = row_valid ? c / P + 1 : 0xffffu;
	const uint32_t inv20 = ((1u << 20) + CW - 1) / CW;         // R / CW == (R * inv20) >> 20 for R < 1024 (no integer division in the step loop)
	const float* Msrc = BWD ? p.E : p.ET;                  // rows = my labels, columns = the contracted label
	const bool timing = p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == MMA_WARP || warp == COMM_WARP || warp == BK_WARP0);
	unsigned long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // cycle counters, kept in registers and flushed once

	// ---------------------------------------------------------------- setup
	if (tid == 0) {
		mbar_init(&ctl->gather[0], 1); mbar_init(&ctl->gather[1], 1); mbar_init(&ctl->psg[0], 1); mbar_init(&ctl->psg[1], 1);
		for (int i = 0; i < MT_MAX; i++) mbar_init(&ctl->mma_bar[i], 1);
		fence_mbar_init();
	}
	if (warp == MMA_WARP) tmem_alloc(&ctl->tmem, p.tmem_cols);
	for (uint32_t i = tid; i < 2 * CS * PS_BYTES / 4; i += N_THREADS) reinterpret_cast<float*>(psb)[i] = 0.0f;   // row groups without a warp stay 0
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem = ctl->tmem;
	if (lane_thread) {
		// tile i, k-step ks of my contraction slice: row R = 128 i + my lane of E (backward) / E^T (forward), 16 contracted labels from
		// c0 + 16 ks on; hi and lo halves both into TENSOR MEMORY (8 packed columns each per k-step).  The four warps of a lane quadrant
		// share the (tile, k-step) pairs.
		for (uint32_t pr = ssub; pr < MT * KS; pr += 4) {
			const uint32_t i = pr / KS, ks = pr % KS, R = i * 128 + srow;
			float x[16];
			const float* src = Msrc + (size_t)R * Lp + c0 + ks * 16;
#pragma unroll
			for (int j = 0; j < 16; j++) x[j] = (R < L && c0 + ks * 16 + j < L) ? __ldg(src + j) : 0.0f;
			const float (&x0)[8] = *reinterpret_cast<const float (*)[8]>(&x[0]);
			const float (&x1)[8] = *reinterpret_cast<const float (*)[8]>(&x[8]);
			uint4 h0, l0, h1, l1;
			split8(x0, h0, l0); split8(x1, h1, l1);
			const uint32_t rh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w}, rl[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
			const uint32_t col = tmem + ((q4 * 32u) << 16) + E_COL + i * CW + ks * 8;
			tmem_st8(col, rh);
			tmem_st8(col + CW / 2, rl);
		}
		tmem_st_wait();
	}
	fence_proxy_async_smem();
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	if (CS > 1) cluster_sync_all();      // every CTA's mbarriers exist before the first remote completion can arrive

	if (warp >= BK_WARP0) {
		// =====================================================================================================
		// bookkeeping warp.  Iteration j (j = -1, 0, 1, ...) runs during step j of the other warps and
		//   [1] turns the partial sums gathered in step j into the exact log-sum of F(j-1),
		//   [2] derives the scale of F(j+1) and publishes everything the lanes need in step j+1,
		//   [3] schedules F(j+2) and fetches its score maxima.
		// NBK such warps, each with its own SPW slots; lane = local slot + SPW*h: the LPS lanes of a slot share the duration loops.
		// =====================================================================================================
		const uint32_t bw = warp - BK_WARP0, slot = bw * SPW + (lane & (SPW - 1)), h = lane / SPW;
		// every slot works through its own utterance list (the host balances the lists over all slots of all clusters)
		const uint32_t list_end = p.cl_off[cl * UB + slot + 1];
		const Slot idle{LAB_BAD, 0, 0, 0, 0.0, 0.0};
		uint32_t list_next = p.cl_off[cl * UB + slot];
		auto refill = [&](bool want, Slot& out) {
			if (want && h == 0 && list_next < list_end) {
				const uint32_t utt = p.cl_list[list_next++];
				out.utt = utt; out.off = p.off[utt]; out.len = p.off[utt + 1] - p.off[utt]; out.t = BWD ? out.len - 1 : 0;
				if (BWD) out.lz = p.logZ[utt];
			}
		};
		// what the scale recursions read from global memory (score maxima of a frame; backward: rho of the frames before it) is
		// loaded into registers right after the frame is scheduled and parked in shared memory only after the current iteration's
		// scale work, so the latency never shows (lane = duration index)
		float pf_sm[SPW]; double pf_rho[SPW];
		auto fetch_sm = [&](uint32_t fidx) {
#pragma unroll
			for (int s = 0; s < SPW; s++) {
				const Slot f = ctl->fr[fidx & 3][bw * SPW + s];
				pf_sm[s] = -INFINITY;
				if (f.utt == LAB_BAD) continue;
				if (!BWD) { if (lane < D) pf_sm[s] = __ldg(p.smaxd + ((size_t)f.off + f.t) * D + lane); }
				else { const uint32_t d = lane + 1; if (d <= min(f.len - 1 - f.t, D)) pf_sm[s] = __ldg(p.smaxd + ((size_t)f.off + f.t + d) * D + d - 1); }
			}
		};
		auto fetch_rho = [&](uint32_t fidx) {
#pragma unroll
			for (int s = 0; s < SPW; s++) {
				const Slot f = ctl->fr[fidx & 3][bw * SPW + s];
				pf_rho[s] = 0.0;
				if (BWD && f.utt != LAB_BAD && lane <= min(f.t, D)) pf_rho[s] = __ldg(p.m + (size_t)f.off + f.t - lane);
			}
		};
		auto park_sm = [&](uint32_t fidx) {
#pragma unroll
			for (int s = 0; s < SPW; s++) ctl->pre_sm[fidx & 1][bw * SPW + s][lane] = pf_sm[s];
		};
		auto park_rho = [&]() {
#pragma unroll
			for (int s = 0; s < SPW; s++) ctl->pre_rho[bw * SPW + s][lane] = pf_rho[s];
		};
		// F(-1) idle, F(0) = the first UB utterances of the list at their first frame (iteration -1 schedules F(1))
		{
			Slot f0 = idle;
			refill(true, f0);
			if (h == 0) { ctl->fr[3][slot] = idle; ctl->fr[0][slot] = f0; }
			__syncwarp();
			fetch_sm(0); park_sm(0);
			if (BWD) { fetch_rho(0); park_rho(); }
			__syncwarp();
		}
		for (int j = -1;; j++) {
			const uint32_t pb = (uint32_t)j & 1, nb = pb ^ 1;
			TICK(b0);
			const Slot fm = ctl->fr[(j - 1) & 3][slot], f0 = ctl->fr[j & 3][slot], f1 = ctl->fr[(j + 1) & 3][slot];
			// ---- [3] (data-independent, so first) schedule F(j+2) and start fetching its maxima / the forward scales of F(j+1) ----
			{
				Slot f2 = idle;
				const bool cont = has_next(f1, BWD);
				if (cont) f2 = next_frame(f1, BWD);
				refill(!cont, f2);
				if (h == 0) ctl->fr[(j + 2) & 3][slot] = f2;
				__syncwarp();
				fetch_sm((uint32_t)(j + 2));
				if (BWD) fetch_rho((uint32_t)(j + 1));
			}
			TICK(b0b); TACC(6, b0, b0b);
			if (j >= 1) mbar_wait(&ctl->psg[pb], ((j - 1) >> 1) & 1);
			TICK(b1); TACC(4, b0b, b1);
			// ---- [1] exact sum of the frame gathered in this step ----
			if (j >= 1) {
				float vsum = 0.0f;
				if (fm.utt != LAB_BAD) {
					const unsigned char* base = psb + (size_t)pb * CS * PS_BYTES;
					for (uint32_t r = h; r < CS; r += LPS)
#pragma unroll
						for (int q = 0; q < 4; q++) vsum += *reinterpret_cast<const float*>(base + r * PS_BYTES + (q * 16 + slot) * 4);
				}
				vsum = slot_sum(vsum);
				if (fm.utt != LAB_BAD && h == 0) {
					if (!BWD) {
						const bool last = fm.t + 1 == fm.len;
						// the log scales are references, not results: only logZ needs the accurate logarithm
						const double ghat = fm.sc + (last ? log((double)vsum) : (double)__logf(vsum));
						ctl->ring_a[slot][fm.t & (DMAX - 1)] = ghat; ctl->ring_af[slot][fm.t & (DMAX - 1)] = (float)ghat;
						if (last && rank == 0) p.logZ[fm.utt] = ghat;               // computeAlphaSum (:447-462)
					} else {
						const bool tail = fm.t + 1 == fm.len;
						const double bl = tail ? 0.0 : fm.sc + (double)__logf(vsum);
						ctl->ring_a[slot][fm.t & (DMAX - 1)] = bl; ctl->ring_af[slot][fm.t & (DMAX - 1)] = (float)bl;
					}
				}
				__syncwarp();
			}
			// ---- [2] scale of F(j+1); F(j) (same utterance, one frame earlier in processing order) is still in flight ----
			const float* sm0 = ctl->pre_sm[pb][slot];      // maxima for F(j)
			const float* sm1 = ctl->pre_sm[nb][slot];      // maxima for F(j+1)
			const bool a1 = f1.utt != LAB_BAD, a0 = f0.utt != LAB_BAD;
			if (!BWD) {
				const uint32_t t1 = f1.t;
				const float Mmaxf = (float)p.Mmax;
				float rhof = -INFINITY, rt = -INFINITY;
				const bool inflight = a1 && t1 > 0;            // F(j) is frame t1-1 of the same utterance
				const uint32_t t0 = t1 - 1;
				// tight bound of the in-flight frame t1-1 from exact sums, + log(label count) bounds its log-sum
				// (shuffles stay outside the per-slot conditions: every lane of the warp must execute them)
				if (inflight) {
					for (uint32_t d = 1 + h; d <= min(t0, D); d += LPS) rt = fmaxf(rt, sm0[d - 1] + Mmaxf + ctl->ring_af[slot][(t0 - d) & (DMAX - 1)]);
					if (h == 0 && t0 < D) rt = fmaxf(rt, sm0[t0]);
				}
				rt = slot_maxf(rt);
				if (inflight) {
					const float ub0 = rt + __logf((float)(P * min(t0 + 1, D)));
					if (h == 0) rhof = sm1[0] + Mmaxf + ub0;
					for (uint32_t d = 2 + h; d <= min(t1, D); d += LPS) rhof = fmaxf(rhof, sm1[d - 1] + Mmaxf + ctl->ring_af[slot][(t1 - d) & (DMAX - 1)]);
				}
				if (a1 && h == 0 && t1 < D) rhof = fmaxf(rhof, sm1[t1]);       // d == t1+1: the segment starts the utterance, alpha = S
				rhof = slot_maxf(rhof);
				const double rho = (double)rhof;
				if (a1) {
					for (uint32_t d = 1 + h; d <= D; d += LPS) {
						float dl = -INFINITY;
						if (d <= t1) dl = (float)(p.Mmax + ctl->ring_b[slot][(t1 - d) & (DMAX - 1)] - rho);
						else if (d == t1 + 1) dl = (float)(-rho);
						ctl->delta[nb][slot][d] = dl;
					}
				}
				__syncwarp();
				if (a1 && h == 0) {
					ctl->ring_b[slot][t1 & (DMAX - 1)] = rho;
					ctl->fr[(j + 1) & 3][slot].sc = rho;
					if (rank == 0) p.m[(size_t)f1.off + t1] = rho;
				}
			} else {
				const uint32_t t1 = f1.t;
				const uint32_t nn1 = a1 ? min(f1.len - 1 - t1, D) : 0;
				const float Mmaxf = (float)p.Mmax;
				float sigf = -INFINITY, st = -INFINITY;
				// frame t1+1 = F(j): bound of its bl from the exact bl of the frames behind it
				const uint32_t nn0 = nn1 ? min(f0.len - 1 - f0.t, D) : 0;
				for (uint32_t d = 1 + h; d <= nn0; d += LPS) st = fmaxf(st, sm0[d - 1] + ctl->ring_af[slot][(t1 + 1 + d) & (DMAX - 1)]);
				st = slot_maxf(st);
				if (nn1) {
					const float ubl0 = nn0 ? Mmaxf + st + __logf((float)(P * nn0)) : 0.0f;   // tail frame: S + beta = S exactly
					if (h == 0) sigf = sm1[0] + ubl0;
					for (uint32_t d = 2 + h; d <= nn1; d += LPS) sigf = fmaxf(sigf, sm1[d - 1] + ctl->ring_af[slot][(t1 + d) & (DMAX - 1)]);
				}
				sigf = slot_maxf(sigf);
				const double sigma = (double)sigf;
				if (a1)
					for (uint32_t d = 1 + h; d <= D; d += LPS)
						ctl->delta[nb][slot][d] = (d <= nn1) ? (float)(ctl->ring_b[slot][(t1 + d) & (DMAX - 1)] - sigma) : -INFINITY;
				// posterior scales of F(j), gathered (and turned into posteriors) in step j+1
				if (a0) {
					const double base0 = f0.sc;
					if (h == 0) {
						ctl->sg[nb][slot] = (float)(ctl->pre_rho[slot][0] + base0 - f0.lz);
						if (rank == 0) p.bbase[(size_t)f0.off + f0.t] = base0;
					}
					for (uint32_t d = 1 + h; d <= D; d += LPS)
						ctl->rsc[nb][slot][d] = (d <= f0.t) ? (float)(base0 + ctl->pre_rho[slot][d] + p.Mmax - f0.lz) : -INFINITY;
				}
				__syncwarp();
				if (a1 && h == 0) {
					const double base1 = nn1 ? p.Mmax + sigma : 0.0;   // tail: beta = 0 (setTailBeta)
					ctl->ring_b[slot][t1 & (DMAX - 1)] = base1;
					ctl->fr[(j + 1) & 3][slot].sc = base1;
				}
			}
			if (h == 0) {
				LaneCtl lc{};
				lc.flags = (a0 ? 1u : 0u) | (a1 ? 2u : 0u);
				if (a0) {
					lc.cn = f0.off + f0.t; lc.crow = lc.cn * Lp; lc.ct = f0.t;
					if (!BWD) { if (f0.t + 1 < f0.len) lc.flags |= 4u; }
					else { if (f0.t + 1 == f0.len) lc.flags |= 4u; lc.navail = P * min(f0.t + 1, D); }
				}
				if (a1) {
					lc.nrow = (f1.off + f1.t) * Lp;
					if (!BWD) { lc.nt = f1.t; lc.navail = P * min(f1.t + 1, D); }
					else lc.nt = min(f1.len - 1 - f1.t, D);
				}
				ctl->lc[nb][slot] = lc;
			}
			const uint32_t any_mine = __ballot_sync(0xffffffffu, a0 || a1);
			if (lane == 0) ctl->any[nb][bw] = any_mine ? 1u : 0u;
			__syncwarp();
			TICK(b2); TACC(5, b1, b2);
			// all warps meet here once per step (the lanes arrive when they finished the previous step's slice), so every warp
			// reads the same "anything left" flags and the loops end together
			bar_sync<BAR_ALL>(BAR_SCALES);
			uint32_t any = 0;
#pragma unroll
			for (int w = 0; w < NBK; w++) any |= ctl->any[nb][w];
			if (!any) break;
			park_sm((uint32_t)(j + 2));      // overwrites the maxima of F(j), consumed above
			if (BWD) park_rho();             // ... and the forward scales of F(j)
			__syncwarp();
		}
		if (timing) for (int i = 4; i < 7; i++) p.dbg[i] = tacc[i];   // bookkeeping warp 0
	} else {
		// =====================================================================================================
		// steps
		// =====================================================================================================
		uint32_t it = 0;
		for (;; it++) {
			const uint32_t pb = it & 1, nb = pb ^ 1;
			TICK(t0);
			bar_sync<BAR_ALL>(BAR_SCALES);
			{
				uint32_t any = 0;
#pragma unroll
				for (int w = 0; w < NBK; w++) any |= ctl->any[pb][w];
				if (!any) break;
			}
			unsigned char* myps = psb + ((size_t)nb * CS + rank) * PS_BYTES;
			unsigned char* recv_pb = recvb + (size_t)pb * CS * CH;
			// tile order of this step: CTA r starts with the tile that holds the first rows of slice r + 1, so that at any moment the CTAs of
			// the cluster push to DIFFERENT destinations and every destination receives a steady trickle instead of 7 chunks at once
			const uint32_t i0 = CS > 1 ? ((rank + 1) % CS) * CW / 128 : 0;
			if (warp == MMA_WARP) {
				// ===================== MMA warp: the product of my contraction slice, one commit per M-tile =====================
				if (it > 0) {
					// my slice of the vector was completed by my own lane threads (BAR_TILE at the end of the previous step)
					tc_fence_after();
					if (elect_one()) {
						// everything the issue loop updates is defined inside the elected region: with exactly one active lane the
						// compiler keeps descriptors and counters in uniform registers (UIADD3 + UTCHMMA, no R2UR per MMA)
						const uint32_t idesc = idesc_bf16_f32(128, UB, false, false);
						const uint64_t v0 = smem_desc(smem_u32(vbuf), 512, 128);
						uint32_t i = i0;
						for (uint32_t jj = 0; jj < MT; jj++) {
							const uint32_t dt = tmem + i * 16;
							uint32_t ta = tmem + E_COL + i * CW;
							uint64_t vhi = v0;
							bool acc = false;
							for (uint32_t k2 = 0; k2 < KS; k2++) {
								mma_ts(dt, ta, vhi, idesc, acc);                        // E_hi * V_hi
								mma_ts(dt, ta, vhi + 16, idesc, true);                  // E_hi * V_lo   (lo tile = +256 bytes)
								mma_ts(dt, ta + CW / 2, vhi, idesc, true);              // E_lo * V_hi
								acc = true; vhi += 64; ta += 8;                         // next k-step: +1024 bytes, +8 TMEM columns
							}
							mma_commit(&ctl->mma_bar[jj]);                              // the lane threads pick tile jj up as soon as it is complete
							i = (i + 1 == MT) ? 0 : i + 1;
						}
					}
					__syncwarp();
					TICK(m2); TACC(1, t0, m2);
				}
				bar_sync<BAR_LANES_2>(BAR_TILE);                        // my slice of the new vector (local MMA operand) is complete
				if (timing) tacc[11]++;
			} else if (warp == COMM_WARP) {
				// ===================== communication warp: reduce-scatter of the partial products, tile by tile =====================
				if (it > 0) {
					// slice r' of my partial products goes to CTA r' (its receive buffer [pb][my rank]) as soon as every tile that holds rows
					// of it has been stored.  The send buffer is double-buffered: I overwrite buffer pb again two steps on, after every peer's
					// NEXT push has reached me -- which it issued only after this push had arrived there.
					// the partial products of my slice arrive as remote stores of the peers' lane threads, each counting its bytes on gather[pb]
					if (lane == 0) mbar_arrive_expect_tx(&ctl->gather[pb], CS * CH);
					TICK(c1); TACC(0, t0, c1);
				}
				TICK(m4);
				bar_sync<BAR_LANES_2>(BAR_TILE);                        // the partial sums of my slice are complete
				// the only thing every CTA needs from every other: 64 partial sums for the scale bookkeeping
				if (lane == 0) mbar_arrive_expect_tx(&ctl->psg[nb], (CS - 1) * PS_BYTES);
				if (lane < CS && lane != rank) {
					const uint32_t dst = mapa(smem_u32(myps), lane), rbar = mapa(smem_u32(&ctl->psg[nb]), lane);
					asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					             ::"r"(dst), "r"(smem_u32(myps)), "r"(PS_BYTES), "r"(rbar) : "memory");
				}
				TICK(m5); TACC(3, m4, m5);
			} else {
				// ===================== lane threads: row `row` of the slice, slots sub*SPT .. sub*SPT+SPT-1 =====================
				TICK(l1); TACC(7, t0, l1);
				float g[SPT] = {0.0f, 0.0f, 0.0f, 0.0f};
				if (it > 0) {
					// my accumulator rows, tile by tile as the MMA warp completes them (row R = 128 i + row of the full product, my 4 slots),
					// stored by destination slice: [slot quad][row in slice][4 slots] so that consecutive lanes write consecutive 16 bytes
					uint32_t i = i0;
					const uint32_t gbar = smem_u32(&ctl->gather[pb]);
					for (uint32_t jj = 0; jj < MT; jj++) {
						mbar_wait(&ctl->mma_bar[jj], (it - 1) & 1);
						tc_fence_after();
						float pa[4];
						tmem_ld4(tmem + ((q4 * 32u) << 16) + i * 16 + ssub * SPT, pa);
						tmem_ld_wait();
						const uint32_t R = i * 128 + srow;
						if (R < K) {
							// straight from the registers into the owner's receive buffer [pb][my rank]: a remote store that counts its 16
							// bytes on the owner's mbarrier (st.async); no staging buffer, no proxy fence, no hand-over to a copy-issuing warp
							const uint32_t rd = (R * inv20) >> 20, rr = R - rd * CW;
							// (the rows of my own slice take the same route, to my own buffer and barrier: one completion covers everything the
							// epilogue threads -- which are not the threads that stored -- are about to read)
							unsigned char* lp = recv_pb + rank * CH + (ssub * CW + rr) * 16;
							const uint32_t ra = mapa(smem_u32(lp), rd), rb = mapa(gbar, rd);
							asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
							             ::"r"(ra), "f"(pa[0]), "f"(pa[1]), "f"(pa[2]), "f"(pa[3]), "r"(rb) : "memory");
						}
						i = (i + 1 == MT) ? 0 : i + 1;
					}
					tc_fence_before();
				}
				TICK(l1b); TACC(8, l1, l1b);
				// what the epilogue reads from global memory is requested only now: the proxy fences of the scatter above would otherwise
				// wait for these loads (MEMBAR), and the exchange that follows hides their latency just as well
				LaneCtl li[SPT];
				float sv[SPT], aux[SPT], old[SPT];
				uint32_t lab[SPT];
#pragma unroll
				for (int k = 0; k < SPT; k++) {
					li[k] = ctl->lc[pb][(sub & 3) * SPT + k];
					sv[k] = 0.0f; aux[k] = 0.0f; old[k] = 0.0f; lab[k] = LAB_BAD;
					if (!row_valid) continue;
					if (!BWD) {
						if (li[k].flags & 2u) {
							sv[k] = __ldg(p.S + li[k].nrow + c);
							if (my_d >= 2 && my_d <= li[k].nt) old[k] = p.G[li[k].nrow - my_d * Lp + c];
						}
					} else {
						if (li[k].flags & 1u) {
							sv[k] = __ldg(p.S + li[k].crow + c); aux[k] = __ldg(p.A + li[k].crow + c); lab[k] = __ldg(p.node_lab + li[k].cn);
						}
						if ((li[k].flags & 2u) && my_d >= 2 && my_d <= li[k].nt) old[k] = p.G[li[k].nrow + my_d * Lp + c];
					}
				}
				if (it > 0) {
					mbar_wait(&ctl->gather[pb], ((it - 1) >> 1) & 1);    // the partial products of my slice from every peer (each reader polls: no hand-over hop)
					if (in_tile) {
						const unsigned char* src = recv_pb + ((size_t)sub * CW + row) * 16;
						for (uint32_t r = 0; r < CS; r++) {
							const float4 v = *reinterpret_cast<const float4*>(src + (size_t)r * CH);
							g[0] += v.x; g[1] += v.y; g[2] += v.z; g[3] += v.w;
						}
					}
					TICK(l2); TACC(10, l1b, l2);
				}
				TICK(l2c);
				// phase A: everything the NEXT step waits for -- the new vector entries into the local MMA operand, their partial sums -- and
				// the barrier; phase B (below): the global stores.  The proxy fence in front of the barrier is a MEMBAR that waits for every
				// outstanding memory operation of the thread: with the stores issued first it held the whole step up until they had reached L2.
				// The arithmetic is written with selects instead of branches so that the compiler interleaves the four slots' dependent chains
				// (log -> add -> exp -> convert): with a branch per slot they ran one after the other.
				float val[SPT], st0[SPT], st1[SPT], st2[SPT];
				const uint32_t dd = row_valid ? my_d : 1u;                  // safe index into the per-duration tables
#pragma unroll
				for (int k = 0; k < SPT; k++) {
					const uint32_t s = (sub & 3) * SPT + k;
					const float dl = ctl->delta[pb][s][dd];
					st0[k] = 0.0f; st1[k] = 0.0f; st2[k] = 0.0f;
					if (!BWD) {
						const bool prod = row_valid && (li[k].flags & 2u) && c < li[k].navail;
						const bool use_log = prod && my_d <= li[k].nt;
						const float lg = __logf(use_log ? (my_d == 1 ? g[k] : old[k]) : 1.0f);
						const float e = __expf(sv[k] + dl + lg);
						val[k] = prod ? e : 0.0f;
					} else {
						const bool a0 = row_valid && (li[k].flags & 1u) && c < li[k].navail;
						const float uu = a0 ? ((li[k].flags & 4u) ? 1.0f : g[k]) : 0.0f;
						const float lu = __logf(a0 ? uu : 1.0f);
						const float lw = a0 ? sv[k] + lu : -INFINITY;
						const float gamma = aux[k] * __expf(lu + ctl->sg[pb][s]);
						const float er = __expf(lw + ctl->rsc[pb][s][dd]);
						st0[k] = a0 ? ((lab[k] == c) ? 1.0f : 0.0f) - gamma : 0.0f;     // Dm
						st1[k] = (a0 && my_d <= li[k].ct) ? er : 0.0f;                     // R
						st2[k] = uu;
						sv[k] = lw;                                                         // G: log-domain S+beta relative to base_t
						const bool a1 = row_valid && (li[k].flags & 2u) && my_d <= li[k].nt;
						const float ev = __expf((my_d == 1 ? lw : old[k]) + dl);
						val[k] = a1 ? ev : 0.0f;
					}
					if (in_tile) {
						const __nv_bfloat16 hi = __float2bfloat16_rn(val[k]);
						const __nv_bfloat16 lo = __float2bfloat16_rn(val[k] - __bfloat162float(hi));
						unsigned char* dst = vbuf + (row / 8) * 512 + (s / 8) * 128 + (s % 8) * 16 + (row % 8) * 2;
						*reinterpret_cast<__nv_bfloat16*>(dst) = hi; *reinterpret_cast<__nv_bfloat16*>(dst + 256) = lo;
					}
				}
				// partial sums over my 32 rows of my 4 slots: transposed butterfly (6 shuffles instead of 20)
				{
					const bool up16 = lane & 16, up8 = lane & 8;
					const float a0 = (up16 ? val[2] : val[0]) + __shfl_xor_sync(0xffffffffu, up16 ? val[0] : val[2], 16);
					const float a1 = (up16 ? val[3] : val[1]) + __shfl_xor_sync(0xffffffffu, up16 ? val[1] : val[3], 16);
					float b = (up8 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, up8 ? a0 : a1, 8);
					b += __shfl_xor_sync(0xffffffffu, b, 4);
					b += __shfl_xor_sync(0xffffffffu, b, 2);
					b += __shfl_xor_sync(0xffffffffu, b, 1);
					// lane (bit4, bit3) now holds the sum of slot k = 2*bit4 + bit3
					if (epi_warp && (lane & 7) == 0) *reinterpret_cast<float*>(myps + (rgrp * 16 + sub * SPT + 2 * (lane >> 4) + ((lane >> 3) & 1)) * 4) = b;
				}
				TICK(l3); TACC(9, l2c, l3);
				fence_proxy_async_smem();
				bar_arrive<BAR_LANES_2>(BAR_TILE);
				// phase B: the lattice arrays (read back by later steps of this thread, by the other direction and by the gradient GEMMs)
				if (row_valid) {
#pragma unroll
					for (int k = 0; k < SPT; k++) {
						if (!BWD) {
							if (li[k].flags & 4u) p.G[li[k].crow + c] = g[k];
							if (li[k].flags & 2u) p.A[li[k].nrow + c] = val[k];
						} else if (li[k].flags & 1u) {
							p.Dm[li[k].crow + c] = st0[k]; p.R[li[k].crow + c] = st1[k];
							if (p.Uvec) p.Uvec[li[k].crow + c] = st2[k];
							p.G[li[k].crow + c] = sv[k];      // read back by this thread d frames earlier
						}
					}
				}
			}
		}
		if (timing) {
			if (warp == MMA_WARP) { p.dbg[1] = tacc[1]; p.dbg[15] = tacc[11]; }
			else if (warp == COMM_WARP) { p.dbg[0] = tacc[0]; p.dbg[2] = tacc[2]; p.dbg[3] = tacc[3]; }
			else { for (int i = 7; i < 11; i++) p.dbg[i] = tacc[i]; }
		}
		// the partial sums pushed during the last executed step are never consumed: wait for them so that no bulk copy is in
		// flight (into this CTA or out of it) when the cluster retires
		if (warp == COMM_WARP && it > 0) mbar_wait(&ctl->psg[it & 1], ((it - 1) >> 1) & 1);
	}
	tc_fence_before();
	__syncthreads();
	if (CS > 1) cluster_sync_all();
	if (warp == MMA_WARP) tmem_dealloc(tmem, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
bool plan_ks_dp(uint32_t L, uint32_t D, int max_smem_optin, KsDpPlan* plan) {
	if (D >= DMAX) return false;   // the bookkeeping warp maps durations 0..D onto its 32 lanes
	for (uint32_t CS = 1; CS <= 8; CS *= 2) {
		const uint32_t CW = ((L + CS - 1) / CS + 15) / 16 * 16;
		if (CW > 128) continue;                                  // one lane thread per row of the slice
		if (CS > 1 && (CS - 1) * CW >= L) continue;              // an empty slice
		const uint32_t K = CS * CW, MT = (K + 127) / 128;
		if (MT > (uint32_t)MT_MAX) continue;
		const uint32_t need = MT * 16 + MT * CW;                 // accumulators + hi and lo halves of every tile
		if (need > 512) continue;
		uint32_t cols = 32; while (cols < need) cols *= 2;
		const size_t CH = (size_t)CW * 64;
		const size_t recv_off = 0, vbuf_off = recv_off + 2 * CS * CH;
		size_t ps_off = vbuf_off + CH;
		ps_off = (ps_off + 127) / 128 * 128;
		const size_t ctl_off = (ps_off + 2 * CS * PS_BYTES + 127) / 128 * 128;
		size_t total = ctl_off + ks_dp_ctl_bytes();
		if (total + 1024 > (size_t)max_smem_optin) continue;
		// resident CTAs own their TMEM columns until they exit: size the shared-memory request so that no more CTAs
		// fit on an SM than its 512 TMEM columns can serve (an over-subscribed tcgen05.alloc would stall a whole cluster)
		const size_t min_smem = (size_t)233472 / (512 / cols + 1) + 1;
		if (total < min_smem) total = min_smem;
		plan->CS = CS; plan->CW = CW; plan->K = K; plan->MT = MT; plan->tmem_cols = cols;
		plan->recv_off = (uint32_t)recv_off; plan->vbuf_off = (uint32_t)vbuf_off; plan->ps_off = (uint32_t)ps_off; plan->ctl_off = (uint32_t)ctl_off; plan->smem = total;
		return true;
	}
	return false;
}

static cudaError_t configure(void* kern, const KsDpPlan& plan, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, uint32_t n_clusters, cudaStream_t s) {
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
	if (e != cudaSuccess) return e;
	cfg->gridDim = dim3(n_clusters * plan.CS); cfg->blockDim = dim3(N_THREADS);
	cfg->dynamicSmemBytes = plan.smem; cfg->stream = s;
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = plan.CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg->attrs = attr; cfg->numAttrs = 1;
	return cudaSuccess;
}

int max_active_ks_clusters(const KsDpPlan& plan) {
	cudaLaunchConfig_t cfg{}; cudaLaunchAttribute attr[1];
	cudaError_t e = configure((void*)dp_ks_kernel<false>, plan, &cfg, attr, 1, nullptr);
	int n = 0;
	if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, dp_ks_kernel<false>, &cfg);
	if (getenv("CRFGPU_VERBOSE")) {
		cudaFuncAttributes fa{};
		cudaFuncGetAttributes(&fa, dp_ks_kernel<false>);
		fprintf(stderr, "[crfgpu] dp_ks_kernel: regs %d, max threads/block %d, static smem %zu, requested dynamic smem %zu, block %d threads, cluster %u -> %d active clusters (%s)\n",
		        fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes, plan.smem, N_THREADS, plan.CS, n, cudaGetErrorString(e));
	}
	if (e != cudaSuccess) {
		if (getenv("CRFGPU_VERBOSE")) fprintf(stderr, "[crfgpu] contraction-sliced lattice kernel not launchable (CS=%u, smem=%zu): %s\n", plan.CS, plan.smem, cudaGetErrorString(e));
		cudaGetLastError();
		return 0;
	}
	return n;
}

cudaError_t launch_ks_dp(bool backward, const KsDpParams& p, const KsDpPlan& plan, cudaStream_t s) {
	if (!p.n_clusters) return cudaSuccess;
	cudaLaunchConfig_t cfg{}; cudaLaunchAttribute attr[1];
	cudaError_t e = configure(backward ? (void*)dp_ks_kernel<true> : (void*)dp_ks_kernel<false>, plan, &cfg, attr, p.n_clusters, s);
	if (e != cudaSuccess) return e;
	return backward ? cudaLaunchKernelEx(&cfg, dp_ks_kernel<true>, p) : cudaLaunchKernelEx(&cfg, dp_ks_kernel<false>, p);
}

}  // namespace crfgpu
