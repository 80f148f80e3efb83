#include "crf_layout.h"

#include <algorithm>
#include <cmath>
#include <functional>
#include <numeric>
#include <stdexcept>

namespace crfgpu {

uint32_t stream_width(uint32_t F, uint32_t D, uint32_t seg, uint32_t lc, uint32_t rc, uint32_t bdelta) {
	// CRF_InFtrStream_SeqMultiWindow ctor (CRF/src/io/CRF_InFtrStream_SeqMultiWindow.cpp:47-117)
	if (D == 1) return (lc + 1 + rc) * F;
	if (seg) return 8 * F + D + (lc + rc) * F;              // left context | sample x5 | avg | max | min | one-hot duration | right context
	if (bdelta) return std::min(lc, rc + 1) * F;            // |left - right| of the frame pairs around the window's first frame
	return (lc + 1 + rc) * F;                               // left context | first frame | right context
}

uint32_t window_width(const crfgpu_config& c) {
	uint32_t w = stream_width(c.n_base_ftrs, c.max_dur, c.extract_seg_ftrs, c.left_ctx, c.right_ctx, c.boundary_delta);
	if (c.n_base_ftrs2) w += stream_width(c.n_base_ftrs2, c.max_dur, c.extract_seg_ftrs2, c.left_ctx2, c.right_ctx2, c.boundary_delta2);   // joined behind the first
	return w;
}

bool has_context_or_join(const crfgpu_config& c) { return c.left_ctx || c.right_ctx || c.boundary_delta || c.n_base_ftrs2; }

Layout build_layout(const crfgpu_config& c) {
	Layout m;
	if (c.n_states == 0 || c.n_labs == 0 || c.n_labs % c.n_states != 0)
		throw std::runtime_error("CRF_StdFeatureMap: invalid state/label combination while computing transitions");
	const uint32_t W = window_width(c);
	if (c.use_state_ftrs && (c.state_fidx_end < c.state_fidx_start || c.state_fidx_end >= W))
		throw std::runtime_error("state feature index range outside the window feature vector");
	if (c.use_trans_ftrs && (c.trans_fidx_end < c.trans_fidx_start || c.trans_fidx_end >= W))
		throw std::runtime_error("transition feature index range outside the window feature vector");
	m.L = c.n_labs;
	m.n_states = c.n_states;
	m.n_act = c.n_labs / c.n_states;
	m.nSf = c.use_state_ftrs ? c.state_fidx_end - c.state_fidx_start + 1 : 0;
	m.nTf = c.use_trans_ftrs ? c.trans_fidx_end - c.trans_fidx_start + 1 : 0;
	m.nS = m.nSf + (c.use_state_bias ? 1u : 0u);
	m.nT = m.nTf + (c.use_trans_bias ? 1u : 0u);
	// end->start block + self transitions + previous-sub-state transitions (:478-485)
	const uint64_t n_pairs = (m.n_states == 1) ? uint64_t(m.n_act) * m.n_act
	                                           : uint64_t(m.n_act) * m.n_act + m.L + (m.L - m.n_act);
	const uint64_t len = uint64_t(m.nS) * m.L + uint64_t(m.nT) * n_pairs;
	if (len >= 0xffffffffull) throw std::runtime_error("lambda vector does not fit 32-bit indexing");
	m.len = uint32_t(len);
	m.sidx.resize(m.L);
	m.tidx.assign(size_t(m.L) * m.L, CRFGPU_NO_IDX);
	const uint32_t L = m.L, N = m.n_states;
	if (N == 1) {
		// each label owns one contiguous block: [state funcs | for every previous label: trans funcs] (:295,365)
		const uint32_t block = m.nS + L * m.nT;
		for (uint32_t cl = 0; cl < L; cl++) {
			m.sidx[cl] = cl * block;
			for (uint32_t pl = 0; pl < L; pl++) m.tidx[size_t(pl) * L + cl] = cl * block + m.nS + pl * m.nT;
		}
	} else {
		// per label: state funcs, self transition, then (start sub-state) one block per phone for the
		// end->start arcs, or (inner sub-state) the single arc from the previous sub-state (:298-315, 369-405)
		uint32_t pos = 0;
		for (uint32_t cl = 0; cl < L; cl++) {
			m.sidx[cl] = pos;
			const uint32_t self = pos + m.nS;
			m.tidx[size_t(cl) * L + cl] = self;
			if (cl % N == 0) {
				for (uint32_t ph = 0; ph < m.n_act; ph++) {
					const uint32_t pl = ph * N + N - 1;
					if (pl != cl) m.tidx[size_t(pl) * L + cl] = self + m.nT + ph * m.nT;
				}
				pos = self + (m.n_act + 1) * m.nT;
			} else {
				m.tidx[size_t(cl - 1) * L + cl] = self + m.nT;
				pos = self + 2 * m.nT;
			}
		}
	}
	return m;
}

void group_labels(uint32_t D, uint32_t T, const uint32_t* labs, uint32_t* out4) {
	for (size_t i = 0; i < size_t(T) * 4; i++) out4[i] = CRFGPU_LAB_BAD;
	uint32_t run_start = 0;
	while (run_start < T) {
		uint32_t run_end = run_start;  // inclusive
		while (run_end + 1 < T && labs[run_end + 1] == labs[run_start]) run_end++;
		const uint32_t lab = labs[run_start], dur = run_end - run_start + 1;
		if (lab != CRFGPU_LAB_BAD) {
			if (dur <= D) {
				uint32_t* o = out4 + 4 * size_t(run_end);
				o[0] = lab; o[1] = run_start; o[2] = run_end; o[3] = 0;
			} else {
				// long runs are cut into ceil(dur/D) near-equal pieces, the longer ones first (:75-106)
				const uint32_t pieces = (dur + D - 1) / D, base = dur / pieces, extra = dur % pieces;
				uint32_t ps = run_start;
				for (uint32_t r = 0; r < pieces; r++) {
					const uint32_t pe = ps + base + (r < extra ? 1u : 0u) - 1;
					uint32_t* o = out4 + 4 * size_t(pe);
					o[0] = lab; o[1] = ps; o[2] = pe; o[3] = 1;
					ps = pe + 1;
				}
			}
		}
		run_start = run_end + 1;
	}
}

std::vector<uint32_t> sample_steps(uint32_t D) {
	std::vector<uint32_t> steps(size_t(D) * 5);
	for (uint32_t d = 1; d <= D; d++) {
		// float one_tenth_win_len = cur_win_len * 0.1;  frameStep = (QNUInt32)ceil(one_tenth_win_len * i) - 1  (:569-573)
		volatile float one_tenth = float(d * 0.1);
		int k = 0;
		for (int i = 1; i < 10; i += 2, k++) {
			volatile float prod = one_tenth * float(i);
			steps[size_t(d - 1) * 5 + k] = uint32_t(std::ceil(prod)) - 1;
		}
	}
	return steps;
}

void shard_views(uint32_t n_utt, uint32_t n_streams, uint32_t* first, uint32_t* count) {
	const uint32_t per = n_streams ? n_utt / n_streams : 0;
	for (uint32_t i = 0; i < n_streams; i++) {
		first[i] = i * per;
		count[i] = (i + 1 == n_streams) ? n_utt - i * per : per;     // view(i*nseg_per_child, last ? QN_ALL : nseg_per_child)
	}
}

uint32_t minibatch_share(uint32_t minibatch, uint32_t n_streams, uint32_t stream) {
	if (minibatch == 0 || minibatch == 0xffffffffu) return 0xffffffffu;    // setMinibatch(0): "totally batch"
	return minibatch / n_streams + (stream < minibatch % n_streams ? 1u : 0u);
}

void balance_utts(uint32_t n_utt, const uint32_t* n_frames, uint32_t n_ranks, uint32_t* rank_of) {
	if (!n_ranks) return;
	std::vector<uint32_t> order(n_utt);
	std::iota(order.begin(), order.end(), 0u);
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return n_frames[a] > n_frames[b]; });
	std::vector<uint64_t> load(n_ranks, 0);
	std::vector<uint32_t> cnt(n_ranks, 0);
	const uint32_t cap_lo = n_utt / n_ranks, n_hi = n_utt % n_ranks;   // ranks 0..n_hi-1 may take one more
	for (uint32_t u : order) {
		uint32_t best = n_ranks;
		for (uint32_t r = 0; r < n_ranks; r++) {
			if (cnt[r] >= cap_lo + (r < n_hi ? 1u : 0u)) continue;
			if (best == n_ranks || load[r] < load[best]) best = r;
		}
		rank_of[u] = best; load[best] += n_frames[u]; cnt[best]++;
	}
}

void balance_utts_cost(uint32_t n_utt, const uint32_t* n_frames, uint32_t n_ranks, uint32_t n_slots, double step_frames, uint32_t* rank_of) {
	if (!n_ranks) return;
	if (!n_slots) n_slots = 1;
	std::vector<uint32_t> order(n_utt);
	std::iota(order.begin(), order.end(), 0u);
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return n_frames[a] > n_frames[b]; });
	// Every rank deals ITS utterances longest first to the least loaded of its n_slots slots (crfgpu_stage_batch), and the global order
	// here is longest first too -- so the slot loads of a rank can be carried along exactly: a min-heap of slot loads per rank, the new
	// utterance lands on the rank's least loaded slot, lock-steps = the largest slot load.
	std::vector<std::vector<uint64_t>> heap(n_ranks, std::vector<uint64_t>(n_slots, 0));      // min-heaps (std::greater)
	std::vector<uint64_t> load(n_ranks, 0), steps(n_ranks, 0);
	for (uint32_t u : order) {
		uint32_t best = 0; double bc = 0.0;
		for (uint32_t r = 0; r < n_ranks; r++) {
			const uint64_t st = std::max<uint64_t>(steps[r], heap[r].front() + n_frames[u]);
			const double c = step_frames * (double)st + (double)(load[r] + n_frames[u]);
			if (r == 0 || c < bc) { best = r; bc = c; }
		}
		std::vector<uint64_t>& h = heap[best];
		std::pop_heap(h.begin(), h.end(), std::greater<uint64_t>());
		h.back() += n_frames[u];
		steps[best] = std::max(steps[best], h.back());
		std::push_heap(h.begin(), h.end(), std::greater<uint64_t>());
		rank_of[u] = best; load[best] += n_frames[u];
	}
}

}  // namespace crfgpu
