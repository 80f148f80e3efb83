// Host-side geometry of the CRF: lambda-vector layout, window widths, label grouping.
// Product code (part of libcrfgpu.so).  Citations are relative to the ASR-CRaFT tree.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/crfgpu.h"

namespace crfgpu {

// Layout of the lambda vector as defined by CRF_StdFeatureMap::recalc and its index functions
// (CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:280-320, 355-410, 472-517).
struct Layout {
	uint32_t L = 0;        // labels (stdseg: phones * max_dur)
	uint32_t n_states = 1; // sub-states per phone
	uint32_t n_act = 0;    // L / n_states ("numActualLabels" of the map)
	uint32_t nSf = 0, nTf = 0;  // state / transition FEATURES per label (pair)
	uint32_t nS = 0, nT = 0;    // state / transition FUNCTIONS (features + bias)
	uint32_t len = 0;      // lambda length
	std::vector<uint32_t> sidx;  // [L]
	std::vector<uint32_t> tidx;  // [L*L] at [plab*L+clab]; CRFGPU_NO_IDX when the topology forbids the pair
};

// Throws std::runtime_error on the geometry errors the reference throws for.
Layout build_layout(const crfgpu_config& c);

uint32_t window_width(const crfgpu_config& c);
// one stream's share of the window vector (CRF_InFtrStream_SeqMultiWindow ctor, CRF/src/io/CRF_InFtrStream_SeqMultiWindow.cpp:47-117)
uint32_t stream_width(uint32_t n_base_ftrs, uint32_t max_dur, uint32_t seg_ftrs, uint32_t left_ctx, uint32_t right_ctx, uint32_t boundary_delta);
// context frames on stream 1, boundary deltas or a joined second stream: the windows are built by expand_joined_kernel
bool has_context_or_join(const crfgpu_config& c);

// CRF_InLabStream_SeqMultiWindow (CRF/src/io/CRF_InLabStream_SeqMultiWindow.cpp:51-306): per frame
// (label,start,end,broken) on the frame where a (possibly split) reference segment ends, else LAB_BAD x4.
void group_labels(uint32_t max_dur, uint32_t n_frames, const uint32_t* frame_labs, uint32_t* out4);

// Sample offsets of CRF_InFtrStream_SeqMultiWindow::sample_ftrs (CRF/src/io/CRF_InFtrStream_SeqMultiWindow.cpp:556-590):
// steps[(d-1)*5+k] = frame offset from the window start of the k-th sampled frame for duration d.
// Computed on the host with the reference's own float arithmetic so the device never re-derives it.
std::vector<uint32_t> sample_steps(uint32_t max_dur);

// Sharding rules of the data-parallel training seam.
// shard_views: stream i of n_streams owns [i*floor(n/n_streams), (i+1)*floor(n/n_streams)), the last one also the remainder
// (CRF/src/io/CRF_FeatureStreamManager.cpp:425-464).
void shard_views(uint32_t n_utt, uint32_t n_streams, uint32_t* first, uint32_t* count);
// minibatch_share: utterances stream i takes per minibatch (CRF_Minibatch_GradAccumulator.cpp:229-257); 0 = whole view.
uint32_t minibatch_share(uint32_t minibatch, uint32_t n_streams, uint32_t stream);
// balance_utts: utterances of ONE global minibatch dealt to n_ranks devices with equal counts (+-1), longest first onto the rank
// with the fewest frames so far among those that still have room.
void balance_utts(uint32_t n_utt, const uint32_t* n_frames, uint32_t n_ranks, uint32_t* rank_of);
// balance_utts_cost: the same deal by a TIME model instead of equal counts.  The lattice kernels are a dependent chain of lock-steps whose
// number is the largest load of the rank's n_slots slots under its own longest-first dealing (carried along exactly), the rest of the step
// streams the rank's frames; with
// step_frames = cost of one lock-step in units of the per-frame cost, a rank costs step_frames * lock-steps + frames.  Utterances go,
// longest first, to the rank that is cheapest afterwards: a rank that holds one of the corpus' longest utterances ends up with fewer
// frames.  Membership of the global minibatch -- hence the gradient -- is unchanged; counts per rank differ.
void balance_utts_cost(uint32_t n_utt, const uint32_t* n_frames, uint32_t n_ranks, uint32_t n_slots, double step_frames, uint32_t* rank_of);

}  // namespace crfgpu
