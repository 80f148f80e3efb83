// libcrfgpu.so -- C ABI (include/crfgpu.h) over the sm_100a kernels.  Host code here only stages
// buffers, derives the device-side tables from lambda and sequences kernel launches on one stream.
// There is no CPU fallback: every compute entry point needs a CUDA device and fails loudly otherwise.
#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <stdexcept>
#include <string>
#include <functional>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include "../../include/crfgpu.h"
#include "crf_kernels.cuh"
#include "crf_layout.h"

using namespace crfgpu;

namespace {

thread_local std::string g_err;

struct ApiError : std::runtime_error {
	int code;
	ApiError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_OK(expr)                                                                                   \
	do {                                                                                                \
		cudaError_t e_ = (expr);                                                                        \
		if (e_ != cudaSuccess)                                                                          \
			throw ApiError(CRFGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));        \
	} while (0)

// growable device buffer
struct DevBuf {
	void* p = nullptr; size_t cap = 0;
	void ensure(size_t bytes) {
		if (bytes <= cap) return;
		if (p) CUDA_OK(cudaFree(p));
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 8 + 256;
		CUDA_OK(cudaMalloc(&p, want));
		cap = want;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
	template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct crfgpu_ctx {
	crfgpu_config cfg{};
	Layout lay;
	int device = 0;
	cudaStream_t stream = nullptr, copy_stream = nullptr;       // copy_stream: chunked H2D of the base features, overlapped with the window expansion
	cudaEvent_t ev_ready = nullptr; std::vector<cudaEvent_t> ev_chunk; std::vector<uint32_t> chunk_end;
	// page-locked arena for the per-batch index tables: their uploads must not serialise the host with the stream (a copy from pageable
	// memory first waits for everything queued on its stream)
	unsigned char* pin = nullptr; size_t pin_cap = 0, pin_used = 0; cudaEvent_t ev_pin = nullptr;
	// crfgpu_prefetch_batch: the NEXT minibatch's base features and windows, copied / expanded into a second buffer set on side streams
	// while the current minibatch computes; crfgpu_stage_batch swaps the sets when it is handed the batch that was prefetched
	DevBuf d_baseB2; const float* pre_ftrs2 = nullptr;             // (second stream of a joined model)
	DevBuf d_base2, d_X2, d_frame_t2, d_bpad2, d_Xa2; bool pre_virt = false; cudaStream_t pre_stream = nullptr; cudaEvent_t ev_pre_done = nullptr, ev_pre_ready = nullptr;
	cudaEvent_t ev_swap = nullptr; bool swap_marked = false;   // main-stream point after which the spare buffer set is free
	std::vector<cudaEvent_t> ev_chunk2; bool pre_valid = false; const float* pre_ftrs = nullptr; std::vector<uint32_t> pre_off;
	// ... and, when the read-ahead was handed the labels too (crfgpu_prefetch_train_batch), the per-frame index / label tables of that batch
	DevBuf d_frame_utt2, d_frame_len2, d_node_lab2, d_prev_lab2, d_next_lab2; const uint32_t* pre_labs = nullptr; bool pre_tabs = false;
	uint32_t W = 0;          // window feature width
	// virtual windows of the training GEMMs (see launch_virtual_windows): padded base stream [N][Fp] + aggregate blocks [N][D][Wa]
	bool virt = false, x_virt_valid = false, x_full_valid = false; int opt_virt = 1; uint32_t Fp = 0, Wa = 0; DevBuf d_bpad, d_Xa, d_bias_dy;
	uint32_t Wp = 0;         // stride between the windows of a frame in X (>= W; currently W, see crfgpu_create)
	uint32_t Lp = 0;         // padded label stride of the lattice arrays
	// label set the lattice kernels run on: the model's labels, or -- for the stdseg_no_dur* models -- the (duration, phone)
	// expansion with tied weights (same recursions as stdseg; the reference's own stdseg run with tied lambda gives the
	// identical logZ / numerators)
	uint32_t Lt = 0; bool tied = false;
	// native O(P^2 + D*P) recursion for the same models (crf_dp_nodur.cu): score / posterior columns are (duration, phone), the
	// transition tables and forward/backward vectors are P wide
	DevBuf d_WdT, d_vt_base, d_negMt; uint32_t vtE = 0;              // decoding with transition features: per-frame transition tables
	bool nodur_tf = false; DevBuf d_next_lab;                       // segmental no_dur model with transition features (nodur && nodur_tf)
	DevBuf d_XtK2, d_XtT2; bool x_tiles_valid = false, pre_tiles = false;   // the X tiles of a joined training batch are made at staging (read-ahead: into the spare set)
	DevBuf d_XtK, d_WtrT;                                           // ... and the transition scores: X slice and weights as K-major tiles (launch_tile_k)
	DevBuf d_XdT, d_XtT; int opt_tf_tiled = 1;                     // transition-feature gradient from pre-split, pre-tiled operands (launch_reduce_gemm_tiled)
	DevBuf d_Eall, d_rowmax;                                        // exp(M_n - max M_n) of every frame and the maxima (launch_transftr_exp)
	bool transftr = false; DevBuf d_Wtr, d_tbias, d_Mall, d_Xd;   // frame-level model with transition FEATURES (crf_dp_transftr.cu)
	bool nodur = false; uint32_t Pp = 0; int opt_nodur_impl = 0; int nodur_groups_max = 0; uint32_t n_nodur_groups = 0;
	DevBuf d_nd_grp, d_nd_batch, d_nd_xch, d_nd_ctr, d_LB;
	std::vector<uint32_t> t_sidx, t_tidx;   // lambda indices of the lattice labels / label pairs
	bool train_ok = false, decode_ok = false;
	std::string train_why, decode_why;
	uint64_t launches = 0;
	int opt_slots = 0, opt_keep_lattice = 0, opt_dp_impl = 3, opt_cluster_slots = 0, opt_max_clusters = 0, opt_gemm_impl = 2, opt_tma_mask = 63; uint32_t opt_k_slab = 1024, opt_k_slab_tc = 2048, opt_k_slab_tma = 4096, opt_k_slab_xi = 8192, opt_prefetch_smem = 1u << 20;
	int max_smem_optin = 0;
	bool cluster_ok = false; ClusterPlan plan{}; uint32_t n_clusters = 0;
	bool tc_ok = false; TcDpPlan tc_plan{}; uint32_t n_tc_clusters = 0;
	bool ks_ok = false; KsDpPlan ks_plan{}; uint32_t n_ks_clusters = 0;   // contraction-sliced tcgen05 lattice kernels (crf_dp_ks.cu): the default where the geometry fits
	uint64_t locksteps = 0;   // frames of the longest slot / cluster list of the staged batch = dependent steps of the lattice kernels
	DevBuf d_cl_off, d_cl_list, d_xch, d_xmax, d_smaxd;

	// model tables
	bool have_lambda = false;
	DevBuf d_lambda, d_sidx, d_tidx, d_Ws, d_Wt, d_bias, d_E, d_ET, d_steps;
	DevBuf d_Wd, d_crossT, d_negDiag, d_negOff, d_sidx0, d_tidx0, d_tmax;
	DevBuf d_lam_acc, d_lam_sqr_acc, d_grad_sqr_acc; bool have_train_state = false;   // crfgpu_sgd_update: accumulators of the averaged model / AdaGrad
	double Mmax = 0.0;

	// staged batch
	uint32_t n_utt = 0, N = 0; bool have_labels = false;
	std::vector<uint32_t> h_off;
	DevBuf d_off, d_base, d_frame_t, d_frame_utt, d_frame_len, d_node_lab, d_prev_lab, d_grp;
	uint32_t n_groups = 0; int U = 1;
	DevBuf d_X, d_S, d_A, d_G, d_m, d_kappa, d_bbase, d_Uvec, d_Dm, d_R, d_logZ, d_numer, d_grad, d_mass;
	int opt_mass_check = 1;
	bool fwdbwd_done = false;
	// viterbi
	DevBuf d_negS, d_candW, d_candP, d_bp, d_bd, d_gmove, d_olab, d_odur, d_ophn, d_nseg, d_cost;
	int opt_vit_eager = 1;   // decode batches: the recursion of each H2D chunk's utterances launched by crfgpu_stage_batch (0: by viterbi_staged, all at once)
	DevBuf d_order16, d_vg_xch, d_vg_final, d_vg_ctr, d_vg_cand; int opt_vit_impl = 0;   // group-sliced Viterbi (large phone sets)
	int opt_frame_impl = 0; bool frame_path = false;
	bool have_lm = false; DevBuf d_lm_start, d_lm_bigT, d_lm_final, d_lm_exit; double beam = 0.0;   // crfgpu_set_beam   // phone-bigram LM of the decoder (crfgpu_set_phone_lm)
	std::vector<cudaStream_t> rec_stream; std::vector<cudaEvent_t> ev_scored, ev_walked;   // one side stream per chunk: the chunks' recursions are latency chains and run beside each other
	bool vit_rec_ready = false; DevBuf d_vorder, d_off2;   // ... and the recursion of each chunk's utterances behind its scores (d_vorder: the chunks' utterances, longest first)
	cudaStream_t aux_stream = nullptr; cudaEvent_t ev_aux_go = nullptr, ev_aux_done = nullptr; int opt_aux_empirical = 1;   // empirical counts beside the recursions
	bool full_windows = false;      // crfgpu_expand_windows: gather every column of every window (no duration-1-only ranges)
	bool vit_score_ready = false;   // the decoder's fp64 scores of the staged batch were launched chunk by chunk behind the H2D copies
	bool viterbi_done = false;

	std::map<std::string, std::pair<cudaEvent_t, cudaEvent_t>> phases;
	ncclComm_t comm = nullptr; int comm_nranks = 0, comm_rank = 0;   // crfgpu_comm_init_*: the communicator crfgpu_allreduce_grad runs on

	// context frames / boundary deltas / joined second stream (crfgpu_*2): the windows are always materialised by expand_joined_kernel
	bool joined = false; DevBuf d_baseB;
	const float* X() const { return (cfg.max_dur == 1 && !joined) ? d_base.as<float>() : d_X.as<float>(); }
	uint64_t ldx() const { return (uint64_t)cfg.max_dur * Wp; }
};

namespace {

void phase_begin(crfgpu_ctx* h, const char* name) {
	auto& ev = h->phases[name];
	if (!ev.first) { CUDA_OK(cudaEventCreate(&ev.first)); CUDA_OK(cudaEventCreate(&ev.second)); }
	CUDA_OK(cudaEventRecord(ev.first, h->stream));
}
void phase_end(crfgpu_ctx* h, const char* name) { CUDA_OK(cudaEventRecord(h->phases[name].second, h->stream)); }

template <class T>
void upload(DevBuf& b, const std::vector<T>& v, cudaStream_t s) {
	b.ensure(v.size() * sizeof(T) + 16);
	if (!v.empty()) CUDA_OK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
}

// upload through the page-locked arena (truly asynchronous); falls back to the pageable path when the arena is full
template <class T>
void upload_async(crfgpu_ctx* h, DevBuf& b, const std::vector<T>& v) {
	const size_t bytes = v.size() * sizeof(T), at = (h->pin_used + 255) & ~(size_t)255;
	if (!bytes || at + bytes > h->pin_cap) { upload(b, v, h->stream); return; }
	b.ensure(bytes + 16);
	std::memcpy(h->pin + at, v.data(), bytes);
	h->pin_used = at + bytes;
	CUDA_OK(cudaMemcpyAsync(b.p, h->pin + at, bytes, cudaMemcpyHostToDevice, h->stream));
}

void check_kernel(crfgpu_ctx* h, int n_launches) {
	h->launches += n_launches;
	CUDA_OK(cudaGetLastError());
}

// Which paths the device implements for this geometry (anything else fails loudly, never emulated).
constexpr size_t TF_SMEM_MAX = 227 * 1024;      // dynamic shared memory one CTA can opt into on sm_100
void classify(crfgpu_ctx* h) {
	const crfgpu_config& c = h->cfg;
	h->train_ok = h->decode_ok = true; h->tied = h->nodur = false;
	h->transftr = h->nodur_tf = false;
	if (c.use_trans_ftrs) {
		// transition FEATURES (crf_featuremap=stdtrans): implemented for training frame-level models with one state per label
		// decoding: the decoder's transition tables become per-frame tables (fp64 scores in the reference's order, one row per frame)
		if (c.model_type != CRFGPU_STDFRAME && c.model_type != CRFGPU_STDSEG_NO_DUR_NO_SEGTRANSFTR) {
			h->decode_ok = false; h->decode_why = "CRFDecode accepts only stdframe / stdseg_no_dur_no_segtransftr (CRFDecode/src/Main.cpp:1065-1076)";
		} else if (c.n_labs > 1024 || c.max_dur > 255) { h->decode_ok = false; h->decode_why = "Viterbi kernel supports crf_label_size <= 1024 and max duration <= 255"; }
		// (N states per label: the illegal pairs of the N-state map score -inf, CRF_StdNStateNode.cpp:65-108)
		if ((c.model_type == CRFGPU_STDFRAME || c.model_type == CRFGPU_STDSEG_NO_DUR_NO_SEGTRANSFTR) && c.max_dur == 1 && c.n_labs <= 192 && transftr_smem_bytes(c.n_labs) <= TF_SMEM_MAX && c.use_state_ftrs) { h->transftr = true; return; }
		// ... and for the segmental production recipe: no duration labels, transition features from the duration-1 window
		// (N states per phone: CRF_StdSegNStateNode_WithoutDurLab_WithoutSegTransFtr.cpp:38-80, every sub-state a segment of its own)
		if (c.model_type == CRFGPU_STDSEG_NO_DUR_NO_SEGTRANSFTR && c.max_dur > 1 && c.max_dur <= 31 && c.n_labs <= 192 && nodur_tf_smem_bytes(c.n_labs) <= TF_SMEM_MAX && c.use_state_ftrs) {
			h->nodur = h->nodur_tf = true; return;
		}
		h->train_ok = false;
		h->train_why = "transition FEATURES (crf_featuremap=stdtrans) are implemented on the device for frame-level models up to 169 labels (phones x states) "
		               "and for stdseg_no_dur_no_segtransftr up to 161 labels, max_dur <= 31 (the double-buffered per-frame label x label matrix must fit "
		               "shared memory); other model types run with transition bias only";
		return;
	}
	if (c.model_type == CRFGPU_STDFRAME) {
		if (c.max_dur != 1) { h->train_ok = h->decode_ok = false; h->train_why = h->decode_why = "stdframe requires label_maximum_duration == 1"; }
	} else if (c.model_type == CRFGPU_STDSEG) {
		h->decode_ok = false; h->decode_why = "CRFDecode accepts only stdframe / stdseg_no_dur_no_segtransftr (CRFDecode/src/Main.cpp:1065-1076)";
		if (c.n_states != 1) { h->train_ok = false; h->train_why = "stdseg with crf_states > 1 throws in the reference (CRF_StateNode.cpp:497-502)"; }
		if (c.n_labs % c.max_dur != 0) { h->train_ok = false; h->train_why = "stdseg: crf_label_size must be phones * label_maximum_duration"; }
	} else if (c.model_type == CRFGPU_STDSEG_NO_DUR || c.model_type == CRFGPU_STDSEG_NO_DUR_NO_TRANSFTR || c.model_type == CRFGPU_STDSEG_NO_DUR_NO_SEGTRANSFTR) {
		// labels = phones, the duration lives in the window features only.  Without transition FEATURES the three no_dur node
		// families compute the same thing (checked against the reference); CRFDecode accepts only the last one.
		if (c.model_type != CRFGPU_STDSEG_NO_DUR_NO_SEGTRANSFTR) {
			h->decode_ok = false; h->decode_why = "CRFDecode accepts only stdframe / stdseg_no_dur_no_segtransftr (CRFDecode/src/Main.cpp:1065-1076)";
		}
		if (c.max_dur > 1) {
			// (N states per phone -- CRF_StdSegNStateNode_WithoutDurLab_WithoutSegTransFtr -- make every sub-state a segment of its own over
			// the N-state map's legal pairs: the same recursions with E = 0 on the illegal pairs)
			// two device paths: the tied (duration, phone) expansion on the dense stdseg kernels (phones * max_dur <= 1024) and the
			// native O(P^2 + D*P) recursion (any phone count whose matrix slices fit the shared memory of one group of CTAs)
			const bool tied_ok = (uint64_t)c.n_labs * c.max_dur <= 1024 && c.max_dur <= 32;
			h->nodur_groups_max = c.max_dur <= 31 ? nodur_max_groups(c.n_labs) : 0;
			const bool native_ok = h->nodur_groups_max >= 1;
			if (h->opt_nodur_impl == 1 ? native_ok : (h->opt_nodur_impl == 2 ? false : (!tied_ok && native_ok))) h->nodur = true;
			else if (tied_ok && h->opt_nodur_impl != 1) h->tied = true;
			else {
				h->train_ok = false;
				h->train_why = "forward-backward for stdseg_no_dur*: the tied (duration, phone) expansion needs phones * max_dur <= 1024, the native "
				               "recursion max_dur <= 31 and a phone count whose 32-column matrix slices fit shared memory (about 1050 phones)";
			}
		}
	} else {
		h->train_ok = h->decode_ok = false;
		h->train_why = h->decode_why = "unknown model type";
	}
	if (h->train_ok && !h->nodur && c.n_labs > 1024) { h->train_ok = false; h->train_why = "dense lattice kernels support crf_label_size <= 1024"; }
	if (h->decode_ok && (c.n_labs > 1024 || c.max_dur > 255)) { h->decode_ok = false; h->decode_why = "Viterbi kernel supports crf_label_size <= 1024 and max duration <= 255"; }
	if (h->train_ok && c.max_dur > 32) { h->train_ok = false; h->train_why = "lattice kernels support label_maximum_duration <= 32"; }
}

// lattice label space of the handle (depends on the no_dur implementation chosen by classify) and its lambda index tables
// virtual windows: segment windows whose whole width is the state-feature range, no transition features, TMA-fed tcgen05 GEMMs with the
// 128-row operand through tensor memory.  The weight tiles follow the virtual chunk order, so the tables are derived again when an option
// flips the decision (crfgpu_set_option).
bool decide_virt(crfgpu_ctx* h) {
	const crfgpu_config& c = h->cfg;
	const bool was = h->virt;
	h->Fp = (c.n_base_ftrs + 31) / 32 * 32; h->Wa = (3 * c.n_base_ftrs + 1 + 31) / 32 * 32;
	h->virt = h->opt_virt && !h->joined && h->train_ok && c.max_dur > 1 && c.extract_seg_ftrs && !c.use_trans_ftrs && c.use_state_ftrs && c.state_fidx_start == 0 &&
	          c.state_fidx_end + 1 == h->W && h->opt_gemm_impl == 2 && (h->opt_tma_mask & 27) == 27 && tma_gemm_eligible(nullptr, c.max_dur, h->Wa, 0);
	return was != h->virt;
}

void setup_label_space(crfgpu_ctx* h) {
	classify(h);
	const uint32_t L0 = h->lay.L, Dd = h->cfg.max_dur;
	h->Lt = (h->tied || h->nodur) ? L0 * Dd : L0;
	h->Lp = (h->Lt + 31) / 32 * 32;
	h->Pp = (L0 + 31) / 32 * 32;
	if (h->nodur) {
		h->t_sidx.resize(h->Lt);
		for (uint32_t q = 0; q < h->Lt; q++) h->t_sidx[q] = h->lay.sidx[q % L0];
		h->t_tidx = h->lay.tidx;                     // P x P
	} else if (!h->tied) { h->t_sidx = h->lay.sidx; h->t_tidx = h->lay.tidx; }
	else {
		h->t_sidx.resize(h->Lt); h->t_tidx.resize((size_t)h->Lt * h->Lt);
		for (uint32_t q = 0; q < h->Lt; q++) {
			h->t_sidx[q] = h->lay.sidx[q % L0];
			for (uint32_t cl = 0; cl < h->Lt; cl++) h->t_tidx[(size_t)q * h->Lt + cl] = h->lay.tidx[(size_t)(q % L0) * L0 + cl % L0];
		}
	}
	h->have_lambda = false; h->fwdbwd_done = false;
	decide_virt(h);
	if (h->stream) {
		upload(h->d_sidx, h->t_sidx, h->stream); upload(h->d_tidx, h->t_tidx, h->stream);
		if (h->decode_ok && (h->tied || h->nodur)) { upload(h->d_sidx0, h->lay.sidx, h->stream); upload(h->d_tidx0, h->lay.tidx, h->stream); }
		if (h->decode_ok && h->cfg.use_trans_ftrs) {
			// the decoder's transition pairs: end(pp) -> start(q) for all phone pairs, self loops, advance arcs (the same pairs the
			// constant tables crossT / negDiag / negOff hold for bias-only transitions)
			const Layout& m = h->lay; const uint32_t L = m.L, NS = m.n_states, P = m.n_act;
			std::vector<uint32_t> base((size_t)P * P + 2 * L, CRFGPU_NO_IDX);
			for (uint32_t pp = 0; pp < P; pp++) for (uint32_t q = 0; q < P; q++) base[(size_t)pp * P + q] = m.tidx[(size_t)(pp * NS + NS - 1) * L + q * NS];
			for (uint32_t l = 0; l < L; l++) {
				base[(size_t)P * P + l] = m.tidx[(size_t)l * L + l];
				if (l % NS != 0) base[(size_t)P * P + L + l] = m.tidx[(size_t)(l - 1) * L + l];
			}
			h->vtE = (uint32_t)base.size();
			upload(h->d_vt_base, base, h->stream);
		}
		CUDA_OK(cudaStreamSynchronize(h->stream));
	}
}

void require_train(crfgpu_ctx* h) {
	if (!h->train_ok) throw ApiError(CRFGPU_ERR_UNSUPPORTED, h->train_why);
	if (!h->have_lambda) throw ApiError(CRFGPU_ERR_ARG, "crfgpu_set_lambda has not been called");
}
void require_decode(crfgpu_ctx* h) {
	if (!h->decode_ok) throw ApiError(CRFGPU_ERR_UNSUPPORTED, h->decode_why);
	if (!h->have_lambda) throw ApiError(CRFGPU_ERR_ARG, "crfgpu_set_lambda has not been called");
}

// ---------------------------------------------------------------------------------------------
// Everything the kernels read from lambda, derived ON THE DEVICE from d_lambda (crf_lambda.cu): state weights / biases, E = exp(M - Mmax)
// with M[p][c] = lambda[tidx]*transBiasVal (CRF_StdFeatureMap.cpp:94-110 with no transition features), the weight tiles of the TMA-fed
// score GEMM and the decoder's tables.  Only the scalar Mmax comes back to the host.
void derive_tables(crfgpu_ctx* h) {
	h->vit_score_ready = h->vit_rec_ready = false;              // scores / paths launched ahead by crfgpu_stage_batch belong to the previous lambda
	const crfgpu_config& c = h->cfg; const Layout& m = h->lay;
	const uint32_t L = m.L, Lt = h->Lt, Lp = h->Lp, nSf = m.nSf;
	cudaStream_t s = h->stream;
	LambdaTablesParams p{};
	p.lam = h->d_lambda.as<double>(); p.nSf = nSf;
	p.use_state_bias = c.use_state_bias; p.use_trans_bias = c.use_trans_bias; p.state_bias_val = c.state_bias_val; p.trans_bias_val = c.trans_bias_val;
	h->d_tmax.ensure(sizeof(double) + 16); p.tmax = h->d_tmax.as<double>();
	if (h->train_ok) {
		// tables over the lattice labels (for tied models label (d,y) reads the weights of phone y; the native no_dur path keeps
		// P-wide tables and shares them between the duration blocks)
		const uint32_t Le = h->nodur ? L : Lt, Lpe = h->nodur ? h->Pp : Lp;
		p.Le = Le; p.Lpe = Lpe; p.sidx = h->d_sidx.as<uint32_t>(); p.tidx = h->d_tidx.as<uint32_t>();
		h->d_Ws.ensure(sizeof(float) * (size_t)Le * std::max(nSf, 1u) + 16); h->d_bias.ensure(sizeof(float) * Le + 16);
		h->d_E.ensure(sizeof(float) * (size_t)Le * Lpe + 16); h->d_ET.ensure(sizeof(float) * (size_t)Le * Lpe + 16);
		CUDA_OK(cudaMemsetAsync(h->d_E.p, 0, sizeof(float) * (size_t)Le * Lpe, s)); CUDA_OK(cudaMemsetAsync(h->d_ET.p, 0, sizeof(float) * (size_t)Le * Lpe, s));
		p.Ws = h->d_Ws.as<float>(); p.bias = h->d_bias.as<float>(); p.E = h->d_E.as<float>(); p.ET = h->d_ET.as<float>();
		if (c.max_dur > 1 && nSf > 0) {   // bf16 hi/lo UMMA tiles of the state weights for the TMA-fed score GEMM
			p.wt_P = h->nodur ? L : Lt / c.max_dur; p.wt_D = h->nodur ? 1 : c.max_dur; p.wt_chunks = score_tma_chunks(nSf);
			if (h->virt) {
				p.wt_virt = 1; p.wt_cpb = h->Fp / 32; p.wt_F = c.n_base_ftrs; p.wt_Dd = c.max_dur; p.wt_chunks = 5 * p.wt_cpb + h->Wa / 32;
				h->d_bias_dy.ensure(sizeof(float) * (size_t)c.max_dur * p.wt_P + 16); p.bias_dy = h->d_bias_dy.as<float>();
			}
			h->d_Wt.ensure((size_t)p.wt_D * ((p.wt_P + 63) / 64) * p.wt_chunks * 8192 + 16);
			p.Wt = h->d_Wt.as<unsigned char>();
		}
	}
	if (h->transftr || h->nodur_tf) {
		const uint32_t nTf = m.nTf;
		h->d_Wtr.ensure(sizeof(float) * (size_t)L * L * nTf + 16); h->d_tbias.ensure(sizeof(float) * (size_t)L * L + 16);
		p.Wtr = h->d_Wtr.as<float>(); p.tbias = h->d_tbias.as<float>(); p.nTf = nTf;
		p.sidx0 = h->d_sidx.as<uint32_t>(); p.tidx0 = h->d_tidx.as<uint32_t>(); p.L0 = L;
	}
	if (h->decode_ok && c.use_trans_ftrs) {
		h->d_WdT.ensure(sizeof(double) * (size_t)(m.nTf + 1) * h->vtE + 16);
		p.WdT = h->d_WdT.as<double>(); p.vt_base = h->d_vt_base.as<uint32_t>(); p.vtE = h->vtE; p.nTf = m.nTf;
	}
	if (h->decode_ok) {
		const bool same = !h->tied && !h->nodur;
		h->d_Wd.ensure(sizeof(double) * (size_t)(nSf + 1) * L + 16); h->d_crossT.ensure(sizeof(float) * (size_t)m.n_act * m.n_act + 16);
		h->d_negDiag.ensure(sizeof(float) * L + 16); h->d_negOff.ensure(sizeof(float) * L + 16);
		p.Wd = h->d_Wd.as<double>(); p.crossT = h->d_crossT.as<float>(); p.negDiag = h->d_negDiag.as<float>(); p.negOff = h->d_negOff.as<float>();
		p.sidx0 = same ? h->d_sidx.as<uint32_t>() : h->d_sidx0.as<uint32_t>(); p.tidx0 = same ? h->d_tidx.as<uint32_t>() : h->d_tidx0.as<uint32_t>();
		p.L0 = L; p.NS = m.n_states; p.P0 = m.n_act;
	}
	CUDA_OK(launch_lambda_tables(p, s));
	if ((h->transftr || h->nodur_tf) && h->opt_tf_tiled) {
		// the transition weights once more as bf16 hi / lo tiles in the byte order of the score GEMM's shared-memory tiles
		h->d_WtrT.ensure(tiled_k_operand_bytes(L * L, m.nTf, 128) + 16);
		CUDA_OK(launch_tile_k(h->d_Wtr.as<float>(), m.nTf, L * L, m.nTf, true, h->d_WtrT.as<unsigned char>(), s));
		h->launches++;
	}
	h->launches += h->train_ok ? 2 : 1;
	h->Mmax = 0.0;
	if (h->train_ok) {
		double t = 0.0;
		CUDA_OK(cudaMemcpyAsync(&t, h->d_tmax.p, sizeof(double), cudaMemcpyDeviceToHost, s));
		CUDA_OK(cudaStreamSynchronize(s));
		h->Mmax = (t == -DBL_MAX) ? 0.0 : t;
	}
	h->have_lambda = true;
}

void set_lambda(crfgpu_ctx* h, const double* lam, uint32_t len) {
	if (len != h->lay.len) throw ApiError(CRFGPU_ERR_ARG, "lambda length " + std::to_string(len) + " != feature map length " + std::to_string(h->lay.len));
	h->d_lambda.ensure(sizeof(double) * (size_t)len + 16);
	CUDA_OK(cudaMemcpyAsync(h->d_lambda.p, lam, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, h->stream));
	derive_tables(h);
	CUDA_OK(cudaStreamSynchronize(h->stream));   // the caller may reuse `lam` at once
}

// ---------------------------------------------------------------------------------------------// Base features to the device in up to 4 chunks cut at utterance boundaries on the copy stream, and -- on stream xs -- the window
// expansion of chunk i (windows never reach across utterances) while chunk i+1 is still in flight.  d_ft must already be queued on xs.
// virt: the virtual-window form (padded base stream d_bp + aggregate blocks d_Xa) instead of the full windows d_X.
void copy_and_expand(crfgpu_ctx* h, uint32_t n_utt, const uint32_t* off, uint32_t N, const float* ftrs, DevBuf& d_base, DevBuf& d_X, DevBuf& d_ft,
                     cudaStream_t xs, cudaEvent_t ev_ready, std::vector<cudaEvent_t>& evs, const char* phase, uint32_t dpart = 0,
                     const std::function<void(uint32_t, uint32_t)>* after_chunk = nullptr, bool virt = false, DevBuf* d_bp = nullptr, DevBuf* d_Xa = nullptr,
                     uint32_t want_chunks = 4) {
	const crfgpu_config& c = h->cfg;
	if (!N) return;
	if (!h->copy_stream) CUDA_OK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
	CUDA_OK(cudaEventRecord(ev_ready, xs));                              // everything queued on xs so far may still read the old contents
	CUDA_OK(cudaStreamWaitEvent(h->copy_stream, ev_ready, 0));
	const uint32_t n_chunks = N >= (1u << 14) ? want_chunks : 1u;
	std::vector<uint32_t> ends;
	uint32_t u = 0, n_prev = 0;
	for (uint32_t k = 1; k <= n_chunks; k++) {
		const uint64_t target = (uint64_t)N * k / n_chunks;
		while (u < n_utt && off[u + 1] <= target) u++;
		const uint32_t n_end = (k == n_chunks) ? N : off[u];
		if (n_end <= n_prev) continue;
		if (evs.size() <= ends.size()) { cudaEvent_t e; CUDA_OK(cudaEventCreate(&e)); evs.push_back(e); }
		CUDA_OK(cudaMemcpyAsync(d_base.as<float>() + (size_t)n_prev * c.n_base_ftrs, ftrs + (size_t)n_prev * c.n_base_ftrs,
		                        sizeof(float) * (size_t)(n_end - n_prev) * c.n_base_ftrs, cudaMemcpyHostToDevice, h->copy_stream));
		CUDA_OK(cudaEventRecord(evs[ends.size()], h->copy_stream));
		ends.push_back(n_end);
		n_prev = n_end;
	}
	if (phase) phase_begin(h, phase);
	n_prev = 0;
	for (size_t k = 0; k < ends.size(); k++) {
		CUDA_OK(cudaStreamWaitEvent(xs, evs[k], 0));
		if (c.max_dur > 1 && virt) {
			launch_virtual_windows(d_base.as<float>(), d_ft.as<uint32_t>(), d_bp->as<float>(), d_Xa->as<float>(), N, c.n_base_ftrs, h->Fp, c.max_dur, h->Wa, n_prev, ends[k], xs);
			check_kernel(h, 2);
		} else if (c.max_dur > 1) {
			ExpandParams ep{d_base.as<float>(), d_ft.as<uint32_t>(), h->d_steps.as<uint32_t>(), d_X.as<float>(),
			                N, c.n_base_ftrs, c.max_dur, h->W, h->Wp, c.extract_seg_ftrs, n_prev, dpart};
			launch_expand_windows(ep, ends[k], xs);
			check_kernel(h, 1);
		}
		if (after_chunk) (*after_chunk)(n_prev, ends[k]);      // work on the frames of this chunk while the next one is in flight
		n_prev = ends[k];
	}
	if (phase) phase_end(h, phase);
	if (phase) h->chunk_end = ends;      // the chunk list of the timed (non-read-ahead) staging: indexes ev_chunk in the verbose timeline
}

void validate_offsets(uint32_t n_utt, const uint32_t* off, const float* ftrs) {
	if (!off || !ftrs) throw ApiError(CRFGPU_ERR_ARG, "null input pointer");
	if (off[0] != 0) throw ApiError(CRFGPU_ERR_ARG, "frame_off[0] must be 0");
	for (uint32_t u = 0; u < n_utt; u++)
		if (off[u + 1] <= off[u])
			throw ApiError(CRFGPU_ERR_ARG, "utterance " + std::to_string(u) + " has no frames (reference: \"No features read from this sentence\")");
}

// node label = (dur-1)*nActualLabs + phone on the frame where a reference segment ends (CRF_NewGradBuilder_StdSeg.cpp:172-187);
// prev_lab = label of the preceding reference segment (:343-351).  Frame-level models label every frame (window length 1 cuts every
// run into 1-frame pieces).  next_lab (transition-feature no_dur models only): phone of the NEXT reference segment at every frame where
// a reference segment ends (the builder's next_lab, CRF_NewGradBuilder_StdSeg_NoDur_NoTrans.cpp:436-446)
void build_label_tables(crfgpu_ctx* h, uint32_t n_utt, const uint32_t* off, const uint32_t* labs, std::vector<uint32_t>& node_lab,
                        std::vector<uint32_t>& prev_lab, std::vector<uint32_t>& next_lab) {
	const crfgpu_config& c = h->cfg;
	const uint32_t N = n_utt ? off[n_utt] : 0;
	node_lab.assign(N, CRFGPU_LAB_BAD); prev_lab.assign(N, CRFGPU_LAB_BAD);
	std::vector<uint32_t> rec;
	for (uint32_t u = 0; u < n_utt; u++) {
		const uint32_t T = off[u + 1] - off[u];
		rec.resize((size_t)T * 4);
		group_labels(c.max_dur, T, labs + off[u], rec.data());
		uint32_t last = CRFGPU_LAB_BAD;
		for (uint32_t t = 0; t < T; t++) {
			prev_lab[off[u] + t] = last;
			if (rec[4 * (size_t)t] != CRFGPU_LAB_BAD) {
				const uint32_t dur = rec[4 * (size_t)t + 2] - rec[4 * (size_t)t + 1] + 1;
				const uint32_t lab = (c.model_type == CRFGPU_STDSEG || h->tied || h->nodur) ? c.n_actual_labs * (dur - 1) + rec[4 * (size_t)t] : rec[4 * (size_t)t];
				node_lab[off[u] + t] = lab; last = lab;
			}
		}
	}
	next_lab.clear();
	if (h->nodur_tf) {
		next_lab.assign(N, CRFGPU_LAB_BAD);
		for (uint32_t u = 0; u < n_utt; u++) {
			uint32_t nxt = CRFGPU_LAB_BAD;
			for (uint32_t n = off[u + 1]; n-- > off[u];)
				if (node_lab[n] != CRFGPU_LAB_BAD) { next_lab[n] = nxt; nxt = node_lab[n] % c.n_actual_labs; }
		}
	}
}

// joined / context window streams of a batch: both streams to the device whole, one gather kernel builds the windows (on stream st,
// into the given buffer set: the staged one or the read-ahead's spare one)
// the transition slice of the duration-1 windows as bf16 hi / lo tiles for the two labels^2 GEMMs (K-major for the scores, frame-major
// with the constant-1 bias column for the gradient): they depend on the windows only, so a training batch gets them at staging
void tile_trans_slice(crfgpu_ctx* h, uint32_t N, const float* X, DevBuf& XtK, DevBuf& XtT, cudaStream_t st) {
	const crfgpu_config& c = h->cfg; const Layout& m = h->lay;
	const uint32_t J = m.nTf + (c.use_trans_bias ? 1 : 0), ones = c.use_trans_bias ? m.nTf : 0xffffffffu;
	XtK.ensure(tiled_k_operand_bytes(N, m.nTf, 128) + 16); XtT.ensure(tiled_operand_bytes(N, J, 128) + 16);
	CUDA_OK(launch_tile_k(X + c.trans_fidx_start, (uint64_t)c.max_dur * h->Wp, N, m.nTf, true, XtK.as<unsigned char>(), st)); check_kernel(h, 1);
	CUDA_OK(launch_tile_mn(X + c.trans_fidx_start, (uint64_t)c.max_dur * h->Wp, J, ones, N, true, XtT.as<unsigned char>(), st)); check_kernel(h, 1);
}

void stage_joined(crfgpu_ctx* h, uint32_t n_utt, uint32_t N, const float* ftrs, const float* ftrs2, DevBuf& base, DevBuf& baseB, DevBuf& X,
                  const uint32_t* d_ft, const uint32_t* d_fu, cudaStream_t st, DevBuf* XtK = nullptr, DevBuf* XtT = nullptr) {
	const crfgpu_config& c = h->cfg;
	if (!N) return;
	const size_t n1 = (size_t)N + (size_t)n_utt * (c.left_ctx + c.right_ctx), n2 = (size_t)N + (size_t)n_utt * (c.left_ctx2 + c.right_ctx2);
	base.ensure(sizeof(float) * n1 * c.n_base_ftrs + 16);
	CUDA_OK(cudaMemcpyAsync(base.p, ftrs, sizeof(float) * n1 * c.n_base_ftrs, cudaMemcpyHostToDevice, st));
	if (c.n_base_ftrs2) {
		baseB.ensure(sizeof(float) * n2 * c.n_base_ftrs2 + 16);
		CUDA_OK(cudaMemcpyAsync(baseB.p, ftrs2, sizeof(float) * n2 * c.n_base_ftrs2, cudaMemcpyHostToDevice, st));
	}
	X.ensure(sizeof(float) * (size_t)N * c.max_dur * h->Wp + 16);
	ExpandJoinedParams ep{};
	ep.part[0] = JoinedPart{base.as<float>(), c.n_base_ftrs, c.extract_seg_ftrs, c.left_ctx, c.right_ctx, c.boundary_delta,
	                        stream_width(c.n_base_ftrs, c.max_dur, c.extract_seg_ftrs, c.left_ctx, c.right_ctx, c.boundary_delta)};
	ep.n_parts = 1;
	if (c.n_base_ftrs2) {
		ep.part[1] = JoinedPart{baseB.as<float>(), c.n_base_ftrs2, c.extract_seg_ftrs2, c.left_ctx2, c.right_ctx2, c.boundary_delta2,
		                        stream_width(c.n_base_ftrs2, c.max_dur, c.extract_seg_ftrs2, c.left_ctx2, c.right_ctx2, c.boundary_delta2)};
		ep.n_parts = 2;
	}
	ep.frame_t = d_ft; ep.frame_utt = d_fu; ep.steps = h->d_steps.as<uint32_t>();
	ep.X = X.as<float>(); ep.N = N; ep.D = c.max_dur; ep.Wp = h->Wp;
	// stdseg_no_dur_no_segtransftr + stdtrans: the transition scores (training and decoding) read the duration-1 window only, so in
	// the longer windows only the state-feature range is gathered -- 62 % of the recipe's 121 kB per frame are transition columns
	if (c.model_type == CRFGPU_STDSEG_NO_DUR_NO_SEGTRANSFTR && c.use_trans_ftrs && c.use_state_ftrs && c.max_dur > 1 && !h->full_windows) {
		ep.keep_lo = c.state_fidx_start; ep.keep_hi = c.state_fidx_end + 1;
	}
	launch_expand_joined(ep, st); check_kernel(h, 1);
	if (XtK && XtT) tile_trans_slice(h, N, X.as<float>(), *XtK, *XtT, st);
}

void prefetch_batch(crfgpu_ctx* h, uint32_t n_utt, const uint32_t* off, const float* ftrs, const uint32_t* labs = nullptr, const float* ftrs2 = nullptr) {
	const crfgpu_config& c = h->cfg;
	validate_offsets(n_utt, off, ftrs);
	const uint32_t N = n_utt ? off[n_utt] : 0;
	h->pre_valid = false;
	if (!N) return;
	if (h->joined && c.n_base_ftrs2 && !ftrs2) return;                    // a joined model without its second stream: nothing to read ahead
	if (!h->pre_stream) {
		CUDA_OK(cudaStreamCreateWithFlags(&h->pre_stream, cudaStreamNonBlocking));
		CUDA_OK(cudaEventCreate(&h->ev_pre_done)); CUDA_OK(cudaEventCreate(&h->ev_pre_ready));
	}
	if (h->swap_marked) CUDA_OK(cudaStreamWaitEvent(h->pre_stream, h->ev_swap, 0));
	if (h->joined) {
		// joined / context window streams (crfgpu_prefetch_train_batch2): the whole staging of crfgpu_stage_batch2 on the side stream,
		// into the spare buffer set
		h->pre_off.assign(off, off + n_utt + 1);
		upload(h->d_off2, h->pre_off, h->pre_stream);
		h->d_frame_t2.ensure(sizeof(uint32_t) * (size_t)N + 16); h->d_frame_utt2.ensure(sizeof(uint32_t) * (size_t)N + 16); h->d_frame_len2.ensure(sizeof(uint32_t) * (size_t)N + 16);
		launch_frame_tables(h->d_off2.as<uint32_t>(), n_utt, N, h->d_frame_t2.as<uint32_t>(), h->d_frame_utt2.as<uint32_t>(), h->d_frame_len2.as<uint32_t>(), h->pre_stream);
		check_kernel(h, 1);
		h->pre_tiles = labs != nullptr && (h->transftr || h->nodur_tf) && h->opt_tf_tiled;
		stage_joined(h, n_utt, N, ftrs, ftrs2, h->d_base2, h->d_baseB2, h->d_X2, h->d_frame_t2.as<uint32_t>(), h->d_frame_utt2.as<uint32_t>(), h->pre_stream,
		             h->pre_tiles ? &h->d_XtK2 : nullptr, h->pre_tiles ? &h->d_XtT2 : nullptr);
		h->pre_tabs = false; h->pre_labs = labs;
		if (labs) {
			std::vector<uint32_t> node_lab, prev_lab, next_lab;
			build_label_tables(h, n_utt, off, labs, node_lab, prev_lab, next_lab);
			upload(h->d_node_lab2, node_lab, h->pre_stream); upload(h->d_prev_lab2, prev_lab, h->pre_stream);
			if (!next_lab.empty()) upload(h->d_next_lab2, next_lab, h->pre_stream);
			h->pre_tabs = true;
		}
		CUDA_OK(cudaEventRecord(h->ev_pre_done, h->pre_stream));
		h->pre_virt = false; h->pre_ftrs = ftrs; h->pre_ftrs2 = ftrs2; h->pre_valid = true;
		return;
	}
	h->d_base2.ensure(sizeof(float) * (size_t)N * c.n_base_ftrs + 16);
	h->pre_virt = h->virt;                                               // the read-ahead of a training loop: the form the training GEMMs read
	if (h->pre_virt) { h->d_bpad2.ensure(sizeof(float) * (size_t)N * h->Fp + 16); h->d_Xa2.ensure(sizeof(float) * (size_t)N * c.max_dur * h->Wa + 16); }
	else if (c.max_dur > 1) h->d_X2.ensure(sizeof(float) * (size_t)N * c.max_dur * h->Wp + 16);
	// per-frame index tables: built on the device from the utterance offsets (launch_frame_tables)
	h->pre_off.assign(off, off + n_utt + 1);
	upload(h->d_off2, h->pre_off, h->pre_stream);                        // pageable: waits only for this side stream's own earlier work
	h->d_frame_t2.ensure(sizeof(uint32_t) * (size_t)N + 16);
	if (labs) { h->d_frame_utt2.ensure(sizeof(uint32_t) * (size_t)N + 16); h->d_frame_len2.ensure(sizeof(uint32_t) * (size_t)N + 16); }
	launch_frame_tables(h->d_off2.as<uint32_t>(), n_utt, N, h->d_frame_t2.as<uint32_t>(), labs ? h->d_frame_utt2.as<uint32_t>() : nullptr,
	                    labs ? h->d_frame_len2.as<uint32_t>() : nullptr, h->pre_stream);
	check_kernel(h, 1);
	h->pre_tabs = false; h->pre_labs = labs;
	if (labs) {
		// the label-derived tables of the next minibatch, built while this one computes (about 1 ms of host work per cfg4 minibatch
		// that crfgpu_stage_batch would otherwise do with the device idle)
		std::vector<uint32_t> node_lab, prev_lab, next_lab;
		build_label_tables(h, n_utt, off, labs, node_lab, prev_lab, next_lab);
		upload(h->d_node_lab2, node_lab, h->pre_stream); upload(h->d_prev_lab2, prev_lab, h->pre_stream);
		if (!next_lab.empty()) upload(h->d_next_lab2, next_lab, h->pre_stream);
		h->pre_tabs = true;
	}
	// optional cap on the expansion's shared memory per CTA (option prefetch_smem): small duration groups fit beside a resident lattice
	// CTA, but measured on cfg4 that SLOWS the step (e2e 13.0 M frames/s at 12 KB vs 14.1 M uncapped: the co-resident CTAs steal issue
	// and shared-memory bandwidth from the latency-bound recursion), so by default the read-ahead uses whole-frame CTAs
	uint32_t dpart = c.max_dur;
	while (dpart > 1 && sizeof(float) * ((size_t)dpart * h->Wp + (size_t)c.max_dur * (c.n_base_ftrs + 5)) > (size_t)h->opt_prefetch_smem) dpart--;
	copy_and_expand(h, n_utt, off, N, ftrs, h->d_base2, h->d_X2, h->d_frame_t2, h->pre_stream, h->ev_pre_ready, h->ev_chunk2, nullptr, dpart, nullptr,
	                h->pre_virt, &h->d_bpad2, &h->d_Xa2);
	CUDA_OK(cudaEventRecord(h->ev_pre_done, h->pre_stream));
	h->pre_ftrs = ftrs; h->pre_valid = true;
}

// ---- decoder plumbing shared by crfgpu_stage_batch (recursion launched chunk by chunk behind the H2D copies) and viterbi_staged ----
void ensure_decode_buffers(crfgpu_ctx* h) {
	const uint32_t N = h->N, L = h->lay.L, D = h->cfg.max_dur;
	h->d_negS.ensure(sizeof(float) * (size_t)N * D * L + 16);
	h->d_candW.ensure(sizeof(float) * (size_t)h->n_utt * D * L + 16); h->d_candP.ensure(sizeof(int32_t) * (size_t)h->n_utt * D * L + 16);
	h->d_bp.ensure(sizeof(uint16_t) * (size_t)N * L + 16); h->d_bd.ensure((size_t)N * L + 16); h->d_gmove.ensure((size_t)N + 16);
	h->d_olab.ensure(sizeof(uint32_t) * (size_t)N + 16); h->d_odur.ensure(sizeof(uint32_t) * (size_t)N + 16); h->d_ophn.ensure(sizeof(uint32_t) * (size_t)N + 16);
	h->d_nseg.ensure(sizeof(uint32_t) * (size_t)h->n_utt + 16); h->d_cost.ensure(sizeof(float) * (size_t)h->n_utt + 16);
}
// large phone sets with one state per phone and constant transition tables: the cross-phone table sliced over groups of CTAs,
// 16 utterances in lock-step per group (crf_viterbi_group.cu); opt_vit_impl 1 forces one CTA per utterance, 2 forces the groups
bool vit_wants_groups(const crfgpu_ctx* h) {
	const crfgpu_config& c = h->cfg; const Layout& m = h->lay;
	const uint32_t L = m.L, NS = m.n_states, P = m.n_act;
	const bool vg_fit = NS == 1 && !c.use_trans_ftrs && P >= 2 && L == P && !h->have_lm && !(h->beam > 0.0);      // (LM weights / beam: the per-utterance kernel)
	const bool vg_auto = vg_fit && (size_t)P * P * sizeof(float) > 96 * 1024;
	return vg_fit && (h->opt_vit_impl == 2 || (h->opt_vit_impl == 0 && vg_auto));
}
// the per-utterance kernel's parameters; order / n_utt are left to the caller (all utterances, or the utterances of one H2D chunk)
VitParams vit_params(crfgpu_ctx* h) {
	const crfgpu_config& c = h->cfg; const Layout& m = h->lay;
	VitParams v{};
	v.L = m.L; v.P = m.n_act; v.NS = m.n_states; v.D = c.max_dur; v.off = h->d_off.as<uint32_t>(); v.negS = h->d_negS.as<float>();
	v.crossT = h->d_crossT.as<float>(); v.negDiag = h->d_negDiag.as<float>(); v.negOff = h->d_negOff.as<float>();
	v.negMt = c.use_trans_ftrs ? h->d_negMt.as<float>() : nullptr; v.E = h->vtE;
	v.beam = h->beam;
	if (h->have_lm) {
		v.lm_start = h->d_lm_start.as<float>(); v.lm_final = h->d_lm_final.as<float>();
		if (m.n_states == 1) v.lm_bigT = h->d_lm_bigT.as<float>(); else v.lm_exit = h->d_lm_exit.as<float>();
	}
	v.candW = h->d_candW.as<float>(); v.candP = h->d_candP.as<int32_t>(); v.keptW = nullptr;
	v.bp = h->d_bp.as<uint16_t>(); v.bd = h->d_bd.as<uint8_t>(); v.gmove = h->d_gmove.as<uint8_t>();
	v.out_lab = h->d_olab.as<uint32_t>(); v.out_dur = h->d_odur.as<uint32_t>(); v.out_phn = h->d_ophn.as<uint32_t>();
	v.n_seg = h->d_nseg.as<uint32_t>(); v.cost = h->d_cost.as<float>();
	return v;
}

void stage_batch(crfgpu_ctx* h, uint32_t n_utt, const uint32_t* off, const float* ftrs, const uint32_t* labs, const float* ftrs2 = nullptr) {
	const crfgpu_config& c = h->cfg;
	validate_offsets(n_utt, off, ftrs);
	if (c.n_base_ftrs2 && !ftrs2) throw ApiError(CRFGPU_ERR_ARG, "the model joins a second feature stream: use the crfgpu_*2 entry points and pass it");
	const uint32_t N = n_utt ? off[n_utt] : 0;
	cudaStream_t s = h->stream;
	h->n_utt = n_utt; h->N = N; h->have_labels = labs != nullptr;
	h->fwdbwd_done = h->viterbi_done = false; h->vit_score_ready = h->vit_rec_ready = false;
	h->x_tiles_valid = false;
	h->h_off.assign(off, off + n_utt + 1);
	{
		const size_t want = ((size_t)N * 6 + (size_t)n_utt * 24 + 65536) * sizeof(uint32_t);
		if (!h->ev_pin) CUDA_OK(cudaEventCreateWithFlags(&h->ev_pin, cudaEventDisableTiming));
		else CUDA_OK(cudaEventSynchronize(h->ev_pin));                  // the previous batch's table uploads have left the arena
		if (want > h->pin_cap) {
			if (h->pin) CUDA_OK(cudaFreeHost(h->pin));
			h->pin = nullptr; h->pin_cap = 0;
			CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&h->pin), want + want / 4));
			h->pin_cap = want + want / 4;
		}
		h->pin_used = 0;
	}

	upload_async(h, h->d_off, h->h_off);
	// per-frame index tables (position in the utterance, utterance, its length): one small kernel over the offsets instead of 12 bytes
	// per frame of host loops and uploads ahead of the first feature byte
	auto frame_tables = [&](bool with_t) {
		if (!N) return;
		if (with_t) h->d_frame_t.ensure(sizeof(uint32_t) * (size_t)N + 16);
		h->d_frame_utt.ensure(sizeof(uint32_t) * (size_t)N + 16); h->d_frame_len.ensure(sizeof(uint32_t) * (size_t)N + 16);
		launch_frame_tables(h->d_off.as<uint32_t>(), n_utt, N, with_t ? h->d_frame_t.as<uint32_t>() : nullptr, h->d_frame_utt.as<uint32_t>(),
		                    h->d_frame_len.as<uint32_t>(), s);
		check_kernel(h, 1);
	};
	const bool want_virt = h->virt && labs != nullptr;                   // training batches of eligible models stage the virtual-window form
	const bool prefetched = h->pre_valid && N && h->pre_ftrs == ftrs && h->pre_off == h->h_off && h->pre_virt == want_virt;
	bool joined_tabs = false;
	const bool tabs_ready0 = prefetched && !h->joined && h->pre_tabs && labs != nullptr && h->pre_labs == labs;   // index / label tables already on the device
	const bool had_tabs = h->pre_tabs;
	h->pre_valid = false; h->pre_tabs = false;
	h->x_virt_valid = want_virt; h->x_full_valid = !want_virt;
	const bool joined_ahead = h->joined && prefetched && h->pre_ftrs2 == ftrs2;      // this joined batch was read ahead (crfgpu_prefetch_train_batch2)
	if (h->joined && joined_ahead) {
		if (!h->ev_swap) CUDA_OK(cudaEventCreateWithFlags(&h->ev_swap, cudaEventDisableTiming));
		CUDA_OK(cudaEventRecord(h->ev_swap, s)); h->swap_marked = true;   // work queued before this point is the last reader of the set handed back
		std::swap(h->d_base, h->d_base2); std::swap(h->d_baseB, h->d_baseB2); std::swap(h->d_X, h->d_X2);
		std::swap(h->d_frame_t, h->d_frame_t2); std::swap(h->d_frame_utt, h->d_frame_utt2); std::swap(h->d_frame_len, h->d_frame_len2);
		if (h->pre_tiles && h->opt_tf_tiled) { std::swap(h->d_XtK, h->d_XtK2); std::swap(h->d_XtT, h->d_XtT2); h->x_tiles_valid = true; }
		if (had_tabs && labs != nullptr && h->pre_labs == labs) {
			std::swap(h->d_node_lab, h->d_node_lab2); std::swap(h->d_prev_lab, h->d_prev_lab2); std::swap(h->d_next_lab, h->d_next_lab2);
			joined_tabs = true;
		}
		CUDA_OK(cudaStreamWaitEvent(s, h->ev_pre_done, 0));
	} else if (h->joined) {
		// general window streams: both streams go to the device whole, one gather kernel builds the joined windows
		frame_tables(true);
		phase_begin(h, "expand");
		const bool tiles = labs != nullptr && N && (h->transftr || h->nodur_tf) && h->opt_tf_tiled;
		stage_joined(h, n_utt, N, ftrs, ftrs2, h->d_base, h->d_baseB, h->d_X, h->d_frame_t.as<uint32_t>(), h->d_frame_utt.as<uint32_t>(), s,
		             tiles ? &h->d_XtK : nullptr, tiles ? &h->d_XtT : nullptr);
		h->x_tiles_valid = tiles;
		phase_end(h, "expand");
	} else if (prefetched) {
		// this batch was copied and expanded by crfgpu_prefetch_batch while the previous one computed: take over its buffers
		if (!h->ev_swap) CUDA_OK(cudaEventCreateWithFlags(&h->ev_swap, cudaEventDisableTiming));
		CUDA_OK(cudaEventRecord(h->ev_swap, s)); h->swap_marked = true;   // work queued before this point is the last reader of the set handed back
		std::swap(h->d_base, h->d_base2); std::swap(h->d_X, h->d_X2); std::swap(h->d_frame_t, h->d_frame_t2);
		std::swap(h->d_bpad, h->d_bpad2); std::swap(h->d_Xa, h->d_Xa2);
		if (tabs_ready0) {
			std::swap(h->d_frame_utt, h->d_frame_utt2); std::swap(h->d_frame_len, h->d_frame_len2);
			std::swap(h->d_node_lab, h->d_node_lab2); std::swap(h->d_prev_lab, h->d_prev_lab2); std::swap(h->d_next_lab, h->d_next_lab2);
		}
		CUDA_OK(cudaStreamWaitEvent(s, h->ev_pre_done, 0));
		if (!tabs_ready0) frame_tables(false);
	} else {
		// what the window expansion needs goes through the copy engine ahead of the feature chunks
		frame_tables(true);
		h->d_base.ensure(sizeof(float) * (size_t)N * c.n_base_ftrs + 16);
		if (want_virt && N) { h->d_bpad.ensure(sizeof(float) * (size_t)N * h->Fp + 16); h->d_Xa.ensure(sizeof(float) * (size_t)N * c.max_dur * h->Wa + 16); }
		else if (c.max_dur > 1 && N) h->d_X.ensure(sizeof(float) * (size_t)N * c.max_dur * h->Wp + 16);
		if (N && !h->ev_ready) CUDA_OK(cudaEventCreate(&h->ev_ready));
		// a decode batch (no labels): the decoder's fp64 state scores of a chunk are launched as soon as the chunk has arrived, so the
		// scoring runs under the remaining H2D copies instead of behind them
		bool eager_rec = false; uint32_t u_next = 0, k_chunk = 0;
		std::function<void(uint32_t, uint32_t)> score_chunk = [&](uint32_t n0, uint32_t n1) {
			const Layout& m = h->lay;
			VitScoreParams vs{};
			vs.X = h->X() + (size_t)n0 * h->ldx(); vs.ldx = h->ldx(); vs.W = h->Wp; vs.sf0 = c.state_fidx_start; vs.nSf = m.nSf; vs.N = n1 - n0; vs.D = c.max_dur; vs.L = m.L;
			vs.frame_t = h->d_frame_t.as<uint32_t>() + n0; vs.Wd = h->d_Wd.as<double>(); vs.use_bias = c.use_state_bias; vs.bias_val = c.state_bias_val;
			vs.negS = h->d_negS.as<float>() + (size_t)n0 * c.max_dur * m.L;
			launch_vit_scores(vs, s); check_kernel(h, 1);
			if (!eager_rec) return;
			// ... and the recursion over the utterances of this chunk (chunks end at utterance boundaries), longest first
			uint32_t u1 = u_next;
			while (u1 < n_utt && off[u1] < n1) u1++;
			const uint32_t cnt = u1 - u_next;
			const size_t bytes = sizeof(uint32_t) * cnt, at = (h->pin_used + 255) & ~(size_t)255;
			if (at + bytes > h->pin_cap) throw ApiError(CRFGPU_ERR_CUDA, "internal: the page-locked table arena is too small for the decode order");
			uint32_t* ord = reinterpret_cast<uint32_t*>(h->pin + at);
			std::iota(ord, ord + cnt, u_next);
			std::stable_sort(ord, ord + cnt, [&](uint32_t a, uint32_t b) { return off[a + 1] - off[a] > off[b + 1] - off[b]; });
			h->pin_used = at + bytes;
			// a recursion lasts as long as its longest utterance however few utterances it walks: every chunk's launch gets a stream of
			// its own, so that it neither delays the next chunk's scores nor waits for the previous chunk's paths
			if (h->rec_stream.size() <= k_chunk) {
				cudaStream_t st; CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); h->rec_stream.push_back(st);
				cudaEvent_t e0, e1; CUDA_OK(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
				h->ev_scored.push_back(e0); h->ev_walked.push_back(e1);
			}
			cudaStream_t rs = h->rec_stream[k_chunk];
			CUDA_OK(cudaEventRecord(h->ev_scored[k_chunk], s));
			CUDA_OK(cudaStreamWaitEvent(rs, h->ev_scored[k_chunk], 0));
			CUDA_OK(cudaMemcpyAsync(h->d_vorder.as<uint32_t>() + u_next, ord, bytes, cudaMemcpyHostToDevice, rs));
			VitParams v = vit_params(h);
			v.order = h->d_vorder.as<uint32_t>() + u_next; v.n_utt = cnt;
			launch_viterbi(v, rs); check_kernel(h, 1);
			CUDA_OK(cudaEventRecord(h->ev_walked[k_chunk], rs));
			u_next = u1; k_chunk++;
		};
		const bool eager_vit = !labs && N && h->decode_ok && h->have_lambda && h->lay.nSf > 0;
		// the per-utterance recursion needs nothing but the state scores and the constant tables: it follows chunk by chunk as well
		// (not with per-frame transition tables, the group-sliced kernel, or a geometry viterbi_staged refuses)
		eager_rec = eager_vit && !c.use_trans_ftrs && !vit_wants_groups(h) && !(h->beam > 0.0 && h->lay.n_states != 1) && !getenv("CRFGPU_DP_TIMING") && h->opt_vit_eager;
		if (eager_vit) h->d_negS.ensure(sizeof(float) * (size_t)N * c.max_dur * h->lay.L + 16);
		if (eager_rec) { ensure_decode_buffers(h); h->d_vorder.ensure(sizeof(uint32_t) * (size_t)n_utt + 16); }
		copy_and_expand(h, n_utt, off, N, ftrs, h->d_base, h->d_X, h->d_frame_t, s, h->ev_ready, h->ev_chunk, "expand", 0, eager_vit ? &score_chunk : nullptr,
		                want_virt, &h->d_bpad, &h->d_Xa, eager_rec ? 8u : 4u);
		for (uint32_t k = 0; k < k_chunk; k++) CUDA_OK(cudaStreamWaitEvent(s, h->ev_walked[k], 0));      // the batch's paths: behind every chunk's recursion
		h->vit_score_ready = eager_vit; h->vit_rec_ready = eager_rec;
	}

	const bool tabs_ready = tabs_ready0 || joined_tabs;
	if (labs && !tabs_ready) {
		std::vector<uint32_t> node_lab, prev_lab, next_lab;
		build_label_tables(h, n_utt, off, labs, node_lab, prev_lab, next_lab);
		upload_async(h, h->d_node_lab, node_lab); upload_async(h, h->d_prev_lab, prev_lab);
		if (!next_lab.empty()) upload_async(h, h->d_next_lab, next_lab);
	}
	// utterances sorted by length (longest first) so the slots of one CTA finish together; reordering inside a
	// minibatch does not change the gradient sum
	std::vector<uint32_t> order(n_utt);
	std::iota(order.begin(), order.end(), 0u);
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return off[a + 1] - off[a] > off[b + 1] - off[b]; });
	int U = h->opt_slots;
	if (U != 1 && U != 2 && U != 4 && U != 8) {
		U = 1;
		while (U < 8 && n_utt / (U * 2) >= 148) U *= 2;
	}
	h->U = U;
	h->n_groups = (n_utt + U - 1) / U;
	std::vector<uint32_t> grp((size_t)h->n_groups * U, CRFGPU_LAB_BAD);
	// deal utterances round-robin over the groups so every group gets a similar mix of lengths?  No:
	// consecutive (similar-length) utterances share a CTA, which minimises idle slots.
	for (uint32_t i = 0; i < n_utt; i++) grp[i] = order[i];
	upload_async(h, h->d_grp, grp);
	{   // the same order padded to a multiple of 16 with LAB_BAD: lock-step batches of the group Viterbi, launch order of the per-utterance kernels
		std::vector<uint32_t> o16((n_utt + 15) / 16 * 16, CRFGPU_LAB_BAD);
		for (uint32_t i = 0; i < n_utt; i++) o16[i] = order[i];
		upload_async(h, h->d_order16, o16);
	}
	// cluster-resident lattice kernels: persistent clusters, utterances dealt longest-first to the least loaded
	// cluster; inside a cluster the list order is the order slots are (re)filled
	std::vector<uint32_t> cl_off, cl_list;
	h->cluster_ok = false; h->tc_ok = false; h->ks_ok = false; h->locksteps = 0;
	for (uint32_t u = 0; u < n_utt; u++) h->locksteps = std::max<uint64_t>(h->locksteps, off[u + 1] - off[u]);   // one chain per utterance unless a plan below deals lists
	// frame-level models with at most 64 labels run one warp per chain (crf_dp_frame.cu): no cluster plan, no slot lists
	h->frame_path = c.max_dur == 1 && h->Lt <= 64 && !h->tied && !h->nodur && !h->transftr && h->opt_dp_impl >= 2 && h->opt_frame_impl != 1;
	auto deal = [&](uint32_t ncl) {
		// utterances dealt longest-first to the least loaded cluster; inside a cluster the list order is the order slots are (re)filled
		std::vector<std::vector<uint32_t>> lists(ncl);
		std::vector<uint64_t> load(ncl, 0);
		for (uint32_t i = 0; i < n_utt; i++) {
			const uint32_t u = order[i];
			const uint32_t k = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
			lists[k].push_back(u); load[k] += off[u + 1] - off[u];
		}
		h->locksteps = load.empty() ? 0 : *std::max_element(load.begin(), load.end());
		cl_off.assign(1, 0); cl_list.clear();
		for (auto& l : lists) { cl_list.insert(cl_list.end(), l.begin(), l.end()); cl_off.push_back((uint32_t)cl_list.size()); }
		upload_async(h, h->d_cl_off, cl_off); upload_async(h, h->d_cl_list, cl_list);
	};
	const bool lattice_tc = h->train_ok && !h->frame_path && !h->nodur && !h->transftr && labs && n_utt && (uint64_t)N * h->Lp < (1ull << 32);   // the lane threads index the lattice arrays with 32 bits
	if (lattice_tc && h->opt_dp_impl == 3) {
		// contraction-sliced tensor-core kernels: 16 slots per cluster, E entirely in tensor memory
		KsDpPlan plan{};
		if (plan_ks_dp(h->Lt, c.max_dur, h->max_smem_optin, &plan)) {
			const int avail_cl = max_active_ks_clusters(plan);
			if (avail_cl > 0) {
				uint32_t ncl = std::min<uint32_t>((uint32_t)avail_cl, (n_utt + TC_DP_SLOTS - 1) / TC_DP_SLOTS);
				if (h->opt_max_clusters > 0) ncl = std::min<uint32_t>(ncl, (uint32_t)h->opt_max_clusters);   // tests: few clusters force slot refills on small batches
				deal(ncl * TC_DP_SLOTS);         // one list per slot, balanced over all slots of all clusters (longest first)
				h->ks_plan = plan; h->n_ks_clusters = ncl; h->ks_ok = true;
				if (getenv("CRFGPU_VERBOSE"))
					fprintf(stderr, "[crfgpu] contraction-sliced lattice plan: CS=%u CW=%u K=%u MT=%u tmem_cols=%u smem=%zu, %u of %d resident clusters, %u utterances\n",
					        plan.CS, plan.CW, plan.K, plan.MT, plan.tmem_cols, plan.smem, ncl, avail_cl, n_utt);
			}
		}
	}
	if (lattice_tc && !h->ks_ok && h->opt_dp_impl >= 2) {
		// output-sliced tensor-core kernels (crf_dp_tc.cu): geometries whose tiles do not fit tensor memory whole
		TcDpPlan plan{};
		if (plan_tc_dp(h->Lt, c.max_dur, h->max_smem_optin, &plan)) {
			const int avail_cl = max_active_tc_clusters(plan);
			if (avail_cl > 0) {
				uint32_t ncl = std::min<uint32_t>((uint32_t)avail_cl, (n_utt + TC_DP_SLOTS - 1) / TC_DP_SLOTS);
				if (h->opt_max_clusters > 0) ncl = std::min<uint32_t>(ncl, (uint32_t)h->opt_max_clusters);
				deal(ncl * TC_DP_SLOTS);
				h->tc_plan = plan; h->n_tc_clusters = ncl; h->tc_ok = true;
				if (getenv("CRFGPU_VERBOSE"))
					fprintf(stderr, "[crfgpu] tensor-core lattice plan: CS=%u CW=%u K=%u tmem_cols=%u smem=%zu, %u of %d resident clusters, %u utterances\n",
					        plan.CS, plan.CW, plan.K, plan.tmem_cols, plan.smem, ncl, avail_cl, n_utt);
			}
		}
	}
	if (h->train_ok && !h->frame_path && !h->nodur && !h->transftr && (h->opt_dp_impl == 1 || (h->opt_dp_impl >= 2 && !h->tc_ok && !h->ks_ok)) && labs && n_utt) {
		ClusterPlan plan{};
		int cap = h->opt_cluster_slots > 0 ? h->opt_cluster_slots : 32;
		if (plan_cluster_dp(h->Lt, c.max_dur, h->max_smem_optin, &plan, cap)) {
			int avail_cl = max_active_clusters(plan);
			// keep every cluster busy: shrink the slot count while there are fewer slot-loads than clusters
			while (h->opt_cluster_slots <= 0 && plan.UB > 4 && avail_cl > 0 && (n_utt + plan.UB - 1) / plan.UB < (uint32_t)avail_cl) {
				ClusterPlan smaller{};
				if (!plan_cluster_dp(h->Lt, c.max_dur, h->max_smem_optin, &smaller, plan.UB == 32 ? 16 : plan.UB == 16 ? 12 : plan.UB == 12 ? 8 : 4)) break;
				if (smaller.CS != plan.CS) break;
				plan = smaller; avail_cl = max_active_clusters(plan);
			}
			if (avail_cl > 0) {
				uint32_t ncl = std::min<uint32_t>((uint32_t)avail_cl, (n_utt + plan.UB - 1) / plan.UB);
				if (h->opt_max_clusters > 0) ncl = std::min<uint32_t>(ncl, (uint32_t)h->opt_max_clusters);
				deal(ncl);
				h->plan = plan; h->n_clusters = ncl; h->cluster_ok = true;
				if (getenv("CRFGPU_VERBOSE"))
					fprintf(stderr, "[crfgpu] cluster plan: CS=%u CW=%u UB=%d threads=%u smem=%zu, %u of %d resident clusters, %u utterances\n",
					        plan.CS, plan.CW, plan.UB, plan.threads, plan.smem, ncl, avail_cl, n_utt);
				h->d_xch.ensure(sizeof(float) * (size_t)ncl * 2 * plan.UB * h->Lp + 16);
				h->d_xmax.ensure(sizeof(float) * (size_t)ncl * 2 * plan.CS * plan.UB + 16);
			}
		}
	}
	std::vector<uint32_t> nd_grp, nd_batch;
	if (h->train_ok && h->nodur && !h->nodur_tf && labs && n_utt) {
		// native no_dur recursion: batches of NODUR_UT utterances of similar length advance in lock-step; batches are dealt
		// (longest first) to the least loaded of the co-resident CTA groups
		const uint32_t nb = (n_utt + NODUR_UT - 1) / NODUR_UT;
		const uint32_t ng = std::min<uint32_t>((uint32_t)h->nodur_groups_max, nb);
		std::vector<std::vector<uint32_t>> lists(ng);
		std::vector<uint64_t> load(ng, 0);
		for (uint32_t b = 0; b < nb; b++) {
			const uint32_t u0 = order[(size_t)b * NODUR_UT];
			const uint32_t k = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
			lists[k].push_back(b); load[k] += off[u0 + 1] - off[u0];
		}
		nd_grp.assign(1, 0);
		for (auto& l : lists) {
			for (uint32_t b : l)
				for (uint32_t i = 0; i < NODUR_UT; i++) nd_batch.push_back((size_t)b * NODUR_UT + i < n_utt ? order[(size_t)b * NODUR_UT + i] : CRFGPU_LAB_BAD);
			nd_grp.push_back((uint32_t)(nd_batch.size() / NODUR_UT));
		}
		h->n_nodur_groups = ng;
		upload_async(h, h->d_nd_grp, nd_grp); upload_async(h, h->d_nd_batch, nd_batch);
	}
	// no host-side wait: the table uploads come from the page-locked arena and the kernels of the step queue behind the expansion
	CUDA_OK(cudaEventRecord(h->ev_pin, s));
}

// [sum numer, sum logZ, n_utt, 0] behind the gradient so that a multi-GPU driver moves gradient and scalars
// with ONE all-reduce (CRF_Minibatch_GradAccumulator.cpp:277-298 sums them next to the gradient)
__global__ void tail_sums_kernel(const double* numer, const double* logZ, uint32_t n_utt, double* tail) {
	__shared__ double sn[32], sz[32];
	double a = 0.0, b = 0.0;
	for (uint32_t i = threadIdx.x; i < n_utt; i += blockDim.x) { a += numer[i]; b += logZ[i]; }
	for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
	if ((threadIdx.x & 31) == 0) { sn[threadIdx.x >> 5] = a; sz[threadIdx.x >> 5] = b; }
	__syncthreads();
	if (threadIdx.x == 0) {
		double ta = 0.0, tb = 0.0;
		for (uint32_t w = 0; w < blockDim.x / 32; w++) { ta += sn[w]; tb += sz[w]; }
		tail[0] = ta; tail[1] = tb; tail[2] = (double)n_utt;      // tail[3] = frames that failed the posterior-mass assertion (posterior_mass_kernel)
	}
}

// The window arrays of the staged batch in the form the next consumer reads, rebuilt from the resident base stream when the batch was
// staged in the other form (a decode call on a training batch, an option changed between staging and the call).
void ensure_windows(crfgpu_ctx* h, bool want_virt) {
	const crfgpu_config& c = h->cfg;
	if (c.max_dur == 1 || !h->N) return;
	cudaStream_t s = h->stream;
	if (want_virt && !h->x_virt_valid) {
		h->d_bpad.ensure(sizeof(float) * (size_t)h->N * h->Fp + 16); h->d_Xa.ensure(sizeof(float) * (size_t)h->N * c.max_dur * h->Wa + 16);
		launch_virtual_windows(h->d_base.as<float>(), h->d_frame_t.as<uint32_t>(), h->d_bpad.as<float>(), h->d_Xa.as<float>(), h->N, c.n_base_ftrs, h->Fp, c.max_dur, h->Wa, 0, h->N, s);
		check_kernel(h, 2);
		h->x_virt_valid = true;
	} else if (!want_virt && !h->x_full_valid) {
		h->d_X.ensure(sizeof(float) * (size_t)h->N * c.max_dur * h->Wp + 16);
		ExpandParams ep{h->d_base.as<float>(), h->d_frame_t.as<uint32_t>(), h->d_steps.as<uint32_t>(), h->d_X.as<float>(),
		                h->N, c.n_base_ftrs, c.max_dur, h->W, h->Wp, c.extract_seg_ftrs, 0, 0};
		launch_expand_windows(ep, h->N, s);
		check_kernel(h, 1);
		h->x_full_valid = true;
	}
}

DpParams dp_params(crfgpu_ctx* h) {
	const crfgpu_config& c = h->cfg;
	DpParams p{};
	p.L = h->Lt; p.Lp = h->Lp; p.D = c.max_dur; p.P = h->Lt / c.max_dur;
	p.n_groups = h->n_groups; p.grp_utt = h->d_grp.as<uint32_t>(); p.off = h->d_off.as<uint32_t>();
	p.S = h->d_S.as<float>(); p.E = h->d_E.as<float>(); p.ET = h->d_ET.as<float>();
	p.A = h->d_A.as<float>(); p.G = h->d_G.as<float>(); p.m = h->d_m.as<double>(); p.kappa = h->d_kappa.as<double>();
	p.bbase = h->d_bbase.as<double>(); p.Uvec = h->opt_keep_lattice ? h->d_Uvec.as<float>() : nullptr;
	p.logZ = h->d_logZ.as<double>(); p.Dm = h->d_Dm.as<float>(); p.R = h->d_R.as<float>();
	p.node_lab = h->d_node_lab.as<uint32_t>(); p.Mmax = h->Mmax; p.mass = nullptr;
	return p;
}

// transition-feature gradient: out[tidx(pair) + f] += sum_n Xd[n][pair] * x_n[tf0 + f] (+ the bias as a constant-1 column), pairs = I.
// Default: both operands split into bf16 hi / lo and tiled once (launch_tile_mn), the product fed by bulk copies; option tf_tiled 0:
// the register-staged kernel on the fp32 arrays.
// transition scores of every frame: M[n][pair] = x_n(trans slice) . lambda_t[pair] + bias[pair] (-inf on the pairs an N-state map lacks)
void trans_score_gemm(crfgpu_ctx* h, uint32_t N, uint32_t I, uint32_t Lq, cudaStream_t s) {
	const crfgpu_config& c = h->cfg; const Layout& m = h->lay;
	if (h->opt_tf_tiled) {
		if (!h->x_tiles_valid) { tile_trans_slice(h, N, h->X(), h->d_XtK, h->d_XtT, s); h->x_tiles_valid = true; }      // (batches staged without the tiles)
		TiledScoreParams t{};
		t.At = h->d_XtK.as<unsigned char>(); t.Bt = h->d_WtrT.as<unsigned char>(); t.bias = h->d_tbias.as<float>();
		t.C = h->d_Mall.as<float>(); t.ldc = Lq; t.M = N; t.Ncols = I; t.K = m.nTf;
		CUDA_OK(launch_score_gemm_tiled(t, s)); check_kernel(h, 1);
		return;
	}
	ScoreGemmParams g{};
	g.A = h->X() + c.trans_fidx_start; g.lda = h->ldx(); g.B = h->d_Wtr.as<float>(); g.ldb = m.nTf; g.bias = h->d_tbias.as<float>();
	g.C = h->d_Mall.as<float>(); g.ldc = Lq; g.M = N; g.Ncols = I; g.K = m.nTf;
	CUDA_OK(launch_score_gemm_tc(g, s)); check_kernel(h, 1);
}

void trans_gradient_gemm(crfgpu_ctx* h, uint32_t N, uint32_t I, uint32_t Lq, cudaStream_t s) {
	const crfgpu_config& c = h->cfg; const Layout& m = h->lay;
	const uint32_t nTf = m.nTf, J = nTf + (c.use_trans_bias ? 1 : 0), ones = c.use_trans_bias ? nTf : 0xffffffffu;
	if (h->opt_tf_tiled) {
		h->d_XdT.ensure(tiled_operand_bytes(N, I, 128) + 16);
		CUDA_OK(launch_tile_mn(h->d_Xd.as<float>(), Lq, I, 0xffffffffu, N, true, h->d_XdT.as<unsigned char>(), s)); check_kernel(h, 1);
		if (!h->x_tiles_valid) { tile_trans_slice(h, N, h->X(), h->d_XtK, h->d_XtT, s); h->x_tiles_valid = true; }
		TiledReduceParams t{};
		t.At = h->d_XdT.as<unsigned char>(); t.Bt = h->d_XtT.as<unsigned char>(); t.N = N; t.I = I; t.J = J;
		t.ones_col = ones; t.scale = 1.0; t.ones_scale = c.trans_bias_val; t.row_idx = h->d_tidx.as<uint32_t>(); t.out = h->d_grad.as<double>();
		CUDA_OK(launch_reduce_gemm_tiled(t, s)); check_kernel(h, 1);
		return;
	}
	ReduceGemmParams r{};
	r.A = h->d_Xd.as<float>(); r.lda = Lq; r.a_row_shift = 0;
	r.B = h->X() + c.trans_fidx_start; r.ldb = h->ldx();
	r.n0 = 0; r.n1 = N; r.I = I; r.J = J;
	r.ones_col = ones;
	r.scale = 1.0; r.ones_scale = c.trans_bias_val; r.mode = 0;
	r.row_idx = h->d_tidx.as<uint32_t>();
	r.out = h->d_grad.as<double>(); r.k_slab = h->opt_k_slab_tc;
	CUDA_OK(launch_reduce_gemm_tc(r, false, s)); check_kernel(h, 1);
}

void fwdbwd_staged(crfgpu_ctx* h) {
	require_train(h);
	if (!h->have_labels) throw ApiError(CRFGPU_ERR_ARG, "training needs frame labels");
	const crfgpu_config& c = h->cfg; const Layout& m = h->lay;
	const uint32_t N = h->N, L = h->Lt, Lp = h->Lp, D = c.max_dur, P = L / D, nSf = m.nSf;
	cudaStream_t s = h->stream;
	const size_t NL = (size_t)N * Lp;
	const size_t NV = h->nodur ? (size_t)N * h->Pp : NL;         // forward/backward vectors: P wide on the native no_dur path
	h->d_S.ensure(sizeof(float) * NL + 16); h->d_A.ensure(sizeof(float) * NV + 16); h->d_G.ensure(sizeof(float) * NV + 16);
	h->d_Dm.ensure(sizeof(float) * NL + 16); h->d_R.ensure(sizeof(float) * (NV + (h->nodur ? h->Pp : 0)) + 16);
	if (h->nodur) h->d_LB.ensure(sizeof(float) * NV + 16);
	if (h->opt_keep_lattice) {
		if (h->nodur) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "keep_lattice / crfgpu_fetch_alpha_beta is not implemented for the native stdseg_no_dur* recursion");
		h->d_Uvec.ensure(sizeof(float) * NL + 16);
	}
	h->d_m.ensure(sizeof(double) * (size_t)N + 16); h->d_kappa.ensure(sizeof(double) * (size_t)N + 16); h->d_bbase.ensure(sizeof(double) * (size_t)N + 16);
	h->d_logZ.ensure(sizeof(double) * (size_t)h->n_utt + 16); h->d_numer.ensure(sizeof(double) * (size_t)h->n_utt + 16);
	h->d_grad.ensure(sizeof(double) * ((size_t)m.len + 4) + 16);
	CUDA_OK(cudaMemsetAsync(h->d_grad.p, 0, sizeof(double) * ((size_t)m.len + 4), s));
	CUDA_OK(cudaMemsetAsync(h->d_numer.p, 0, sizeof(double) * (size_t)h->n_utt, s));
	if (!N) { h->fwdbwd_done = true; return; }

	// K1: state scores.  TMA-fed kernel: all durations in one launch, per-duration maxima fused; otherwise one GEMM per duration block
	const bool virt = h->virt;                                    // virtual windows: sampled-frame blocks as row shifts of the padded base stream
	ensure_windows(h, virt);
	const bool tma = virt || (h->opt_gemm_impl == 2 && (h->opt_tma_mask & 1) && D > 1 && nSf > 0 && tma_gemm_eligible(h->X() + c.state_fidx_start, D, h->Wp, c.state_fidx_start));
	bool smax_done = false;
	if (nSf == 0) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "models without state features are not implemented on the device");
	// the empirical counts and numerators (labels, windows and lambda only) do not wait for anything the recursions produce: on a side
	// stream they run beside the score GEMM and the lattice kernels (which leave 28 SMs idle) instead of behind the state gradient
	auto make_emp = [&]() {
		EmpiricalParams e{};
		e.X = h->X(); e.ldx = h->ldx(); e.W = h->Wp; e.sf0 = c.state_fidx_start; e.nSf = nSf;
		e.node_lab = h->d_node_lab.as<uint32_t>(); e.prev_lab = h->d_prev_lab.as<uint32_t>(); e.frame_utt = h->d_frame_utt.as<uint32_t>();
		e.N = N; e.L = L; e.P = P; e.tL = h->nodur ? P : L; e.lambda = h->d_lambda.as<double>(); e.sidx = h->d_sidx.as<uint32_t>(); e.tidx = h->d_tidx.as<uint32_t>();
		e.use_state_bias = c.use_state_bias; e.use_trans_bias = c.use_trans_bias;
		e.state_bias_val = c.state_bias_val; e.trans_bias_val = c.trans_bias_val;
		e.grad = h->d_grad.as<double>(); e.numer = h->d_numer.as<double>();
		if (virt) { e.virt = 1; e.F = c.n_base_ftrs; e.D = D; e.base = h->d_base.as<float>(); e.steps = h->d_steps.as<uint32_t>(); e.X = h->d_Xa.as<float>(); e.W = h->Wa; }
		return e;
	};
	const bool emp_early = h->opt_aux_empirical && !h->nodur_tf && !h->transftr;
	if (emp_early) {
		if (!h->aux_stream) {
			CUDA_OK(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
			CUDA_OK(cudaEventCreateWithFlags(&h->ev_aux_go, cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&h->ev_aux_done, cudaEventDisableTiming));
		}
		CUDA_OK(cudaEventRecord(h->ev_aux_go, s));                  // behind the memsets of the gradient / numerators and the windows
		CUDA_OK(cudaStreamWaitEvent(h->aux_stream, h->ev_aux_go, 0));
		launch_empirical(make_emp(), h->aux_stream); check_kernel(h, 1);
		CUDA_OK(cudaEventRecord(h->ev_aux_done, h->aux_stream));
	}
	phase_begin(h, "score");
	if (tma) {
		ScoreTmaParams g{};
		g.Bt = h->d_Wt.as<unsigned char>(); g.bias = h->d_bias.as<float>(); g.C = h->d_S.as<float>(); g.ldc = Lp;
		g.M = N; g.P = P; g.K = nSf; g.D = D; g.n_chunks = score_tma_chunks(nSf); g.ntile = (P + 63) / 64;
		g.frame_t = h->d_frame_t.as<uint32_t>(); g.shared_w = h->nodur ? 1u : 0u; g.a_from_tmem = (h->opt_tma_mask & 8) ? 1u : 0u;
		if ((h->tc_ok || h->ks_ok || h->nodur) && g.ntile == 1) { h->d_smaxd.ensure(sizeof(float) * (size_t)N * D + 16); g.smaxd = h->d_smaxd.as<float>(); smax_done = true; }
		if (virt) {
			g.virt = 1; g.cpb = h->Fp / 32; g.base2 = h->d_bpad.as<float>(); g.steps = h->d_steps.as<uint32_t>();
			g.bias = h->d_bias_dy.as<float>(); g.n_chunks = 5 * g.cpb + h->Wa / 32;
			CUDA_OK(launch_score_gemm_tma(h->d_Xa.as<float>(), h->Wa, g, s));
		} else
		CUDA_OK(launch_score_gemm_tma(h->X() + c.state_fidx_start, h->Wp, g, s));
		check_kernel(h, 1);
	} else for (uint32_t d = 0; d < D; d++) {
		ScoreGemmParams g{};
		g.A = h->X() + (size_t)d * h->Wp + c.state_fidx_start; g.lda = h->ldx();
		// the lattice label (d,y) has its own row of the device tables (tied models repeat phone y's weights; the native no_dur path shares P rows)
		g.B = h->d_Ws.as<float>() + (h->nodur ? 0 : (size_t)d * P * nSf); g.ldb = nSf;
		g.bias = h->d_bias.as<float>() + (h->nodur ? 0 : (size_t)d * P);
		g.C = h->d_S.as<float>() + (size_t)d * P; g.ldc = Lp;
		g.M = N; g.Ncols = P; g.K = nSf;
		if (h->opt_gemm_impl >= 1) CUDA_OK(launch_score_gemm_tc(g, s)); else launch_score_gemm(g, s);
		check_kernel(h, 1);
	}
	phase_end(h, "score");

	if (h->transftr) {
		// frame-level model with transition FEATURES: the L*L transition scores of every frame are one more tensor-core GEMM, the
		// recursions stream them (crf_dp_transftr.cu), and both gradients (with their empirical counts) are reduce-GEMMs
		const uint32_t Lq = (L * L + 3) / 4 * 4, tf0 = c.trans_fidx_start, nTf = m.nTf;
		h->d_Mall.ensure(sizeof(float) * (size_t)N * Lq + 16); h->d_Xd.ensure(sizeof(float) * (size_t)N * Lq + 16);
		phase_begin(h, "forward");
		trans_score_gemm(h, N, L * L, Lq, s);
		h->d_Eall.ensure(sizeof(float) * (size_t)N * Lq + 16); h->d_rowmax.ensure(sizeof(float) * (size_t)N + 16);
		launch_transftr_exp(h->d_Mall.as<float>(), h->d_Eall.as<float>(), h->d_rowmax.as<float>(), N, L * L, Lq, s); check_kernel(h, 1);
		TransFtrParams q{};
		q.E = h->d_Eall.as<float>(); q.rowmax = h->d_rowmax.as<float>();
		q.L = L; q.Lp = Lp; q.Lq = Lq; q.n_utt = h->n_utt; q.off = h->d_off.as<uint32_t>();
		q.S = h->d_S.as<float>(); q.M = h->d_Mall.as<float>(); q.A = h->d_A.as<float>(); q.rho = h->d_m.as<double>();
		q.logZ = h->d_logZ.as<double>(); q.numer = h->d_numer.as<double>(); q.Dm = h->d_Dm.as<float>(); q.Xd = h->d_Xd.as<float>();
		q.labs = h->d_node_lab.as<uint32_t>(); q.tidx = h->d_tidx.as<uint32_t>();
		CUDA_OK(launch_transftr_dp(false, q, s)); check_kernel(h, 1);
		phase_end(h, "forward");
		phase_begin(h, "backward");
		CUDA_OK(launch_transftr_dp(true, q, s)); check_kernel(h, 1);
		phase_end(h, "backward");
		phase_begin(h, "xi");
		{
			// transition weights: out[tidx(p,c) + f] += sum_n ([ref pair] - xi_n[p][c]) * x_n[tf0 + f]   (computeTransExpF, :197-223)
			trans_gradient_gemm(h, N, L * L, Lq, s);
		}
		phase_end(h, "xi");
		phase_begin(h, "grad");
		{
			ReduceGemmParams r{};
			r.A = h->d_Dm.as<float>(); r.lda = Lp; r.a_row_shift = 0;
			r.B = h->X() + c.state_fidx_start; r.ldb = h->ldx();
			r.n0 = 0; r.n1 = N; r.I = L; r.J = nSf + (c.use_state_bias ? 1 : 0);
			r.ones_col = c.use_state_bias ? nSf : 0xffffffffu;
			r.scale = 1.0; r.ones_scale = c.state_bias_val; r.mode = 0;
			r.row_idx = h->d_sidx.as<uint32_t>();
			r.out = h->d_grad.as<double>(); r.k_slab = h->opt_k_slab_tc;
			CUDA_OK(launch_reduce_gemm_tc(r, true, s)); check_kernel(h, 1);
		}
		tail_sums_kernel<<<1, 256, 0, s>>>(h->d_numer.as<double>(), h->d_logZ.as<double>(), h->n_utt, h->d_grad.as<double>() + m.len);
		check_kernel(h, 1);
		phase_end(h, "grad");
		h->fwdbwd_done = true;
		return;
	}
	DpParams p = dp_params(h);
	// frames per CTA in the reduce-GEMMs.  The defaults are tuned on cfg4 (10 duration blocks of CTAs); a frame-level model has one
	// block, so the same slabs would leave most SMs idle (cfg2: 17 and 68 CTAs): one to two slabs per SM instead
	uint32_t ks_xi = h->opt_k_slab_xi, ks_tc = h->opt_k_slab_tc;
	if (D == 1 && N) {
		ks_xi = std::min(ks_xi, std::max(512u, (N / 148u + 31u) / 32u * 32u));
		ks_tc = std::min(ks_tc, std::max(256u, (N / 296u + 31u) / 32u * 32u));
	}
	const uint32_t Lq = (P * P + 3) / 4 * 4;        // nodur_tf: row stride of the per-frame transition scores
	if (h->nodur_tf) {
		const uint32_t tf0 = c.trans_fidx_start, nTf = m.nTf;
		h->d_Mall.ensure(sizeof(float) * (size_t)N * Lq + 16); h->d_Xd.ensure(sizeof(float) * (size_t)N * Lq + 16);
		phase_begin(h, "forward");
		phase_begin(h, "trans_score");      // (nested in "forward": the transition-score GEMM and its exp pre-pass on their own)
		trans_score_gemm(h, N, P * P, Lq, s);      // M_n[y'][y] from the duration-1 window of frame n
		h->d_Eall.ensure(sizeof(float) * (size_t)N * Lq + 16); h->d_rowmax.ensure(sizeof(float) * (size_t)N + 16);
		launch_transftr_exp(h->d_Mall.as<float>(), h->d_Eall.as<float>(), h->d_rowmax.as<float>(), N, P * P, Lq, s); check_kernel(h, 1);
		phase_end(h, "trans_score");
		NodurTfParams q{};
		q.E = h->d_Eall.as<float>(); q.rowmax = h->d_rowmax.as<float>();
		q.P = P; q.Pp = h->Pp; q.D = D; q.Lp = Lp; q.Lq = Lq; q.n_utt = h->n_utt; q.off = h->d_off.as<uint32_t>();
		q.S = h->d_S.as<float>(); q.M = h->d_Mall.as<float>(); q.A = h->d_A.as<float>(); q.LG = h->d_G.as<float>(); q.rho = h->d_m.as<double>();
		q.logZ = h->d_logZ.as<double>(); q.numer = h->d_numer.as<double>(); q.Dm = h->d_Dm.as<float>(); q.Xd = h->d_Xd.as<float>();
		q.node_lab = h->d_node_lab.as<uint32_t>(); q.next_lab = h->d_next_lab.as<uint32_t>(); q.tidx = h->d_tidx.as<uint32_t>();
		q.LB = h->d_LB.as<float>(); q.kappa = h->d_kappa.as<double>();
		CUDA_OK(launch_nodur_tf_dp(false, q, s)); check_kernel(h, 1);
		phase_end(h, "forward");
		phase_begin(h, "backward");
		CUDA_OK(launch_nodur_tf_dp(true, q, s)); check_kernel(h, 1);
		{   // Dm = [ref] - gamma: the posterior pass of the native no_dur path (this recursion keeps no Mmax in its scales)
			NodurParams pp{};
			pp.P = P; pp.Pp = h->Pp; pp.D = D; pp.Lp = Lp; pp.S = h->d_S.as<float>(); pp.LG = h->d_G.as<float>(); pp.LB = h->d_LB.as<float>();
			pp.rho = h->d_m.as<double>(); pp.kappa = h->d_kappa.as<double>(); pp.logZ = h->d_logZ.as<double>(); pp.Mmax = 0.0;
			pp.node_lab = h->d_node_lab.as<uint32_t>(); pp.Dm = h->d_Dm.as<float>();
			launch_nodur_post(pp, h->d_frame_t.as<uint32_t>(), h->d_frame_utt.as<uint32_t>(), N, s); check_kernel(h, 1);
		}
		phase_end(h, "backward");
	} else if (h->nodur) {
		h->d_smaxd.ensure(sizeof(float) * (size_t)N * D + 16);
		const uint32_t npt = (P + 31) / 32, Pk = npt * 32;
		h->d_nd_xch.ensure(sizeof(float) * (size_t)h->n_nodur_groups * 2 * ((size_t)Pk + npt) * NODUR_UT + 16);
		h->d_nd_ctr.ensure(sizeof(uint32_t) * (size_t)h->n_nodur_groups + 16);
		NodurParams q{};
		q.P = P; q.Pp = h->Pp; q.D = D; q.Lp = Lp; q.n_groups = h->n_nodur_groups; q.npt = npt;
		q.grp_off = h->d_nd_grp.as<uint32_t>(); q.batch_utt = h->d_nd_batch.as<uint32_t>(); q.off = h->d_off.as<uint32_t>();
		q.S = h->d_S.as<float>(); q.smaxd = h->d_smaxd.as<float>(); q.E = h->d_E.as<float>(); q.ET = h->d_ET.as<float>(); q.Mmax = h->Mmax;
		q.A = h->d_A.as<float>(); q.LG = h->d_G.as<float>(); q.rho = h->d_m.as<double>(); q.logZ = h->d_logZ.as<double>();
		q.LB = h->d_LB.as<float>(); q.R = h->d_R.as<float>(); q.Dm = h->d_Dm.as<float>(); q.node_lab = h->d_node_lab.as<uint32_t>();
		q.xch = h->d_nd_xch.as<float>(); q.ctr = h->d_nd_ctr.as<uint32_t>(); q.kappa = h->d_kappa.as<double>();
		static DevBuf ndbg; const bool ntiming = getenv("CRFGPU_DP_TIMING") != nullptr;
		if (ntiming) { ndbg.ensure(16 * 8); q.dbg = ndbg.as<unsigned long long>(); }
		auto nreport = [&](const char* what) {
			unsigned long long v[16]; CUDA_OK(cudaMemcpyAsync(v, ndbg.p, sizeof(v), cudaMemcpyDeviceToHost, s)); CUDA_OK(cudaStreamSynchronize(s));
			const double n = v[8] ? (double)v[8] : 1.0;
			fprintf(stderr, "[crfgpu] no_dur %s: %.0f steps; cycles/step: operand requests %.0f phaseA %.0f barrier %.0f partial-sum requests %.0f store+sync %.0f reduce+scales+prefetch %.0f multiply %.0f other %.0f\n",
			        what, n, v[0] / n, v[1] / n, v[2] / n, v[3] / n, v[4] / n, v[5] / n, v[6] / n, v[7] / n);
		};
		phase_begin(h, "forward");
		if (!smax_done) { launch_block_max(h->d_S.as<float>(), h->d_frame_t.as<uint32_t>(), h->d_smaxd.as<float>(), N, Lp, P, D, s); check_kernel(h, 1); }
		CUDA_OK(launch_nodur_dp(false, q, s)); check_kernel(h, 1);
		if (ntiming) nreport("forward");
		phase_end(h, "forward");
		phase_begin(h, "backward");
		CUDA_OK(launch_nodur_dp(true, q, s)); check_kernel(h, 1);
		if (ntiming) nreport("backward");
		launch_nodur_post(q, h->d_frame_t.as<uint32_t>(), h->d_frame_utt.as<uint32_t>(), N, s); check_kernel(h, 1);      // Dm = [ref] - gamma
		phase_end(h, "backward");
	} else if (h->frame_path) {
		// frame-level models with at most 64 labels: one warp per utterance, the transition matrix in registers (crf_dp_frame.cu);
		// d_grp holds the utterances longest first, so the four warps of a CTA carry similar lengths
		if (h->opt_frame_impl == 2) {     // the two chains one after the other, posteriors fused into the backward pass
			phase_begin(h, "forward");
			CUDA_OK(launch_frame_dp(false, p, h->d_grp.as<uint32_t>(), h->n_utt, s)); check_kernel(h, 1);
			phase_end(h, "forward");
			phase_begin(h, "backward");
			CUDA_OK(launch_frame_dp(true, p, h->d_grp.as<uint32_t>(), h->n_utt, s)); check_kernel(h, 1);
			phase_end(h, "backward");
		} else {
			// the alpha and beta chains of an utterance are independent given the scores: both run in ONE launch (the step is bound by the
			// longest utterance's chain, not by throughput), then the posteriors of all frames are formed without any recursion
			h->d_Uvec.ensure(sizeof(float) * NL + 16); p.Uvec = h->d_Uvec.as<float>();
			phase_begin(h, "forward");
			CUDA_OK(launch_frame_dp_pair(p, h->d_grp.as<uint32_t>(), h->n_utt, s)); check_kernel(h, 1);
			phase_end(h, "forward");
			phase_begin(h, "backward");
			CUDA_OK(launch_frame_post(p, h->d_frame_t.as<uint32_t>(), h->d_frame_utt.as<uint32_t>(), N, s)); check_kernel(h, 1);
			phase_end(h, "backward");
		}
	} else if (h->ks_ok) {
		h->d_smaxd.ensure(sizeof(float) * (size_t)N * D + 16);
		KsDpParams kp{};
		static_cast<DpParams&>(kp) = p;
		kp.CS = h->ks_plan.CS; kp.CW = h->ks_plan.CW; kp.K = h->ks_plan.K; kp.MT = h->ks_plan.MT; kp.tmem_cols = h->ks_plan.tmem_cols;
		kp.recv_off = h->ks_plan.recv_off; kp.vbuf_off = h->ks_plan.vbuf_off; kp.ps_off = h->ks_plan.ps_off; kp.ctl_off = h->ks_plan.ctl_off;
		kp.n_clusters = h->n_ks_clusters; kp.cl_off = h->d_cl_off.as<uint32_t>(); kp.cl_list = h->d_cl_list.as<uint32_t>();
		kp.smaxd = h->d_smaxd.as<float>();
		static DevBuf kdbg2; const bool timing = getenv("CRFGPU_DP_TIMING") != nullptr;
		if (timing) { kdbg2.ensure(64 * 8); }
		auto report = [&](const char* what) {
			unsigned long long v[32]; CUDA_OK(cudaMemcpyAsync(v, kdbg2.p, sizeof(v), cudaMemcpyDeviceToHost, s)); CUDA_OK(cudaStreamSynchronize(s));
			const double n = v[15] ? (double)v[15] : 1.0;
			fprintf(stderr, "[crfgpu] %s (contraction-sliced): %.0f steps; cycles/step MMA warp: scales-wait+issue %.0f mma-wait %.0f scatter-wait+push+reduce-wait %.0f tile-wait+push %.0f | bookkeeping: sums-wait %.0f scales %.0f schedule+prefetch %.0f | lanes: scales-wait+loads %.0f acc-wait+scatter %.0f recv-wait+sum %.0f epilogue %.0f\n",
			        what, n, v[1] / n, v[0] / n, v[2] / n, v[3] / n, v[4] / n, v[5] / n, v[6] / n, v[7] / n, v[8] / n, v[10] / n, v[9] / n);
		};
		kp.dbg = timing ? kdbg2.as<unsigned long long>() : nullptr;
		phase_begin(h, "forward");
		if (!smax_done) { launch_block_max(h->d_S.as<float>(), h->d_frame_t.as<uint32_t>(), h->d_smaxd.as<float>(), N, Lp, P, D, s); check_kernel(h, 1); }
		if (timing) CUDA_OK(cudaMemsetAsync(kdbg2.p, 0, 64 * 8, s));
		CUDA_OK(launch_ks_dp(false, kp, h->ks_plan, s)); check_kernel(h, 1);
		if (timing) report("forward");
		phase_end(h, "forward");
		phase_begin(h, "backward");
		if (timing) CUDA_OK(cudaMemsetAsync(kdbg2.p, 0, 64 * 8, s));
		CUDA_OK(launch_ks_dp(true, kp, h->ks_plan, s)); check_kernel(h, 1);
		if (timing) report("backward");
		phase_end(h, "backward");
	} else if (h->tc_ok) {
		h->d_smaxd.ensure(sizeof(float) * (size_t)N * D + 16);
		TcDpParams tp{};
		static_cast<DpParams&>(tp) = p;
		tp.CS = h->tc_plan.CS; tp.CW = h->tc_plan.CW; tp.K = h->tc_plan.K; tp.tmem_cols = h->tc_plan.tmem_cols; tp.ctl_off = h->tc_plan.ctl_off;
		tp.n_clusters = h->n_tc_clusters; tp.cl_off = h->d_cl_off.as<uint32_t>(); tp.cl_list = h->d_cl_list.as<uint32_t>();
		tp.smaxd = h->d_smaxd.as<float>();
		static DevBuf dbg; const bool timing = getenv("CRFGPU_DP_TIMING") != nullptr;
		if (timing) { dbg.ensure(64 * 8); }
		phase_begin(h, "forward");
		if (!smax_done) { launch_block_max(h->d_S.as<float>(), h->d_frame_t.as<uint32_t>(), h->d_smaxd.as<float>(), N, Lp, P, D, s); check_kernel(h, 1); }
		auto report = [&](const char* what) {
			unsigned long long v[32]; CUDA_OK(cudaMemcpyAsync(v, dbg.p, sizeof(v), cudaMemcpyDeviceToHost, s)); CUDA_OK(cudaStreamSynchronize(s));
			const double n = v[15] ? (double)v[15] : 1.0;
			fprintf(stderr, "[crfgpu] %s: %.0f steps; cycles/step MMA warp: scales+gather-wait %.0f issue %.0f tile-wait+copy %.0f (%.0f) | bookkeeping: gather-wait %.0f scales %.0f schedule+prefetch %.0f | lanes: scales-wait+loads %.0f acc-wait %.0f epilogue %.0f (%.0f)\n",
			        what, n, v[0] / n, v[1] / n, v[2] / n, v[3] / n, v[4] / n, v[5] / n, v[6] / n, v[7] / n, v[8] / n, v[9] / n, v[10] / n);
		};
		tp.dbg = timing ? dbg.as<unsigned long long>() : nullptr;
		if (timing) CUDA_OK(cudaMemsetAsync(dbg.p, 0, 64 * 8, s));
		CUDA_OK(launch_tc_dp(false, tp, h->tc_plan, s)); check_kernel(h, 1);
		if (timing) report("forward");
		phase_end(h, "forward");
		phase_begin(h, "backward");
		if (timing) CUDA_OK(cudaMemsetAsync(dbg.p, 0, 64 * 8, s));
		CUDA_OK(launch_tc_dp(true, tp, h->tc_plan, s)); check_kernel(h, 1);
		if (timing) report("backward");
		phase_end(h, "backward");
	} else if (h->cluster_ok) {
		ClusterDpParams cp{};
		static_cast<DpParams&>(cp) = p;
		cp.CS = h->plan.CS; cp.CW = h->plan.CW; cp.CWp = h->plan.CWp; cp.CWt = h->plan.CWt; cp.n_clusters = h->n_clusters;
		cp.cl_off = h->d_cl_off.as<uint32_t>(); cp.cl_list = h->d_cl_list.as<uint32_t>();
		cp.xch = h->d_xch.as<float>(); cp.xmax = h->d_xmax.as<float>();
		phase_begin(h, "forward");
		CUDA_OK(launch_cluster_dp(false, cp, h->plan, s)); check_kernel(h, 1);
		phase_end(h, "forward");
		phase_begin(h, "backward");
		CUDA_OK(launch_cluster_dp(true, cp, h->plan, s)); check_kernel(h, 1);
		phase_end(h, "backward");
	} else {
		phase_begin(h, "forward");
		launch_forward(p, h->U, s); check_kernel(h, 1);
		phase_end(h, "forward");
		phase_begin(h, "backward");
		launch_backward(p, h->U, s); check_kernel(h, 1);
		phase_end(h, "backward");
	}

	// K4: expected-minus-empirical counts as two families of reduce-GEMMs
	phase_begin(h, "xi");
	const uint32_t Lv = h->nodur ? h->Pp : Lp;                  // row stride of the forward / right-factor vectors
	const bool lat_tma = h->opt_gemm_impl == 2 && lattice_tma_eligible(h->d_A.as<float>(), Lv) && lattice_tma_eligible(h->d_R.as<float>(), Lv) &&
	                     lattice_tma_eligible(h->d_Dm.as<float>(), Lp);
	if (virt && !(lat_tma && tma)) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "internal: the virtual-window form needs the TMA-fed gradient kernels");
	if (h->nodur && !(lat_tma && tma)) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "the native stdseg_no_dur* path needs the TMA-fed gradient kernels (gemm_impl 2)");
	if (h->nodur_tf) {
		// transition weights: out[tidx(y',y) + f] += sum_n ([ref pair] - xi)[n][y'][y] * x_{n,1}[tf0 + f]  (duration-1 window of frame n)
		trans_gradient_gemm(h, N, P * P, Lq, s);
	} else if (h->nodur) {
		if (c.use_trans_bias && N > 1) {
			// xi_t[y',y] = a_t[y'] E[y'][y] R_{t+1}[y]: one duration block of P columns with row shift 1
			FrameGemmParams x{};
			x.N = N; x.P = P; x.D = 1; x.ntile = (P + FRAME_GEMM_TILE - 1) / FRAME_GEMM_TILE; x.k_slab = h->opt_k_slab_xi; x.Mext = P; x.ones_col = 0xffffffffu;
			x.scale = -c.trans_bias_val; x.pair_idx = h->d_tidx.as<uint32_t>(); x.L = P; x.Ew = h->d_E.as<float>(); x.e_ld = h->Pp; x.out = h->d_grad.as<double>();
			x.a_from_tmem = (h->opt_tma_mask & 32) ? 1u : 0u;
			CUDA_OK(launch_xi_gemm_tma(h->d_A.as<float>(), h->d_R.as<float>(), h->Pp, x, s)); check_kernel(h, 1);
		}
	} else
	if (c.use_trans_bias && lat_tma && (h->opt_tma_mask & 4) && N > 1) {
		FrameGemmParams x{};
		x.N = N; x.P = P; x.D = D; x.ntile = (P + FRAME_GEMM_TILE - 1) / FRAME_GEMM_TILE; x.k_slab = ks_xi; x.Mext = L; x.ones_col = 0xffffffffu;
		x.scale = -c.trans_bias_val; x.pair_idx = h->d_tidx.as<uint32_t>(); x.L = L; x.Ew = h->d_E.as<float>(); x.e_ld = Lp; x.out = h->d_grad.as<double>();
		x.a_from_tmem = (h->opt_tma_mask & 32) ? 1u : 0u;
		CUDA_OK(launch_xi_gemm_tma(h->d_A.as<float>(), h->d_R.as<float>(), Lp, x, s)); check_kernel(h, 1);
	} else if (c.use_trans_bias && h->opt_gemm_impl >= 1 && N > 1) {
		XiGemmParams x{};
		x.A = h->d_A.as<float>(); x.lda = Lp; x.R = h->d_R.as<float>(); x.ldb = Lp;
		x.n_frames = N; x.n0 = 1; x.n1 = N; x.k_slab = h->opt_k_slab_tc;
		x.L = L; x.P = P; x.D = D; x.pair_idx = h->d_tidx.as<uint32_t>(); x.Ew = h->d_E.as<float>(); x.e_ld = Lp;
		x.scale = -c.trans_bias_val; x.out = h->d_grad.as<double>();
		CUDA_OK(launch_xi_gemm_tc(x, s)); check_kernel(h, 1);
	} else if (c.use_trans_bias) {
		for (uint32_t d = 1; d <= D; d++) {
			if (d >= N) break;
			ReduceGemmParams r{};
			r.A = h->d_A.as<float>(); r.lda = Lp; r.a_row_shift = d;
			r.B = h->d_R.as<float>() + (size_t)(d - 1) * P; r.ldb = Lp;
			r.n0 = d; r.n1 = N; r.I = L; r.J = P; r.ones_col = 0xffffffffu;
			r.scale = -c.trans_bias_val; r.ones_scale = 0.0; r.mode = 1;
			r.pair_idx = h->d_tidx.as<uint32_t>() + (size_t)(d - 1) * P; r.pair_ld = L;
			r.Ew = h->d_E.as<float>() + (size_t)(d - 1) * P; r.e_ld = Lp;
			r.out = h->d_grad.as<double>(); r.k_slab = h->opt_k_slab;
			launch_reduce_gemm(r, s); check_kernel(h, 1);
		}
	}
	phase_end(h, "xi");
	phase_begin(h, "grad");
	if (tma && lat_tma && (h->opt_tma_mask & 2)) {
		FrameGemmParams r{};
		r.N = N; r.P = P; r.D = D; r.ntile = (P + FRAME_GEMM_TILE - 1) / FRAME_GEMM_TILE; r.k_slab = h->opt_k_slab_tma; r.Mext = nSf + (c.use_state_bias ? 1 : 0);
		r.ones_col = c.use_state_bias ? nSf : 0xffffffffu; r.scale = 1.0; r.ones_scale = c.state_bias_val;
		r.row_idx = h->d_sidx.as<uint32_t>(); r.out = h->d_grad.as<double>();
		r.a_from_tmem = (h->opt_tma_mask & 16) ? 1u : 0u;
		if (virt) {
			r.virt = 1; r.F = c.n_base_ftrs; r.Fp = h->Fp; r.tpb = (5 * (h->Fp / 32) + 3) / 4; r.base2 = h->d_bpad.as<float>(); r.steps = h->d_steps.as<uint32_t>();
			CUDA_OK(launch_state_grad_tma(h->d_Xa.as<float>(), h->Wa, nSf, h->d_Dm.as<float>(), Lp, r, s));
		} else
		CUDA_OK(launch_state_grad_tma(h->X() + c.state_fidx_start, h->Wp, nSf, h->d_Dm.as<float>(), Lp, r, s));
		check_kernel(h, 1);
	} else for (uint32_t d = 0; d < D; d++) {
		ReduceGemmParams r{};
		r.A = h->d_Dm.as<float>() + (size_t)d * P; r.lda = Lp; r.a_row_shift = 0;
		r.B = h->X() + (size_t)d * h->Wp + c.state_fidx_start; r.ldb = h->ldx();
		r.n0 = 0; r.n1 = N; r.I = P; r.J = nSf + (c.use_state_bias ? 1 : 0);
		r.ones_col = c.use_state_bias ? nSf : 0xffffffffu;
		r.scale = 1.0; r.ones_scale = c.state_bias_val; r.mode = 0;
		r.row_idx = h->d_sidx.as<uint32_t>() + (size_t)d * P;
		r.out = h->d_grad.as<double>(); r.k_slab = h->opt_gemm_impl >= 1 ? ks_tc : h->opt_k_slab;
		if (h->opt_gemm_impl >= 1) CUDA_OK(launch_reduce_gemm_tc(r, true, s)); else launch_reduce_gemm(r, s);
		check_kernel(h, 1);
	}
	if (emp_early) CUDA_OK(cudaStreamWaitEvent(s, h->ev_aux_done, 0));      // the empirical counts ran beside the recursions
	else if (!h->nodur_tf) { launch_empirical(make_emp(), s); check_kernel(h, 1); }      // nodur_tf: numerators come from the forward kernel, counts from Dm / Xd
	if (h->opt_mass_check) {
		// the nodes' posterior-mass assertion (frame-level: 0.9..1.1, segmental: the probability that a segment ends here, 0..1)
		h->d_mass.ensure(sizeof(float) * (size_t)N + 16);
		const bool seg = D > 1;
		launch_posterior_mass(h->d_Dm.as<float>(), Lp, L, h->d_node_lab.as<uint32_t>(), N, seg ? -1e-3f : 0.9f, seg ? 1.001f : 1.1f,
		                      h->d_mass.as<float>(), h->d_grad.as<double>() + m.len + 3, s);
		check_kernel(h, 1);
	}
	tail_sums_kernel<<<1, 256, 0, s>>>(h->d_numer.as<double>(), h->d_logZ.as<double>(), h->n_utt, h->d_grad.as<double>() + m.len);
	check_kernel(h, 1);
	phase_end(h, "grad");
	h->fwdbwd_done = true;
}

void viterbi_staged(crfgpu_ctx* h) {
	require_decode(h);
	ensure_windows(h, false);
	const crfgpu_config& c = h->cfg; const Layout& m = h->lay;
	const uint32_t N = h->N, L = m.L, D = c.max_dur, NS = m.n_states, P = m.n_act;
	cudaStream_t s = h->stream;
	ensure_decode_buffers(h);
	if (!N) { h->viterbi_done = true; return; }
	if (h->beam > 0.0 && NS != 1) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "beam pruning is implemented for one state per phone (with N states a pruned node can leave phones without any candidate, and the hypothesis order becomes data-dependent)");
	const bool scored = h->vit_score_ready;      // launched chunk by chunk by crfgpu_stage_batch for this batch and this lambda
	const bool walked = scored && h->vit_rec_ready;   // ... and so were the recursions of each chunk's utterances
	h->vit_score_ready = h->vit_rec_ready = false;
	if (walked) { h->viterbi_done = true; return; }
	if (!scored) phase_begin(h, "viterbi_score");
	VitScoreParams vs{};
	vs.X = h->X(); vs.ldx = h->ldx(); vs.W = h->Wp; vs.sf0 = c.state_fidx_start; vs.nSf = m.nSf; vs.N = N; vs.D = D; vs.L = L;
	vs.frame_t = h->d_frame_t.as<uint32_t>(); vs.Wd = h->d_Wd.as<double>(); vs.use_bias = c.use_state_bias; vs.bias_val = c.state_bias_val;
	vs.negS = h->d_negS.as<float>();
	if (!scored) { launch_vit_scores(vs, s); check_kernel(h, 1); phase_end(h, "viterbi_score"); }
	if (c.use_trans_ftrs) {
		// (float)(-M_n[p][c]) of every frame's decoder pairs from its duration-1 window, in the reference's fp64 arithmetic
		h->d_negMt.ensure(sizeof(float) * (size_t)N * h->vtE + 16);
		VitScoreParams ts{};
		ts.X = h->X(); ts.ldx = h->ldx(); ts.W = h->Wp; ts.sf0 = c.trans_fidx_start; ts.nSf = m.nTf; ts.N = N; ts.D = 1; ts.L = h->vtE;
		ts.frame_t = h->d_frame_t.as<uint32_t>(); ts.Wd = h->d_WdT.as<double>(); ts.use_bias = c.use_trans_bias; ts.bias_val = c.trans_bias_val;
		ts.negS = h->d_negMt.as<float>();
		launch_vit_scores(ts, s); check_kernel(h, 1);
	}
	if (vit_wants_groups(h)) {
		const int gmax = vitg_max_groups(P);
		if (gmax >= 1) {
			phase_begin(h, "viterbi");
			VitGroupParams g{};
			g.n_utt = h->n_utt; g.P = P; g.D = D; g.npt = (P + 31) / 32;
			g.n_batches = (h->n_utt + VITG_UT - 1) / VITG_UT;
			g.n_groups = std::min<uint32_t>((uint32_t)gmax, g.n_batches);
			const uint32_t Pk = (P + 31) / 32 * 32;
			h->d_vg_xch.ensure(sizeof(float) * (size_t)g.n_groups * 2 * Pk * VITG_UT + 16);
			h->d_vg_final.ensure(sizeof(float) * (size_t)h->n_utt * P + 16); h->d_vg_ctr.ensure(sizeof(uint32_t) * g.n_groups + 16);
			g.off = h->d_off.as<uint32_t>(); g.slot_utt = h->d_order16.as<uint32_t>(); g.negS = h->d_negS.as<float>();
			g.crossT = h->d_crossT.as<float>(); g.negDiag = h->d_negDiag.as<float>();
			h->d_vg_cand.ensure(sizeof(float2) * (size_t)h->n_utt * D * P + 16); g.cand = h->d_vg_cand.as<float2>(); g.bp = h->d_bp.as<uint16_t>(); g.bd = h->d_bd.as<uint8_t>();
			g.xch = h->d_vg_xch.as<float>(); g.finalW = h->d_vg_final.as<float>(); g.ctr = h->d_vg_ctr.as<uint32_t>();
			g.out_lab = h->d_olab.as<uint32_t>(); g.out_dur = h->d_odur.as<uint32_t>(); g.out_phn = h->d_ophn.as<uint32_t>();
			g.n_seg = h->d_nseg.as<uint32_t>(); g.cost = h->d_cost.as<float>();
			static DevBuf vdbg; const bool vtiming = getenv("CRFGPU_DP_TIMING") != nullptr;
			if (vtiming) { vdbg.ensure(16 * 8); g.dbg = vdbg.as<unsigned long long>(); }
			CUDA_OK(launch_viterbi_group(g, s)); check_kernel(h, 2);
			if (vtiming) {
				unsigned long long v[16]; CUDA_OK(cudaMemcpyAsync(v, vdbg.p, sizeof(v), cudaMemcpyDeviceToHost, s)); CUDA_OK(cudaStreamSynchronize(s));
				const double n = v[8] ? (double)v[8] : 1.0;
				fprintf(stderr, "[crfgpu] group viterbi: %llu steps; cycles/step: dur>=2 %.0f wait %.0f staging %.0f scan %.0f update+publish %.0f\n",
				        v[8], v[0] / n, v[1] / n, v[2] / n, v[3] / n, v[4] / n);
			}
			phase_end(h, "viterbi");
			h->viterbi_done = true;
			return;
		}
		if (h->opt_vit_impl == 2) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "group-sliced Viterbi: the table slices of this phone count do not fit shared memory");
	}
	phase_begin(h, "viterbi");
	VitParams v = vit_params(h);
	v.n_utt = h->n_utt;
	v.order = h->d_order16.as<uint32_t>();      // longest utterances first
	static DevBuf kdbg; const bool ktiming = getenv("CRFGPU_DP_TIMING") != nullptr;
	if (ktiming) { kdbg.ensure(8 * 8); v.dbg = kdbg.as<unsigned long long>(); }
	launch_viterbi(v, s); check_kernel(h, 1);
	if (ktiming) {
		unsigned long long t[8]; CUDA_OK(cudaMemcpyAsync(t, kdbg.p, sizeof(t), cudaMemcpyDeviceToHost, s)); CUDA_OK(cudaStreamSynchronize(s));
		const double n = t[4] ? (double)t[4] : 1.0;
		fprintf(stderr, "[crfgpu] viterbi (utterance 0, %llu frames); cycles/frame: top %.0f scan %.0f candidates %.0f node %.0f\n", t[4], t[3] / n, t[0] / n, t[1] / n, t[2] / n);
	}
	phase_end(h, "viterbi");
	h->viterbi_done = true;
}

// ---------------------------------------------------------------------------------------------
// NCCL, loaded lazily (dlopen) so that libcrfgpu.so itself loads on a box without NCCL and a process that already carries an NCCL
// (e.g. PyTorch's bundled one) shares it.  Only the lambda-gradient all-reduce of the training seam uses it
// (CRF_Minibatch_GradAccumulator.cpp:277-308 sums the per-stream gradients and scalars serially on the host).
struct NcclApi {
	void* lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi& nccl() {
	static NcclApi api;
	if (api.lib) return api;
	const char* names[] = {getenv("CRFGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	for (const char* n : names) { if (n && *n && (api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
	if (!api.lib) throw ApiError(CRFGPU_ERR_UNSUPPORTED, std::string("NCCL is not loadable (libnccl.so.2; set CRFGPU_NCCL_LIB): ") + (dlerror() ? dlerror() : ""));
	auto sym = [&](const char* n) { void* f = dlsym(api.lib, n); if (!f) { api.lib = nullptr; throw ApiError(CRFGPU_ERR_UNSUPPORTED, std::string("NCCL lacks ") + n); } return f; };
	api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
	api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
	api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
	api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
	api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
	api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
	api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
	api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
	api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
	return api;
}
#define NCCL_OK(expr)                                                                                   \
	do {                                                                                                \
		ncclResult_t r_ = (expr);                                                                       \
		if (r_ != ncclSuccess)                                                                          \
			throw ApiError(CRFGPU_ERR_CUDA, std::string(#expr) + ": " + nccl().GetErrorString(r_));     \
	} while (0)

void comm_release(crfgpu_ctx* h) {
	if (h->comm) { nccl().CommDestroy(h->comm); h->comm = nullptr; h->comm_nranks = 0; h->comm_rank = 0; }
}

// one line of text: which kernels run for the staged batch (crfgpu_plan_info)
std::string plan_text(crfgpu_ctx* h) {
	const crfgpu_config& c = h->cfg;
	char b[512];
	std::string s;
	snprintf(b, sizeof b, "model_type=%u labels=%u lattice_labels=%u max_dur=%u states=%u batch=%u utts/%u frames; ", c.model_type, h->lay.L, h->Lt, c.max_dur, h->lay.n_states, h->n_utt, h->N);
	s += b;
	if (!h->train_ok) s += "train: unsupported (" + h->train_why + "); ";
	else if (!h->have_labels) s += "train: no labels staged; ";
	else if (h->transftr) s += "lattice=transftr_forward/backward_kernel (one CTA per utterance, per-frame transition scores); ";
	else if (h->nodur_tf) s += "lattice=nodur_tf_forward/backward_kernel (one CTA per utterance, per-frame transition scores); ";
	if (h->transftr || h->nodur_tf) s += h->opt_tf_tiled ? "labels^2 GEMMs=score_gemm_tiled/reduce_gemm_tiled_kernel (pre-split bf16 tiles, bulk copies); "
	                                                      : "labels^2 GEMMs=score_gemm_tc/reduce_gemm_tc_kernel (register-staged); ";
	else if (h->nodur) { snprintf(b, sizeof b, "lattice=nodur_dp_kernel (native O(P^2+D*P), %u groups x %u CTAs, 16 utterances in lock-step); ", h->n_nodur_groups, (h->lay.L + 31) / 32); s += b; }
	else if (h->frame_path) s += h->opt_frame_impl == 2 ? "lattice=frame_dp_kernel (one warp per utterance, chains one after the other); " : "lattice=frame_dp_pair_kernel+frame_post_kernel (one warp per chain); ";
	else if (h->ks_ok) { snprintf(b, sizeof b, "lattice=dp_ks_kernel (tcgen05, contraction-sliced: cluster of %u CTAs x %u labels, %u M-tiles in tensor memory, %u clusters x 16 slots%s); ", h->ks_plan.CS, h->ks_plan.CW, h->ks_plan.MT, h->n_ks_clusters, h->tied ? ", tied (duration, phone) expansion" : ""); s += b; }
	else if (h->tc_ok) { snprintf(b, sizeof b, "lattice=dp_tc_kernel (tcgen05, cluster of %u CTAs x %u labels, %u clusters x 16 slots%s); ", h->tc_plan.CS, h->tc_plan.CW, h->n_tc_clusters, h->tied ? ", tied (duration, phone) expansion" : ""); s += b; }
	else if (h->cluster_ok) { snprintf(b, sizeof b, "lattice=cluster_dp_kernel (FFMA fallback: the tcgen05 plan does not hold this geometry or dp_impl=1; cluster of %u CTAs, %u clusters x %d slots); ", h->plan.CS, h->n_clusters, h->plan.UB); s += b; }
	else { snprintf(b, sizeof b, "lattice=forward/backward_kernel (FFMA fallback, E from L2, %d slots per CTA); ", h->U); s += b; }
	snprintf(b, sizeof b, "locksteps=%llu; ", (unsigned long long)h->locksteps); s += b;
	s += h->opt_gemm_impl == 2 ? "gemm=tcgen05 (TMA-fed where the window stream is TMA-addressable, else register-staged); " : h->opt_gemm_impl == 1 ? "gemm=tcgen05 register-staged; " : "gemm=FFMA; ";
	if (!h->decode_ok) s += "decode: unsupported (" + h->decode_why + ")";
	else {
		const uint32_t P = h->lay.n_act;
		const bool vg = h->lay.n_states == 1 && !c.use_trans_ftrs && P >= 2 && h->lay.L == P && (h->opt_vit_impl == 2 || (h->opt_vit_impl == 0 && (size_t)P * P * sizeof(float) > 96 * 1024));
		s += vg ? "decode=viterbi_group_kernel (table sliced over CTA groups)" : "decode=viterbi_kernel (one CTA per utterance)";
	}
	if (h->comm) { snprintf(b, sizeof b, "; comm=nccl rank %d of %d", h->comm_rank, h->comm_nranks); s += b; }
	return s;
}

// what the reference throws for from inside the nodes (runtime_error "Probability sums greater / less than ...", overflow_error from
// CRF_LogMath) surfaces here as CRFGPU_ERR_NUMERIC when the results are read
void tail_numeric(const double* tail4) {
	if (!std::isfinite(tail4[0]) || !std::isfinite(tail4[1]))
		throw ApiError(CRFGPU_ERR_NUMERIC, "non-finite numerator / log partition function in this batch (empty or overflowed log-sum)");
	if (tail4[3] != 0.0)
		throw ApiError(CRFGPU_ERR_NUMERIC, "computeExpF: posterior probability sums outside the reference's band on " + std::to_string((long long)tail4[3]) +
		                                   " frame(s) (CRF_StdStateNode.cpp:252-275 / CRF_StdSegStateNode.cpp:417-436); crfgpu_fetch_posterior_mass shows them");
}
void check_tail_numeric(crfgpu_ctx* h) {
	double t[4];
	CUDA_OK(cudaMemcpyAsync(t, h->d_grad.as<double>() + h->lay.len, sizeof(t), cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(cudaStreamSynchronize(h->stream));
	if (h->N) tail_numeric(t);
}

template <class F>
int guarded(F&& f) {
	try { f(); return CRFGPU_OK; }
	catch (const ApiError& e) { g_err = e.what(); return e.code; }
	catch (const std::exception& e) { g_err = e.what(); return CRFGPU_ERR_ARG; }
}

}  // namespace

extern "C" {

const char* crfgpu_last_error(void) { return g_err.c_str(); }

uint32_t crfgpu_window_width(const crfgpu_config* cfg) { return cfg ? window_width(*cfg) : 0; }

int crfgpu_create(const crfgpu_config* cfg, int device, crfgpu_handle* out) {
	if (out) *out = nullptr;
	crfgpu_ctx* h = nullptr;
	int rc = guarded([&] {
		if (!cfg || !out) throw ApiError(CRFGPU_ERR_ARG, "null argument");
		int n_dev = 0;
		cudaError_t e = cudaGetDeviceCount(&n_dev);
		if (e != cudaSuccess || n_dev == 0)
			throw ApiError(CRFGPU_ERR_CUDA, std::string("no usable CUDA device (libcrfgpu has no CPU fallback): ") +
			                                    (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
		if (device < 0 || device >= n_dev) throw ApiError(CRFGPU_ERR_ARG, "device index out of range");
		CUDA_OK(cudaSetDevice(device));
		h = new crfgpu_ctx();
		h->cfg = *cfg; h->device = device;
		h->lay = build_layout(*cfg);
		h->W = window_width(*cfg);
		h->joined = has_context_or_join(*cfg);
		if (cfg->boundary_delta && cfg->max_dur > 1 && cfg->extract_seg_ftrs) throw ApiError(CRFGPU_ERR_ARG, "extract_segment_features must be false to use boundary_delta_ftrs.");
		if (cfg->n_base_ftrs2 && cfg->boundary_delta2 && cfg->max_dur > 1 && cfg->extract_seg_ftrs2) throw ApiError(CRFGPU_ERR_ARG, "extract_segment_features must be false to use boundary_delta_ftrs.");
		if ((cfg->boundary_delta || (cfg->n_base_ftrs2 && cfg->boundary_delta2)) && cfg->max_dur == 1)
			throw ApiError(CRFGPU_ERR_UNSUPPORTED, "boundary delta features with window length 1 (the reference leaves most of that window vector unwritten)");
		// measured on cfg4 (score + state-gradient GEMM ms per step): unpadded 2.01 + 3.18, padded to 8 floats 1.86 + 3.43, to 32 floats
		// 1.83 + 3.56 -- padding helps the K-major reader and hurts the MN-major one, so the windows stay packed
		h->Wp = cfg->max_dur > 1 ? (h->W + 3) / 4 * 4 : h->W;   // 16-byte window rows: TMA-addressable (max_dur == 1 aliases the base stream)
		if (cfg->max_dur == 0) throw ApiError(CRFGPU_ERR_ARG, "the maximum duration of labels must be larger than 0");
		CUDA_OK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
		setup_label_space(h);
		CUDA_OK(cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
		std::vector<uint32_t> steps = sample_steps(cfg->max_dur);
		upload(h->d_steps, steps, h->stream);
		CUDA_OK(cudaStreamSynchronize(h->stream));
		const char* env = getenv("CRFGPU_SLOTS");
		if (env) h->opt_slots = atoi(env);
		if ((env = getenv("CRFGPU_PREFETCH_SMEM"))) h->opt_prefetch_smem = (uint32_t)atoi(env);
		*out = h;
	});
	if (rc != CRFGPU_OK && h) { delete h; }
	return rc;
}

int crfgpu_destroy(crfgpu_handle h) {
	if (!h) return CRFGPU_OK;
	cudaSetDevice(h->device);
	cudaStreamSynchronize(h->stream);
	try { comm_release(h); } catch (...) {}
	DevBuf* bufs[] = {&h->d_lambda, &h->d_sidx, &h->d_tidx, &h->d_Ws, &h->d_Wt, &h->d_bias, &h->d_E, &h->d_ET, &h->d_steps, &h->d_Wd, &h->d_crossT,
	                  &h->d_negDiag, &h->d_negOff, &h->d_off, &h->d_base, &h->d_frame_t, &h->d_frame_utt, &h->d_frame_len, &h->d_node_lab,
	                  &h->d_prev_lab, &h->d_grp, &h->d_X, &h->d_S, &h->d_A, &h->d_G, &h->d_m, &h->d_kappa, &h->d_bbase, &h->d_Uvec, &h->d_Dm,
	                  &h->d_R, &h->d_logZ, &h->d_numer, &h->d_grad, &h->d_mass, &h->d_negS, &h->d_candW, &h->d_candP, &h->d_bp, &h->d_bd, &h->d_gmove,
	                  &h->d_olab, &h->d_odur, &h->d_ophn, &h->d_nseg, &h->d_cost, &h->d_cl_off, &h->d_cl_list, &h->d_xch, &h->d_xmax, &h->d_smaxd,
	                  &h->d_nd_grp, &h->d_nd_batch, &h->d_nd_xch, &h->d_nd_ctr, &h->d_LB,
	                  &h->d_sidx0, &h->d_tidx0, &h->d_tmax, &h->d_lam_acc, &h->d_lam_sqr_acc, &h->d_grad_sqr_acc, &h->d_base2, &h->d_X2, &h->d_frame_t2, &h->d_bpad, &h->d_Xa, &h->d_bias_dy, &h->d_baseB, &h->d_lm_start, &h->d_lm_bigT, &h->d_lm_final, &h->d_lm_exit, &h->d_frame_utt2, &h->d_frame_len2, &h->d_node_lab2, &h->d_prev_lab2, &h->d_next_lab2, &h->d_bpad2, &h->d_Xa2, &h->d_Wtr, &h->d_tbias, &h->d_Mall, &h->d_Xd, &h->d_next_lab, &h->d_WdT, &h->d_vt_base, &h->d_negMt,
	                  &h->d_Eall, &h->d_rowmax, &h->d_XdT, &h->d_XtT, &h->d_XtK, &h->d_WtrT, &h->d_baseB2, &h->d_XtK2, &h->d_XtT2, &h->d_order16, &h->d_vg_xch, &h->d_vg_final, &h->d_vg_ctr, &h->d_vg_cand, &h->d_vorder, &h->d_off2};
	for (DevBuf* b : bufs) b->release();
	for (auto& kv : h->phases) { cudaEventDestroy(kv.second.first); cudaEventDestroy(kv.second.second); }
	for (cudaEvent_t e : h->ev_chunk) cudaEventDestroy(e);
	for (cudaEvent_t e : h->ev_chunk2) cudaEventDestroy(e);
	if (h->ev_pre_done) cudaEventDestroy(h->ev_pre_done);
	if (h->ev_swap) cudaEventDestroy(h->ev_swap);
	if (h->ev_pre_ready) cudaEventDestroy(h->ev_pre_ready);
	if (h->pre_stream) { cudaStreamSynchronize(h->pre_stream); cudaStreamDestroy(h->pre_stream); }
	if (h->ev_ready) cudaEventDestroy(h->ev_ready);
	if (h->ev_pin) cudaEventDestroy(h->ev_pin);
	if (h->pin) cudaFreeHost(h->pin);
	if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
	if (h->aux_stream) { cudaStreamSynchronize(h->aux_stream); cudaStreamDestroy(h->aux_stream); cudaEventDestroy(h->ev_aux_go); cudaEventDestroy(h->ev_aux_done); }
	for (cudaStream_t st : h->rec_stream) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
	for (cudaEvent_t e : h->ev_scored) cudaEventDestroy(e);
	for (cudaEvent_t e : h->ev_walked) cudaEventDestroy(e);
	cudaStreamDestroy(h->stream);
	delete h;
	return CRFGPU_OK;
}

uint32_t crfgpu_lambda_len(crfgpu_handle h) { return h ? h->lay.len : 0; }

int crfgpu_index_maps(crfgpu_handle h, uint32_t* state_idx, uint32_t* trans_idx) {
	return guarded([&] {
		if (!h || !state_idx || !trans_idx) throw ApiError(CRFGPU_ERR_ARG, "null argument");
		std::memcpy(state_idx, h->lay.sidx.data(), sizeof(uint32_t) * h->lay.sidx.size());
		std::memcpy(trans_idx, h->lay.tidx.data(), sizeof(uint32_t) * h->lay.tidx.size());
	});
}

int crfgpu_set_lambda(crfgpu_handle h, const double* lambda, uint32_t len) {
	return guarded([&] {
		if (!h || !lambda) throw ApiError(CRFGPU_ERR_ARG, "null argument");
		CUDA_OK(cudaSetDevice(h->device));
		set_lambda(h, lambda, len);
	});
}

int crfgpu_sgd_update(crfgpu_handle h, const crfgpu_sgd* opt, double n_active) {
	return guarded([&] {
		if (!h || !opt) throw ApiError(CRFGPU_ERR_ARG, "null argument");
		if (!h->fwdbwd_done || !h->have_lambda) throw ApiError(CRFGPU_ERR_ARG, "crfgpu_sgd_update needs a staged gradient (crfgpu_fwdbwd_staged / crfgpu_fwdbwd_batch)");
		if (!(n_active > 0.0)) throw ApiError(CRFGPU_ERR_ARG, "n_active must be positive");
		CUDA_OK(cudaSetDevice(h->device));
		const size_t len = h->lay.len, bytes = sizeof(double) * len;
		if (!h->have_train_state) {
			h->d_lam_acc.ensure(bytes + 16); h->d_lam_sqr_acc.ensure(bytes + 16); h->d_grad_sqr_acc.ensure(bytes + 16);
			CUDA_OK(cudaMemsetAsync(h->d_lam_acc.p, 0, bytes, h->stream)); CUDA_OK(cudaMemsetAsync(h->d_lam_sqr_acc.p, 0, bytes, h->stream));
			CUDA_OK(cudaMemsetAsync(h->d_grad_sqr_acc.p, 0, bytes, h->stream));
			h->have_train_state = true;
		}
		SgdParams p{};
		p.lambda = h->d_lambda.as<double>(); p.grad = h->d_grad.as<double>(); p.len = len; p.n_active = n_active; p.lr = opt->lr;
		p.use_gvar = opt->use_gvar; p.inv_square_var = opt->inv_square_var; p.use_adagrad = opt->use_adagrad; p.eta = opt->eta; p.eps = opt->eps;
		p.grad_sqr_acc = h->d_grad_sqr_acc.as<double>(); p.lambda_acc = h->d_lam_acc.as<double>(); p.lambda_sqr_acc = h->d_lam_sqr_acc.as<double>();
		CUDA_OK(launch_sgd_update(p, h->stream)); check_kernel(h, 1);
		h->fwdbwd_done = false;                   // the gradient has been consumed
		derive_tables(h);
	});
}

int crfgpu_get_lambda(crfgpu_handle h, double* lambda, double* lambda_acc, double* lambda_sqr_acc, double* grad_sqr_acc) {
	return guarded([&] {
		if (!h || !h->have_lambda) throw ApiError(CRFGPU_ERR_ARG, "no lambda on the device");
		CUDA_OK(cudaSetDevice(h->device));
		const size_t bytes = sizeof(double) * (size_t)h->lay.len;
		if ((lambda_acc || lambda_sqr_acc || grad_sqr_acc) && !h->have_train_state) throw ApiError(CRFGPU_ERR_ARG, "no training state: crfgpu_sgd_update has not run");
		if (lambda) CUDA_OK(cudaMemcpyAsync(lambda, h->d_lambda.p, bytes, cudaMemcpyDeviceToHost, h->stream));
		if (lambda_acc) CUDA_OK(cudaMemcpyAsync(lambda_acc, h->d_lam_acc.p, bytes, cudaMemcpyDeviceToHost, h->stream));
		if (lambda_sqr_acc) CUDA_OK(cudaMemcpyAsync(lambda_sqr_acc, h->d_lam_sqr_acc.p, bytes, cudaMemcpyDeviceToHost, h->stream));
		if (grad_sqr_acc) CUDA_OK(cudaMemcpyAsync(grad_sqr_acc, h->d_grad_sqr_acc.p, bytes, cudaMemcpyDeviceToHost, h->stream));
		CUDA_OK(cudaStreamSynchronize(h->stream));
	});
}

int crfgpu_set_train_state(crfgpu_handle h, const double* lambda_acc, const double* lambda_sqr_acc, const double* grad_sqr_acc) {
	return guarded([&] {
		if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle");
		CUDA_OK(cudaSetDevice(h->device));
		const size_t bytes = sizeof(double) * (size_t)h->lay.len;
		h->d_lam_acc.ensure(bytes + 16); h->d_lam_sqr_acc.ensure(bytes + 16); h->d_grad_sqr_acc.ensure(bytes + 16);
		auto put = [&](DevBuf& b, const double* src) {
			if (src) CUDA_OK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
			else CUDA_OK(cudaMemsetAsync(b.p, 0, bytes, h->stream));
		};
		put(h->d_lam_acc, lambda_acc); put(h->d_lam_sqr_acc, lambda_sqr_acc); put(h->d_grad_sqr_acc, grad_sqr_acc);
		CUDA_OK(cudaStreamSynchronize(h->stream));
		h->have_train_state = true;
	});
}

int crfgpu_prefetch_batch(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs) {
	return guarded([&] {
		if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle");
		CUDA_OK(cudaSetDevice(h->device));
		prefetch_batch(h, n_utt, frame_off, base_ftrs);
	});
}

int crfgpu_prefetch_train_batch(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const uint32_t* frame_labs) {
	return guarded([&] {
		if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle");
		CUDA_OK(cudaSetDevice(h->device));
		prefetch_batch(h, n_utt, frame_off, base_ftrs, frame_labs);
	});
}

int crfgpu_prefetch_train_batch2(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2, const uint32_t* frame_labs) {
	return guarded([&] {
		if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle");
		CUDA_OK(cudaSetDevice(h->device));
		prefetch_batch(h, n_utt, frame_off, base_ftrs, frame_labs, base_ftrs2);
	});
}

int crfgpu_stage_batch(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const uint32_t* frame_labs) {
	return guarded([&] {
		if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle");
		CUDA_OK(cudaSetDevice(h->device));
		stage_batch(h, n_utt, frame_off, base_ftrs, frame_labs);
	});
}

int crfgpu_stage_batch2(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2, const uint32_t* frame_labs) {
	return guarded([&] {
		if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle");
		CUDA_OK(cudaSetDevice(h->device));
		stage_batch(h, n_utt, frame_off, base_ftrs, frame_labs, base_ftrs2);
	});
}

int crfgpu_fwdbwd_staged(crfgpu_handle h) {
	return guarded([&] { if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle"); CUDA_OK(cudaSetDevice(h->device)); fwdbwd_staged(h); });
}
int crfgpu_viterbi_staged(crfgpu_handle h) {
	return guarded([&] { if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle"); CUDA_OK(cudaSetDevice(h->device)); viterbi_staged(h); });
}

int crfgpu_device_results(crfgpu_handle h, double** d_grad, double** d_numer, double** d_logZ) {
	return guarded([&] {
		if (!h || !h->fwdbwd_done) throw ApiError(CRFGPU_ERR_ARG, "no forward-backward results staged");
		if (d_grad) *d_grad = h->d_grad.as<double>();
		if (d_numer) *d_numer = h->d_numer.as<double>();
		if (d_logZ) *d_logZ = h->d_logZ.as<double>();
	});
}

int crfgpu_fetch_fwdbwd(crfgpu_handle h, double* grad, double* numer, double* logZ) {
	return guarded([&] {
		if (!h || !h->fwdbwd_done) throw ApiError(CRFGPU_ERR_ARG, "no forward-backward results staged");
		CUDA_OK(cudaSetDevice(h->device));
		if (grad) CUDA_OK(cudaMemcpyAsync(grad, h->d_grad.p, sizeof(double) * (size_t)h->lay.len, cudaMemcpyDeviceToHost, h->stream));
		if (numer && h->n_utt) CUDA_OK(cudaMemcpyAsync(numer, h->d_numer.p, sizeof(double) * (size_t)h->n_utt, cudaMemcpyDeviceToHost, h->stream));
		if (logZ && h->n_utt) CUDA_OK(cudaMemcpyAsync(logZ, h->d_logZ.p, sizeof(double) * (size_t)h->n_utt, cudaMemcpyDeviceToHost, h->stream));
		CUDA_OK(cudaStreamSynchronize(h->stream));
		// the lattice kernels write NaN/Inf only when a log-sum was empty or overflowed -- the cases in which the
		// reference throws from logE/expE (CRF_LogMath.cpp:192-224)
		if (logZ) for (uint32_t u = 0; u < h->n_utt; u++)
			if (!std::isfinite(logZ[u])) throw ApiError(CRFGPU_ERR_NUMERIC, "non-finite log partition function for utterance " + std::to_string(u));
		check_tail_numeric(h);
	});
}

int crfgpu_fetch_viterbi(crfgpu_handle h, uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg, float* path_cost) {
	return guarded([&] {
		if (!h || !h->viterbi_done) throw ApiError(CRFGPU_ERR_ARG, "no Viterbi results staged");
		CUDA_OK(cudaSetDevice(h->device));
		const size_t nb = sizeof(uint32_t) * (size_t)h->N;
		if (out_lab && h->N) CUDA_OK(cudaMemcpyAsync(out_lab, h->d_olab.p, nb, cudaMemcpyDeviceToHost, h->stream));
		if (out_dur && h->N) CUDA_OK(cudaMemcpyAsync(out_dur, h->d_odur.p, nb, cudaMemcpyDeviceToHost, h->stream));
		if (out_phn && h->N) CUDA_OK(cudaMemcpyAsync(out_phn, h->d_ophn.p, nb, cudaMemcpyDeviceToHost, h->stream));
		if (n_seg && h->n_utt) CUDA_OK(cudaMemcpyAsync(n_seg, h->d_nseg.p, sizeof(uint32_t) * (size_t)h->n_utt, cudaMemcpyDeviceToHost, h->stream));
		if (path_cost && h->n_utt) CUDA_OK(cudaMemcpyAsync(path_cost, h->d_cost.p, sizeof(float) * (size_t)h->n_utt, cudaMemcpyDeviceToHost, h->stream));
		CUDA_OK(cudaStreamSynchronize(h->stream));
	});
}

int crfgpu_fwdbwd_batch(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const uint32_t* frame_labs,
                        double* grad, double* numer, double* logZ) {
	return crfgpu_fwdbwd_batch2(h, n_utt, frame_off, base_ftrs, nullptr, frame_labs, grad, numer, logZ);
}

int crfgpu_fwdbwd_batch2(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                         const uint32_t* frame_labs, double* grad, double* numer, double* logZ) {
	const bool verbose = getenv("CRFGPU_VERBOSE") != nullptr;
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
	const auto t0 = now();
	std::chrono::steady_clock::time_point t1, t2;
	int rc = guarded([&] {
		if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle");
		CUDA_OK(cudaSetDevice(h->device));
		require_train(h);
		if (!frame_labs) throw ApiError(CRFGPU_ERR_ARG, "training needs frame labels");
		stage_batch(h, n_utt, frame_off, base_ftrs, frame_labs, base_ftrs2);
		t1 = now();
		fwdbwd_staged(h);
		t2 = now();
	});
	if (rc != CRFGPU_OK) return rc;
	rc = crfgpu_fetch_fwdbwd(h, grad, numer, logZ);
	if (verbose && h->ev_ready) {
		fprintf(stderr, "[crfgpu] device timeline from the start of staging (ms):");
		float t = 0.0f;
		for (size_t k = 0; k < std::min(h->chunk_end.size(), h->ev_chunk.size()); k++) if (cudaEventElapsedTime(&t, h->ev_ready, h->ev_chunk[k]) == cudaSuccess) fprintf(stderr, " H2D chunk %zu done %.3f", k, t);
		for (const char* ph : {"expand", "score", "grad"}) {
			auto it = h->phases.find(ph);
			if (it != h->phases.end() && cudaEventElapsedTime(&t, h->ev_ready, it->second.first) == cudaSuccess) {
				float t2 = 0.0f; cudaEventElapsedTime(&t2, h->ev_ready, it->second.second);
				fprintf(stderr, " | %s %.3f..%.3f", ph, t, t2);
			}
		}
		fprintf(stderr, "\n");
	}
	if (verbose) fprintf(stderr, "[crfgpu] fwdbwd_batch host timeline: stage (prep + H2D enqueue + sync) %.3f ms, launch %.3f ms, kernels + D2H %.3f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, now()));
	return rc;
}

int crfgpu_viterbi_batch(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs,
                         uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg, float* path_cost) {
	return crfgpu_viterbi_batch2(h, n_utt, frame_off, base_ftrs, nullptr, out_lab, out_dur, out_phn, n_seg, path_cost);
}

int crfgpu_viterbi_batch2(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                          uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg, float* path_cost) {
	const bool verbose = getenv("CRFGPU_VERBOSE") != nullptr;
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
	const auto t0 = now();
	std::chrono::steady_clock::time_point t1, t2;
	int rc = guarded([&] {
		if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle");
		CUDA_OK(cudaSetDevice(h->device));
		require_decode(h);
		stage_batch(h, n_utt, frame_off, base_ftrs, nullptr, base_ftrs2);
		t1 = now();
		viterbi_staged(h);
		t2 = now();
	});
	if (rc != CRFGPU_OK) return rc;
	rc = crfgpu_fetch_viterbi(h, out_lab, out_dur, out_phn, n_seg, path_cost);
	if (verbose && h->ev_ready) {
		fprintf(stderr, "[crfgpu] device timeline from the start of staging (ms):");
		float t = 0.0f;
		for (size_t k = 0; k < std::min(h->chunk_end.size(), h->ev_chunk.size()); k++) if (cudaEventElapsedTime(&t, h->ev_ready, h->ev_chunk[k]) == cudaSuccess) fprintf(stderr, " H2D chunk %zu done %.3f", k, t);
		for (const char* ph : {"expand", "viterbi_score", "viterbi"}) {
			auto it = h->phases.find(ph);
			if (it != h->phases.end() && cudaEventElapsedTime(&t, h->ev_ready, it->second.first) == cudaSuccess) {
				float t2 = 0.0f; cudaEventElapsedTime(&t2, h->ev_ready, it->second.second);
				fprintf(stderr, " | %s %.3f..%.3f", ph, t, t2);
			}
		}
		fprintf(stderr, "\n");
	}
	if (verbose) fprintf(stderr, "[crfgpu] viterbi_batch host timeline: stage (prep + H2D enqueue) %.3f ms, launch %.3f ms, kernels + D2H %.3f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, now()));
	return rc;
}

int crfgpu_set_phone_lm(crfgpu_handle h, const float* lm_start, const float* lm_bigram, const float* lm_final) {
	return guarded([&] {
		if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle");
		CUDA_OK(cudaSetDevice(h->device));
		if (!lm_start && !lm_bigram && !lm_final) { h->have_lm = false; h->vit_rec_ready = false; return; }
		if (!lm_start || !lm_bigram || !lm_final) throw ApiError(CRFGPU_ERR_ARG, "the phone-bigram LM needs all three weight arrays (or none, to drop it)");
		if (!h->decode_ok) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "decoding: " + h->decode_why);
		if (h->lay.n_states != 1) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "the phone-bigram LM is implemented for one state per phone (the N-state free-phone LM of the reference returns to its start state through epsilon arcs)");
		const uint32_t P = h->lay.n_act;
		std::vector<float> st(lm_start, lm_start + P), fin(lm_final, lm_final + P), bt((size_t)P * P);
		for (uint32_t a = 0; a < P; a++) for (uint32_t b = 0; b < P; b++) {
			const float w = lm_bigram[(size_t)a * P + b];
			if (a != b && !std::isfinite(w)) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "the phone-bigram LM must be complete (a finite weight on every arc p -> q, p != q): a missing arc changes the hypothesis order the decoder's tie-breaks depend on");
			bt[(size_t)b * P + a] = w;                      // transposed: [to][from]
		}
		for (uint32_t a = 0; a < P; a++) if (!std::isfinite(st[a])) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "the phone-bigram LM must have a finite start weight for every phone");
		upload(h->d_lm_start, st, h->stream); upload(h->d_lm_bigT, bt, h->stream); upload(h->d_lm_final, fin, h->stream);
		CUDA_OK(cudaStreamSynchronize(h->stream));
		h->have_lm = true; h->viterbi_done = false; h->vit_rec_ready = false;
	});
}

int crfgpu_set_beam(crfgpu_handle h, double beam) {
	return guarded([&] {
		if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle");
		if (!(beam >= 0.0) || !std::isfinite(beam)) throw ApiError(CRFGPU_ERR_ARG, "the beam must be a finite number >= 0 (0 = no pruning)");
		if (beam > 0.0 && h->lay.n_states != 1) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "beam pruning is implemented for one state per phone");
		h->beam = beam; h->viterbi_done = false; h->vit_rec_ready = false;
	});
}

int crfgpu_set_phone_unigram_lm(crfgpu_handle h, const float* lm_unigram, const float* lm_exit, const float* lm_final) {
	return guarded([&] {
		if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle");
		CUDA_OK(cudaSetDevice(h->device));
		if (!lm_unigram && !lm_exit && !lm_final) { h->have_lm = false; h->vit_rec_ready = false; return; }
		if (!lm_unigram || !lm_exit || !lm_final) throw ApiError(CRFGPU_ERR_ARG, "the phone-unigram LM needs all three weight arrays (or none, to drop it)");
		if (!h->decode_ok) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "decoding: " + h->decode_why);
		if (h->lay.n_states < 2) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "the unigram + exit-cost LM has the topology of the reference's N-state free-phone LM (epsilon arcs back to the start state); with one state per phone use crfgpu_set_phone_lm");
		const uint32_t P = h->lay.n_act;
		std::vector<float> un(lm_unigram, lm_unigram + P), ex(lm_exit, lm_exit + P), fin(lm_final, lm_final + P);
		for (uint32_t a = 0; a < P; a++) if (!std::isfinite(un[a]) || !std::isfinite(ex[a])) throw ApiError(CRFGPU_ERR_UNSUPPORTED, "the phone-unigram LM must have a finite unigram and exit cost for every phone");
		upload(h->d_lm_start, un, h->stream); upload(h->d_lm_exit, ex, h->stream); upload(h->d_lm_final, fin, h->stream);
		CUDA_OK(cudaStreamSynchronize(h->stream));
		h->have_lm = true; h->viterbi_done = false; h->vit_rec_ready = false;
	});
}

int crfgpu_expand_windows(crfgpu_handle h, uint32_t n_frames, const float* base_ftrs, float* out) {
	return crfgpu_expand_windows2(h, n_frames, base_ftrs, nullptr, out);
}

int crfgpu_expand_windows2(crfgpu_handle h, uint32_t n_frames, const float* base_ftrs, const float* base_ftrs2, float* out) {
	return guarded([&] {
		if (!h || !base_ftrs || !out) throw ApiError(CRFGPU_ERR_ARG, "null argument");
		CUDA_OK(cudaSetDevice(h->device));
		const uint32_t off[2] = {0, n_frames};
		if (!n_frames) return;
		h->full_windows = true;                       // the caller reads every column of every window back
		try { stage_batch(h, 1, off, base_ftrs, nullptr, base_ftrs2); } catch (...) { h->full_windows = false; throw; }
		h->full_windows = false;
		CUDA_OK(cudaMemcpy2DAsync(out, sizeof(float) * h->W, h->X(), sizeof(float) * h->Wp, sizeof(float) * h->W,
		                          (size_t)n_frames * h->cfg.max_dur, cudaMemcpyDeviceToHost, h->stream));
		CUDA_OK(cudaStreamSynchronize(h->stream));
	});
}

int crfgpu_group_labels(const crfgpu_config* cfg, uint32_t n_frames, const uint32_t* frame_labs, uint32_t* out4) {
	return guarded([&] {
		if (!cfg || !frame_labs || !out4) throw ApiError(CRFGPU_ERR_ARG, "null argument");
		if (cfg->max_dur == 0) throw ApiError(CRFGPU_ERR_ARG, "the maximum duration of labels must be larger than 0");
		group_labels(cfg->max_dur, n_frames, frame_labs, out4);
	});
}

int crfgpu_synchronize(crfgpu_handle h) {
	return guarded([&] { if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle"); CUDA_OK(cudaSetDevice(h->device)); CUDA_OK(cudaStreamSynchronize(h->stream)); });
}
void* crfgpu_stream(crfgpu_handle h) { return h ? (void*)h->stream : nullptr; }
uint64_t crfgpu_launch_count(crfgpu_handle h) { return h ? h->launches : 0; }

double crfgpu_phase_ms(crfgpu_handle h, const char* phase) {
	if (!h || !phase) return -1.0;
	auto it = h->phases.find(phase);
	if (it == h->phases.end()) return -1.0;
	float ms = -1.0f;
	if (cudaEventSynchronize(it->second.second) != cudaSuccess) return -1.0;
	if (cudaEventElapsedTime(&ms, it->second.first, it->second.second) != cudaSuccess) return -1.0;
	return ms;
}

int crfgpu_fetch_posterior_mass(crfgpu_handle h, float* mass) {
	return guarded([&] {
		if (!h || !h->fwdbwd_done || !mass) throw ApiError(CRFGPU_ERR_ARG, "no forward-backward results staged");
		if (!h->opt_mass_check || h->transftr) throw ApiError(CRFGPU_ERR_ARG, "the posterior-mass pass did not run for this batch");
		CUDA_OK(cudaSetDevice(h->device));
		if (h->N) CUDA_OK(cudaMemcpyAsync(mass, h->d_mass.p, sizeof(float) * (size_t)h->N, cudaMemcpyDeviceToHost, h->stream));
		CUDA_OK(cudaStreamSynchronize(h->stream));
	});
}

int crfgpu_fetch_alpha_beta(crfgpu_handle h, double* alpha, double* beta) {
	return guarded([&] {
		if (!h || !h->fwdbwd_done || !alpha || !beta) throw ApiError(CRFGPU_ERR_ARG, "no forward-backward results staged");
		if (!h->opt_keep_lattice) throw ApiError(CRFGPU_ERR_ARG, "set option keep_lattice=1 before running forward-backward");
		CUDA_OK(cudaSetDevice(h->device));
		const size_t n = (size_t)h->N * h->Lt;
		DevBuf da, db; da.ensure(sizeof(double) * n + 16); db.ensure(sizeof(double) * n + 16);
		DpParams p = dp_params(h);
		launch_dump_alpha_beta(p, h->N, h->d_frame_t.as<uint32_t>(), h->d_frame_len.as<uint32_t>(), da.as<double>(), db.as<double>(), h->stream);
		check_kernel(h, 1);
		CUDA_OK(cudaMemcpyAsync(alpha, da.p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
		CUDA_OK(cudaMemcpyAsync(beta, db.p, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
		CUDA_OK(cudaStreamSynchronize(h->stream));
		da.release(); db.release();
	});
}

int crfgpu_set_option(crfgpu_handle h, const char* name, int64_t value) {
	return guarded([&] {
		if (!h || !name) throw ApiError(CRFGPU_ERR_ARG, "null argument");
		const std::string n(name);
		if (n == "slots") h->opt_slots = (int)value;
		else if (n == "k_slab") { if (value < 16) throw ApiError(CRFGPU_ERR_ARG, "k_slab must be >= 16"); h->opt_k_slab = (uint32_t)value; }
		else if (n == "keep_lattice") h->opt_keep_lattice = value != 0;
		else if (n == "virt_windows") h->opt_virt = value != 0;           // 0: the training GEMMs read fully materialised windows (34 kB per cfg4 frame instead of 12.8)
		else if (n == "gemm_impl") h->opt_gemm_impl = (int)value;        // 0: fp32 FFMA tiles; 1: tcgen05 split-bf16, register-staged; 2: + TMA-fed window GEMMs
		else if (n == "nodur_impl") {                                    // stdseg_no_dur* lattice: 0 auto, 1 native O(P^2 + D*P), 2 tied (duration, phone) expansion
			h->opt_nodur_impl = (int)value;
			setup_label_space(h);                                          // new label space: crfgpu_set_lambda must be called again
		}
		else if (n == "frame_impl") h->opt_frame_impl = (int)value;      // frame-level models with <= 64 labels: 0 one warp per chain, both chains in one launch (crf_dp_frame.cu), 2 the chains one after the other, 1 the cluster lattice kernels
		else if (n == "tf_tiled") { h->opt_tf_tiled = value != 0 ? 1 : 0; h->have_lambda = false; h->x_tiles_valid = false; }      // transition-feature GEMMs: 1 pre-tiled operands + bulk copies, 0 register-staged kernels (set_lambda again: the weight tiles belong to the tiled path)
		else if (n == "aux_empirical") h->opt_aux_empirical = value != 0 ? 1 : 0;      // empirical counts on a side stream beside the recursions (1) or behind the state gradient (0)
		else if (n == "vit_eager") h->opt_vit_eager = value != 0.0 ? 1 : 0;
		else if (n == "vit_impl") { h->opt_vit_impl = (int)value; h->vit_rec_ready = false; }         // Viterbi recursion: 0 auto, 1 one CTA per utterance, 2 table sliced over groups of CTAs (one state per phone)
		else if (n == "tma_mask") h->opt_tma_mask = (int)value;          // debug: 1 score, 2 state gradient, 4 Xi through the TMA-fed kernels, 8 / 16 / 32 the 128-row operand of the score / state-gradient / Xi GEMM through tensor memory
		else if (n == "prefetch_smem") h->opt_prefetch_smem = (uint32_t)value;   // shared-memory cap of the read-ahead expansion's CTAs
		else if (n == "k_slab_xi") { if (value < 32 || value % 32) throw ApiError(CRFGPU_ERR_ARG, "k_slab_xi must be a multiple of 32"); h->opt_k_slab_xi = (uint32_t)value; }
		else if (n == "k_slab_tma") { if (value < 32 || value % 32) throw ApiError(CRFGPU_ERR_ARG, "k_slab_tma must be a multiple of 32"); h->opt_k_slab_tma = (uint32_t)value; }
		else if (n == "k_slab_tc") { if (value < 32) throw ApiError(CRFGPU_ERR_ARG, "k_slab_tc must be >= 32"); h->opt_k_slab_tc = (uint32_t)value; }
		else if (n == "dp_impl") h->opt_dp_impl = (int)value;            // 0: one CTA per utterance group, E from L2; 1: cluster-resident E, FFMA; 2: cluster-resident E, tcgen05, output labels sliced; 3 (default): tcgen05, contraction sliced, E whole in tensor memory
		else if (n == "cluster_slots") h->opt_cluster_slots = (int)value; // utterance slots per cluster (4,8,12,16,32; 0 auto)
		else if (n == "mass_check") h->opt_mass_check = value != 0;       // 0 skips the posterior-mass assertion pass (one read of the posterior array)
		else if (n == "max_clusters") h->opt_max_clusters = (int)value;   // cap on the resident clusters of the lattice kernels (0 = all): a small cap makes every slot work through a long utterance list
		else throw ApiError(CRFGPU_ERR_ARG, "unknown option " + n);
		if ((n == "virt_windows" || n == "gemm_impl" || n == "tma_mask") && decide_virt(h) && h->have_lambda) {
			CUDA_OK(cudaSetDevice(h->device));
			derive_tables(h);                                              // the weight tiles follow the chunk order of the window form
		}
	});
}

// ---- multi-GPU: the one exchange step of the training path --------------------------------------------------------------------
int crfgpu_comm_unique_id(void* id128) {
	return guarded([&] {
		if (!id128) throw ApiError(CRFGPU_ERR_ARG, "null argument");
		static_assert(sizeof(ncclUniqueId) == CRFGPU_COMM_ID_BYTES, "id size");
		ncclUniqueId id;
		NCCL_OK(nccl().GetUniqueId(&id));
		std::memcpy(id128, &id, sizeof(id));
	});
}

int crfgpu_comm_init_rank(crfgpu_handle h, int n_ranks, int rank, const void* id128) {
	return guarded([&] {
		if (!h || !id128) throw ApiError(CRFGPU_ERR_ARG, "null argument");
		if (n_ranks < 1 || rank < 0 || rank >= n_ranks) throw ApiError(CRFGPU_ERR_ARG, "rank out of range");
		CUDA_OK(cudaSetDevice(h->device));
		comm_release(h);
		ncclUniqueId id;
		std::memcpy(&id, id128, sizeof(id));
		NCCL_OK(nccl().CommInitRank(&h->comm, n_ranks, id, rank));
		h->comm_nranks = n_ranks; h->comm_rank = rank;
	});
}

int crfgpu_comm_init_all(crfgpu_handle* handles, int n) {
	return guarded([&] {
		if (!handles || n < 1) throw ApiError(CRFGPU_ERR_ARG, "null argument");
		std::vector<int> devs(n);
		for (int i = 0; i < n; i++) {
			if (!handles[i]) throw ApiError(CRFGPU_ERR_ARG, "null handle");
			devs[i] = handles[i]->device;
			for (int j = 0; j < i; j++) if (devs[j] == devs[i]) throw ApiError(CRFGPU_ERR_ARG, "crfgpu_comm_init_all needs one handle per DEVICE (two handles share device " + std::to_string(devs[i]) + ")");
			if (handles[i]->lay.len != handles[0]->lay.len) throw ApiError(CRFGPU_ERR_ARG, "the handles of one communicator must share the model geometry");
			comm_release(handles[i]);
		}
		std::vector<ncclComm_t> comms(n);
		NCCL_OK(nccl().CommInitAll(comms.data(), n, devs.data()));
		for (int i = 0; i < n; i++) { handles[i]->comm = comms[i]; handles[i]->comm_nranks = n; handles[i]->comm_rank = i; }
	});
}

int crfgpu_comm_destroy(crfgpu_handle h) {
	return guarded([&] { if (!h) throw ApiError(CRFGPU_ERR_ARG, "null handle"); CUDA_OK(cudaSetDevice(h->device)); comm_release(h); });
}

int crfgpu_comm_size(crfgpu_handle h) { return h ? (h->comm ? h->comm_nranks : 0) : 0; }

int crfgpu_group_start(void) { return guarded([&] { NCCL_OK(nccl().GroupStart()); }); }
int crfgpu_group_end(void) { return guarded([&] { NCCL_OK(nccl().GroupEnd()); }); }

int crfgpu_allreduce_grad(crfgpu_handle h) {
	return guarded([&] {
		if (!h || !h->fwdbwd_done) throw ApiError(CRFGPU_ERR_ARG, "crfgpu_allreduce_grad needs a staged gradient (crfgpu_fwdbwd_staged)");
		if (!h->comm) throw ApiError(CRFGPU_ERR_ARG, "no communicator: call crfgpu_comm_init_rank / crfgpu_comm_init_all first");
		CUDA_OK(cudaSetDevice(h->device));
		// gradient + [sum numer, sum logZ, n_utt, 0] in ONE collective, in place, ordered on the handle's stream behind the gradient kernels
		NCCL_OK(nccl().AllReduce(h->d_grad.p, h->d_grad.p, (size_t)h->lay.len + 4, ncclDouble, ncclSum, h->comm, h->stream));
		h->launches += 1;
	});
}

int crfgpu_fetch_tail(crfgpu_handle h, double* tail4) {
	return guarded([&] {
		if (!h || !h->fwdbwd_done || !tail4) throw ApiError(CRFGPU_ERR_ARG, "no forward-backward results staged");
		CUDA_OK(cudaSetDevice(h->device));
		CUDA_OK(cudaMemcpyAsync(tail4, h->d_grad.as<double>() + h->lay.len, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
		CUDA_OK(cudaStreamSynchronize(h->stream));
		tail_numeric(tail4);
	});
}

int crfgpu_shard_views(uint32_t n_utt, uint32_t n_streams, uint32_t* first, uint32_t* count) {
	return guarded([&] {
		if (!first || !count || !n_streams) throw ApiError(CRFGPU_ERR_ARG, "null argument / zero streams");
		shard_views(n_utt, n_streams, first, count);
	});
}
uint32_t crfgpu_minibatch_share(uint32_t minibatch, uint32_t n_streams, uint32_t stream) { return n_streams ? minibatch_share(minibatch, n_streams, stream) : 0; }
int crfgpu_balance_utts(uint32_t n_utt, const uint32_t* n_frames, uint32_t n_ranks, uint32_t* rank_of) {
	return guarded([&] {
		if ((n_utt && (!n_frames || !rank_of)) || !n_ranks) throw ApiError(CRFGPU_ERR_ARG, "null argument / zero ranks");
		balance_utts(n_utt, n_frames, n_ranks, rank_of);
	});
}

int crfgpu_balance_utts_cost(uint32_t n_utt, const uint32_t* n_frames, uint32_t n_ranks, uint32_t n_slots, double step_frames, uint32_t* rank_of) {
	return guarded([&] {
		if ((n_utt && (!n_frames || !rank_of)) || !n_ranks || !(step_frames >= 0.0)) throw ApiError(CRFGPU_ERR_ARG, "crfgpu_balance_utts_cost: bad argument");
		balance_utts_cost(n_utt, n_frames, n_ranks, n_slots, step_frames, rank_of);
	});
}

uint32_t crfgpu_plan_info(crfgpu_handle h, char* buf, uint32_t cap) {
	if (!h || !buf || !cap) return 0;
	std::string s = plan_text(h);
	const uint32_t n = (uint32_t)std::min<size_t>(s.size(), cap - 1);
	std::memcpy(buf, s.data(), n); buf[n] = 0;
	return n;
}

int crfgpu_host_alloc(void** p, uint64_t bytes) {
	return guarded([&] { if (!p) throw ApiError(CRFGPU_ERR_ARG, "null argument"); CUDA_OK(cudaMallocHost(p, bytes ? bytes : 1)); });
}
int crfgpu_host_free(void* p) {
	return guarded([&] { if (p) CUDA_OK(cudaFreeHost(p)); });
}

}  // extern "C"
