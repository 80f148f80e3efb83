/*
 * crfgpu.h -- C ABI of libcrfgpu.so, the B200 (sm_100a) implementation of ASR-CRaFT's CRF lattice
 * hot path.  Plain pointers and sizes only; every entry point returns 0 on success or a CRFGPU_ERR_*
 * code, never throws across the boundary, and fails loudly (no CPU fallback) when no CUDA device
 * is usable.  One handle is thread-compatible (use one handle per host thread / per GPU).
 *
 * Each entry point names the reference interface it replaces (paths relative to the ASR-CRaFT tree):
 *
 *   crfgpu_create / crfgpu_lambda_len / crfgpu_index_maps
 *       CRF_FeatureMap::createFeatureMap(cfg) + CRF_StdFeatureMap::recalc()       CRF/src/ftrmaps/CRF_FeatureMap.cpp:55-74,
 *       and CRF_Model::setFeatureMap / setLabMaxDur / setNActualLabs / setModelType  CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:472-517,
 *                                                                                   CRF/src/CRF_Model.cpp:75-87
 *   crfgpu_set_lambda            CRF_Model::getLambda() being read by the nodes     CRF/src/CRF_Model.cpp:105-131
 *   crfgpu_sgd_update            the per-minibatch update of CRF_SGTrainer::sgtrainMinibatch (SGD / AdaGrad / gvar quirk /
 *   crfgpu_get_lambda            lambdaAcc, lambdaSqrAcc)                            CRF/src/trainers/CRF_SGTrainer.cpp:299-325
 *   crfgpu_set_train_state       and the arrays its checkpoints write                CRF/src/trainers/CRF_SGTrainer.cpp:336-380
 *   crfgpu_fwdbwd_batch          CRF_Minibatch_GradAccumulator::accumulateGradient  CRF/src/trainers/accumulators/CRF_Minibatch_GradAccumulator.cpp:201-322
 *                                = a loop of CRF_GradBuilder::buildGradient         CRF/src/trainers/gradbuilders/CRF_GradBuilder.h:40
 *                                (CRF_NewGradBuilder.cpp:48-382, CRF_NewGradBuilder_StdSeg.cpp:31-377)
 *   crfgpu_viterbi_batch         CRF_ViterbiDecoder_StdSeg_NoSegTransFtr<CRF_ViterbiNode>::nStateDecode
 *                                with lm_fst==NULL, beam 0                           CRF/src/decoders/CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:1369-2398
 *   crfgpu_expand_windows        CRF_InFtrStream_SeqMultiWindow::read_ftrs           CRF/src/io/CRF_InFtrStream_SeqMultiWindow.cpp:209-328
 *   crfgpu_*2                    the same calls for models whose window stream has context frames (first_frame_left_ctx_ftrs,
 *                                first/last_frame_right_ctx_ftrs, boundary_delta_ftrs, .cpp:897-1110) and / or a second stream joined
 *                                behind the first (CRF_FeatureStreamManager::join, CRF/src/io/CRF_FeatureStreamManager.cpp:482-500)
 *   crfgpu_prefetch_batch        the bunch read-ahead of the feature streams         CRF/src/io/CRF_FeatureStream.cpp:116-136
 *   crfgpu_group_labels          CRF_InLabStream_SeqMultiWindow::nextseg/read_labs   CRF/src/io/CRF_InLabStream_SeqMultiWindow.cpp:51-306
 *   crfgpu_comm_* /              the serial sum of the per-stream gradients and scalars CRF/src/trainers/accumulators/CRF_Minibatch_GradAccumulator.cpp:277-298
 *   crfgpu_allreduce_grad        (one NCCL all-reduce of lambda_len+4 doubles per minibatch; the caller keeps the "/ nStreams_active" of :306-308)
 *   crfgpu_shard_views /         the contiguous per-stream corpus views and the per-stream minibatch shares
 *   crfgpu_balance_utts          CRF/src/io/CRF_FeatureStreamManager.cpp:425-464, CRF_Minibatch_GradAccumulator.cpp:229-257
 *
 * Inputs are the UN-windowed streams (what the pfile / ilab hold): a ragged batch of utterances,
 * `frame_off[n_utt+1]` frame offsets, `base_ftrs[sum T][n_base_ftrs]` floats and one phone label per
 * frame.  The segment windows (CRF_InFtrStream_SeqMultiWindow) and the (label,start,end) grouping
 * (CRF_InLabStream_SeqMultiWindow) are applied on the device / inside the call, so the D-fold inflated
 * window stream never crosses PCIe.
 */
#ifndef CRFGPU_H
#define CRFGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* modeltype, CRF/src/CRF.h:50 */
enum {
	CRFGPU_STDFRAME = 0,
	CRFGPU_STDSEG = 1,
	CRFGPU_STDSEG_NO_DUR = 2,
	CRFGPU_STDSEG_NO_DUR_NO_TRANSFTR = 3,
	CRFGPU_STDSEG_NO_DUR_NO_SEGTRANSFTR = 4
};

enum {
	CRFGPU_OK = 0,
	CRFGPU_ERR_ARG = 1,          /* bad argument / geometry (what the reference would throw runtime_error for) */
	CRFGPU_ERR_UNSUPPORTED = 2,  /* valid in the reference, not implemented on the device yet (never silently emulated) */
	CRFGPU_ERR_CUDA = 3,         /* CUDA runtime failure, including "no device" */
	CRFGPU_ERR_NUMERIC = 4       /* posterior-mass check failed / empty log-sum (reference: overflow_error / runtime_error) */
};

#define CRFGPU_LAB_BAD 0xffffffffu  /* CRF_LAB_BAD, CRF/src/io/CRF_FeatureStream.h:15 */
#define CRFGPU_NO_IDX 0xffffffffu   /* illegal N-state transition, CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:389,401 */

/* Mirrors CRF_FeatureMap_config (CRF/src/ftrmaps/CRF_FeatureMap.h:24-47) + the CRF_Model geometry set by
 * CRFTrain (CRFTrain/src/Main.cpp:539-575) + the window-stream options (CRFTrain/src/Main.cpp:508-515). */
typedef struct crfgpu_config {
	uint32_t model_type;        /* CRFGPU_STD* */
	uint32_t n_labs;            /* crf_label_size (stdseg: phones * max_dur) */
	uint32_t n_base_ftrs;       /* width of the un-windowed feature stream */
	uint32_t n_states;          /* crf_states */
	uint32_t max_dur;           /* label_maximum_duration == ftr1_window_len */
	uint32_t n_actual_labs;     /* num_actual_labs */
	uint32_t extract_seg_ftrs;  /* ftr1_extract_seg_ftr: windows carry [5 samples|avg|max|min|one-hot dur] */
	uint32_t use_state_ftrs, state_fidx_start, state_fidx_end;   /* indices into the WINDOW feature vector, inclusive */
	uint32_t use_trans_ftrs, trans_fidx_start, trans_fidx_end;   /* crf_featuremap=stdtrans: training with <= 169 labels (phones x states) frame-level, <= 161 in stdseg_no_dur_no_segtransftr; decoding for those two model kinds up to 1024 labels */
	uint32_t use_state_bias, use_trans_bias;
	double state_bias_val, trans_bias_val;
	/* Context frames of feature stream 1 (ftr1_left_context_len, ftr1_right_context_len, ftr1_use_boundary_delta_ftr;
	 * CRFTrain/src/Main.cpp:508-515, CRF/src/io/CRF_InFtrStream_SeqMultiWindow.cpp:47-117): the stream then carries
	 * left_ctx + right_ctx more frames per utterance than there are labelled frames (a padded pfile). */
	uint32_t left_ctx, right_ctx, boundary_delta;
	/* Optional second feature stream whose windows are joined behind those of the first (ftr2_*; CRFTrain/src/Main.cpp:516-526,
	 * CRF_FeatureStream::join, CRF/src/io/CRF_FeatureStream.cpp:172-184).  n_base_ftrs2 == 0: absent.  The TIMIT recipe
	 * (demo/segmental-timit-demo.cfg.in:16-33) takes its 1162 state features from stream 1 (segment features of 144 inputs, maxDur 10)
	 * and its 1872 transition features from stream 2 (13 context frames x 144 around the first frame of the window). */
	uint32_t n_base_ftrs2, extract_seg_ftrs2, left_ctx2, right_ctx2, boundary_delta2;
} crfgpu_config;

typedef struct crfgpu_ctx* crfgpu_handle;

/* Error text of the last failing call on this thread (valid also when crfgpu_create failed). */
const char* crfgpu_last_error(void);

int crfgpu_create(const crfgpu_config* cfg, int device, crfgpu_handle* out);
int crfgpu_destroy(crfgpu_handle h);

uint32_t crfgpu_window_width(const crfgpu_config* cfg);
uint32_t crfgpu_lambda_len(crfgpu_handle h);
/* state_idx[n_labs], trans_idx[n_labs*n_labs] indexed [plab*n_labs+clab]; CRFGPU_NO_IDX for illegal pairs */
int crfgpu_index_maps(crfgpu_handle h, uint32_t* state_idx, uint32_t* trans_idx);

/* Copies lambda (host, double) to the device and derives the device-side tables. */
int crfgpu_set_lambda(crfgpu_handle h, const double* lambda, uint32_t len);

/* Trainer update on the device, so that lambda never leaves HBM between minibatches: with g = staged gradient / n_active
 * (CRF_Minibatch_GradAccumulator.cpp:306-308; after an all-reduce the gradient at crfgpu_device_results is the global sum),
 *   if use_gvar:    g -= g * inv_square_var                      (CRF_SGTrainer.cpp:300-303, the reference's own formula)
 *   if use_adagrad: gradSqrAcc += g*g; lambda += eta / (sqrt(gradSqrAcc) + eps) * g
 *   else:           lambda += lr * g                             (lr: the trainer's float learning rate promoted to double)
 *   lambdaAcc += lambda; lambdaSqrAcc += lambda*lambda           (what the .avg.out checkpoints are built from)
 * and every lambda-derived device table is rebuilt by a kernel.  The staged gradient is consumed. */
typedef struct crfgpu_sgd {
	double lr;
	uint32_t use_gvar; double inv_square_var;
	uint32_t use_adagrad; double eta, eps;
} crfgpu_sgd;
int crfgpu_sgd_update(crfgpu_handle h, const crfgpu_sgd* opt, double n_active);
/* Device -> host copies of lambda and (any may be NULL) lambdaAcc, lambdaSqrAcc, gradSqrAcc [lambda_len doubles each]. */
int crfgpu_get_lambda(crfgpu_handle h, double* lambda, double* lambda_acc, double* lambda_sqr_acc, double* grad_sqr_acc);
/* Resume: host -> device copies of the accumulators (NULL = zeros). */
int crfgpu_set_train_state(crfgpu_handle h, const double* lambda_acc, const double* lambda_sqr_acc, const double* grad_sqr_acc);

/* ---- host-buffer entry points (the drop-in calls; H2D + kernels + D2H inside) --------------------
 * grad[lambda_len] is OVERWRITTEN with sum over the batch of (empirical - expected) feature counts,
 * numer[n_utt] = sum lambda.f on the reference path, logZ[n_utt] = log partition function. */
int crfgpu_fwdbwd_batch(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off,
                        const float* base_ftrs, const uint32_t* frame_labs,
                        double* grad, double* numer, double* logZ);

/* Best path per utterance as segments, written at out_*[frame_off[u] .. frame_off[u]+n_seg[u]):
 * out_lab = sub-state label of the segment, out_dur = its duration in frames, out_phn = phone id
 * emitted when the segment starts a phone (else CRFGPU_LAB_BAD), path_cost = float cost of the path
 * (99999.0 and n_seg 0 when the end state cannot be reached, as the reference reports). */
int crfgpu_viterbi_batch(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs,
                         uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                         float* path_cost);

/* Decoding against a phone-bigram language model (nStateDecode's lm_fst, CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:1369-1372, for the
 * class of LMs that has the topology of the decoder's own free-phone LM, createFreePhoneLmFst :1270-1348, with one state per phone:
 * state "the last phone was p", an arc p -> q for every q != p, every phone state possibly final).  lm_start[P] = weight of the arc
 * start -> q, lm_bigram[from*P + to] = weight of the arc from -> to (diagonal unused; every other entry finite: the LM must be complete),
 * lm_final[P] = final weight of state p (+inf: not final).  Weights are costs (negative log probabilities) added float by float where
 * the reference adds them -- (hypothesis + arc) + transition -- and the best final hypothesis is taken over the LM states in state
 * order with the final weight added (expandFinalNode :746-758, :2138-2153); path_cost then includes it.  All three NULL drops the LM.
 * Epsilon / back-off arcs, word-level LMs, beam pruning and lattice output are not implemented (CRFGPU_ERR_UNSUPPORTED). */
int crfgpu_set_phone_lm(crfgpu_handle h, const float* lm_start, const float* lm_bigram, const float* lm_final);
/* The same for models with N > 1 states per phone, whose free-phone LM in the reference returns from every phone state to the start
 * state through an epsilon arc and leaves it again on one arc per phone (:1313-1330): lm_unigram[P] = cost of the arc start -> q (also
 * paid by the first phone), lm_exit[P] = cost of the epsilon arc of phone p (a phone insertion penalty when all equal), lm_final[P] as
 * above.  Added where the reference adds them: ((hypothesis + exit) + unigram) + transition. */
int crfgpu_set_phone_unigram_lm(crfgpu_handle h, const float* lm_unigram, const float* lm_exit, const float* lm_final);
/* nStateDecode's input_beam (one state per phone): after every frame the decoder drops the hypotheses whose weight is not below the frame's
 * minimum + beam (pruning(), .cpp:976-1106) and expands only the survivors (:573); 0 = no pruning (the default).  The results are the
 * reference's for the same beam, bit for bit.  min_hyps / max_hyps / beam_inc are accepted by the reference's signature and never used. */
int crfgpu_set_beam(crfgpu_handle h, double beam);

/* Window features for one utterance, out[(t*max_dur + d-1)*window_width ...]; slots with d > t+1 are zero. */
int crfgpu_expand_windows(crfgpu_handle h, uint32_t n_frames, const float* base_ftrs, float* out);

/* ---- models with context frames and / or a joined second feature stream -------------------------------------------------------
 * frame_off[] are the offsets of the LABELLED frames, as everywhere else.  Stream s (1 or 2) carries left_ctx_s + right_ctx_s more
 * frames per utterance: utterance u's rows start at row frame_off[u] + u * (left_ctx_s + right_ctx_s) of base_ftrs_s and number
 * T_u + left_ctx_s + right_ctx_s; labelled frame t of the utterance is row left_ctx_s + t of them (nextseg(),
 * CRF_InFtrStream_SeqMultiWindow.cpp:163-199).  base_ftrs2 is ignored (may be NULL) when the configuration has no second stream.
 * Without context frames and without a second stream these calls equal the ones above. */
int crfgpu_stage_batch2(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                        const uint32_t* frame_labs /* may be NULL for decode */);
int crfgpu_fwdbwd_batch2(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                         const uint32_t* frame_labs, double* grad, double* numer, double* logZ);
int crfgpu_viterbi_batch2(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                          uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg, float* path_cost);
/* base_ftrs: [left_ctx + n_frames + right_ctx][n_base_ftrs], base_ftrs2 likewise with the second stream's context lengths */
int crfgpu_expand_windows2(crfgpu_handle h, uint32_t n_frames, const float* base_ftrs, const float* base_ftrs2, float* out);
/* (label,start,end,broken) records per frame or CRFGPU_LAB_BAD x4 (host-side, no device work). */
int crfgpu_group_labels(const crfgpu_config* cfg, uint32_t n_frames, const uint32_t* frame_labs, uint32_t* out4);

/* ---- device-resident entry points (inputs already in HBM; used for kernel-only timing and by
 * multi-GPU drivers that all-reduce the gradient in place before reading it back) ----------------- */
/* Stage a batch: copies offsets/features/labels to the device (async on the handle's stream). */
int crfgpu_stage_batch(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off,
                       const float* base_ftrs, const uint32_t* frame_labs /* may be NULL for decode */);
/* Optional: start copying and window-expanding the NEXT minibatch into a second buffer set on side streams, to be called after
 * crfgpu_fwdbwd_staged of the current one (the data-loader prefetch of a training loop; it does not touch lambda).  The next
 * crfgpu_stage_batch / crfgpu_fwdbwd_batch that is handed the same frame_off contents and base_ftrs pointer takes the buffers over
 * instead of copying again; any other batch is staged normally.  base_ftrs must stay valid and unchanged until then. */
int crfgpu_prefetch_batch(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs);
/* The same read-ahead for a training loop that also hands over the labels of the next minibatch: the per-frame index and label tables
 * (reference-segment labels of CRF_InLabStream_SeqMultiWindow's grouping, previous / next segment labels) are built on the host and
 * copied while the current minibatch computes, so that the crfgpu_stage_batch that takes the batch over (same frame_off contents, same
 * base_ftrs and frame_labs pointers) has no per-frame host work left. */
int crfgpu_prefetch_train_batch(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const uint32_t* frame_labs);
/* ... and of a model with context frames / a joined second stream (crfgpu_stage_batch2's layout of both streams; frame_labs may be
 * NULL): copies, joined windows and label tables of the NEXT minibatch on the side stream, taken over by crfgpu_stage_batch2 /
 * crfgpu_fwdbwd_batch2 when handed the same pointers and offsets. */
int crfgpu_prefetch_train_batch2(crfgpu_handle h, uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                                 const uint32_t* frame_labs);
/* Run forward-backward+gradient on the staged batch; results stay on the device. */
int crfgpu_fwdbwd_staged(crfgpu_handle h);
/* Run Viterbi + traceback on the staged batch; results stay on the device. */
int crfgpu_viterbi_staged(crfgpu_handle h);
/* Device pointers of the staged results (valid until the next stage/destroy).  d_grad has lambda_len+4
 * doubles: the gradient followed by [sum numer, sum logZ, n_utt, 0], so one all-reduce moves everything. */
int crfgpu_device_results(crfgpu_handle h, double** d_grad, double** d_numer, double** d_logZ);
int crfgpu_fetch_fwdbwd(crfgpu_handle h, double* grad, double* numer, double* logZ);
int crfgpu_fetch_viterbi(crfgpu_handle h, uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn,
                         uint32_t* n_seg, float* path_cost);
int crfgpu_synchronize(crfgpu_handle h);
/* The CUDA stream (cudaStream_t) all work of this handle is issued on. */
void* crfgpu_stream(crfgpu_handle h);
/* Kernel launches issued by this handle since creation (for bench.py's gpu_launches). */
uint64_t crfgpu_launch_count(crfgpu_handle h);
/* Elapsed device milliseconds of the named phase in the last staged run (CUDA events on the stream):
 * "expand","score","forward","backward","xi","grad","viterbi_score","viterbi". <0 if unknown. */
double crfgpu_phase_ms(crfgpu_handle h, const char* phase);

/* Debug/test access to the lattice of the staged batch after crfgpu_fwdbwd_staged: log-domain
 * alpha/beta [sum T][n_labs] as doubles (entries the reference never computes are -DBL_MAX). */
int crfgpu_fetch_alpha_beta(crfgpu_handle h, double* alpha, double* beta);

/* The posterior-mass assertion of the nodes' computeExpF (CRF_StdStateNode.cpp:252-275, CRF_StdSegStateNode.cpp:417-436) runs on the device
 * after every forward-backward: mass[n] = sum over labels of the posterior of frame n (frame-level models: 1; segmental models: the
 * probability that a segment ends on frame n).  Frames outside the reference's band (frame-level 0.9..1.1; segmental 0..1, +-1e-3 for the
 * device's fp32 posteriors) or NaN make crfgpu_fetch_fwdbwd / crfgpu_fetch_tail / crfgpu_fwdbwd_batch return CRFGPU_ERR_NUMERIC, the way
 * the reference throws.  This call returns the masses themselves [sum T floats] (test / diagnosis access). */
int crfgpu_fetch_posterior_mass(crfgpu_handle h, float* mass);

/* Tuning / debug switches (results do not depend on them beyond the stated tolerance; decoding stays bit-exact):
 *  "slots" (utterances per CTA in the one-CTA lattice kernels: 0 auto,1,2,4,8), "k_slab" / "k_slab_tc" / "k_slab_tma" / "k_slab_xi"
 *  (frames per CTA in the reduce-GEMMs), "keep_lattice" (1: keep what crfgpu_fetch_alpha_beta needs),
 *  "dp_impl" (dense lattice: 0 one CTA per utterance group, 1 cluster-resident FFMA, 2 cluster-resident tcgen05 = default),
 *  "frame_impl" (frame-level models with <= 64 labels: 0 one warp per utterance = default, 1 the cluster lattice kernels),
 *  "nodur_impl" (stdseg_no_dur*: 0 auto, 1 native O(P^2 + D*P) recursion, 2 tied (duration, label) expansion; set_lambda again after it),
 *  "vit_impl" (Viterbi recursion: 0 auto, 1 one CTA per utterance, 2 transition table sliced over groups of CTAs -- one state per phone),
 *  "aux_empirical" (1 = default: the empirical counts / numerators run on a side stream beside the score GEMM and the recursions; 0: behind
 *  the state gradient),
 *  "tf_tiled" (transition-feature GEMMs: 1 = default, operands split into bf16 hi / lo and tiled once, ring stages by bulk copies; 0 = the
 *  register-staged kernels on the fp32 arrays; set_lambda again after it),
 *  "vit_eager" (decode batches: 1 = default, crfgpu_stage_batch launches the recursion of each H2D chunk's utterances behind the chunk on
 *  a side stream; 0 = one launch for the whole batch in crfgpu_viterbi_staged), "virt_windows" (0: the training GEMMs read fully
 *  materialised windows),
 *  "gemm_impl" / "tma_mask" (which GEMM kernels run: FFMA, register-staged tcgen05, TMA-fed tcgen05 with the 128-row operand in TMEM),
 *  "cluster_slots", "prefetch_smem".  Environment: CRFGPU_VERBOSE (plans, device timeline), CRFGPU_DP_TIMING (cycle counters of the
 *  recursion kernels on stderr). */
int crfgpu_set_option(crfgpu_handle h, const char* name, int64_t value);

/* ---- multi-GPU training: one process (or host thread) per GPU, one handle each ----------------------------------------------
 * The data-parallel exchange of the reference is the serial host sum of nStreams gradients (CRF_Minibatch_GradAccumulator.cpp:277-298);
 * here it is ONE ncclAllReduce(sum) over the staged device gradient and its tail [sum numer, sum logZ, n_utt, 0], issued on the
 * handle's stream behind the gradient kernels, so crfgpu_sgd_update can consume the global sum without a host round trip.
 * NCCL is loaded with dlopen("libnccl.so.2") on first use (CRFGPU_NCCL_LIB overrides the name); CRFGPU_ERR_UNSUPPORTED if absent.
 *   multi-process:  rank 0 calls crfgpu_comm_unique_id, ships the 128 bytes to the other ranks by any means (file, MPI, torch
 *                   broadcast ...), every rank calls crfgpu_comm_init_rank;
 *   one process:    crfgpu_comm_init_all over one handle per device; collectives of several handles issued by ONE thread must be
 *                   bracketed by crfgpu_group_start / crfgpu_group_end (ncclGroupStart/End). */
#define CRFGPU_COMM_ID_BYTES 128
int crfgpu_comm_unique_id(void* id128);
int crfgpu_comm_init_rank(crfgpu_handle h, int n_ranks, int rank, const void* id128);
int crfgpu_comm_init_all(crfgpu_handle* handles, int n);
int crfgpu_comm_destroy(crfgpu_handle h);
int crfgpu_comm_size(crfgpu_handle h);            /* ranks of the handle's communicator, 0 if none */
int crfgpu_group_start(void);
int crfgpu_group_end(void);
/* In-place all-reduce (sum) of the staged gradient + tail over the handle's communicator; asynchronous on crfgpu_stream(h). */
int crfgpu_allreduce_grad(crfgpu_handle h);
/* The 4 tail doubles behind the staged gradient: [sum numer, sum logZ, n_utt, 0] (global sums after crfgpu_allreduce_grad). */
int crfgpu_fetch_tail(crfgpu_handle h, double* tail4);

/* Host-side sharding rules (no device work).
 * crfgpu_shard_views: the reference's contiguous corpus views -- stream i of n_streams owns utterances
 *   [i*floor(n/n_streams), (i+1)*floor(n/n_streams)), the last stream also the remainder (CRF_FeatureStreamManager.cpp:425-464);
 *   writes first[n_streams] and count[n_streams].
 * crfgpu_minibatch_share: utterances stream i takes per minibatch, floor(mb/n_streams) + (i < mb % n_streams)
 *   (CRF_Minibatch_GradAccumulator.cpp:229-257); minibatch 0 = the whole view (0xffffffff is returned).
 * crfgpu_balance_utts: assignment of the utterances of ONE global minibatch to n_ranks devices, equal counts (+-1), longest-first
 *   onto the rank with the fewest frames so far -- membership of the minibatch is untouched and the gradient sum is order-free
 *   (SURVEY.md 8e), only the device that computes an utterance changes; writes rank_of[n_utt]. */
int crfgpu_shard_views(uint32_t n_utt, uint32_t n_streams, uint32_t* first, uint32_t* count);
uint32_t crfgpu_minibatch_share(uint32_t minibatch, uint32_t n_streams, uint32_t stream);
int crfgpu_balance_utts(uint32_t n_utt, const uint32_t* n_frames, uint32_t n_ranks, uint32_t* rank_of);
/* The same deal by a time model instead of equal counts: cost of a rank = step_frames * lock-steps + frames, lock-steps = the largest
 * load of the rank's n_slots slots when it deals its utterances longest first to the least loaded slot, as crfgpu_stage_batch does (the
 * model carries those loads along exactly) -- the lattice kernels are a dependent chain of lock-steps (n_slots = utterances a device
 * advances side by side: resident clusters x 16, see crfgpu_plan_info), everything else streams the frames;
 * step_frames = measured time of one lock-step / measured time per frame of the rest of the step.  A rank holding one of the corpus'
 * longest utterances gets fewer frames.  Membership of the global minibatch is unchanged. */
int crfgpu_balance_utts_cost(uint32_t n_utt, const uint32_t* n_frames, uint32_t n_ranks, uint32_t n_slots, double step_frames, uint32_t* rank_of);

/* Which kernels the handle runs for the staged batch, as one line of text: lattice implementation and its plan (clusters, slots,
 * lock-steps = frames of the longest slot list), GEMM family, Viterbi variant -- so that no geometry changes implementation
 * without the caller being able to see it.  Returns the number of bytes written (without the NUL), 0 on error. */
uint32_t crfgpu_plan_info(crfgpu_handle h, char* buf, uint32_t cap);

/* Pinned host memory helpers for callers that want asynchronous copies. */
int crfgpu_host_alloc(void** p, uint64_t bytes);
int crfgpu_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
