/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of ASR-CRaFT's CRF lattice hot path in plain C.
 *
 * This is the checker the GPU path is compared against in tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg.  Nothing under asr-craft_b200/ may include, link or call it.
 *
 * Parity pinning: every function here is itself checked (tests/test_oracle_vs_reference.py,
 * tests/golden/) against the reference's own C++ compiled from /root/reference by
 * `make -C oracle ref` (oracle/_ref/libcrfref.so), and against the known answers recorded in
 * SURVEY.md section 8(c) for the bundled CRFTrain/test.ascii toy set.
 *
 * All reference citations are relative to /root/reference/.
 */
#ifndef CRF_ORACLE_H
#define CRF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* modeltype, CRF/src/CRF.h:50 */
enum {
	CRFO_STDFRAME = 0, CRFO_STDSEG = 1, CRFO_STDSEG_NO_DUR = 2,
	CRFO_STDSEG_NO_DUR_NO_TRANSFTR = 3, CRFO_STDSEG_NO_DUR_NO_SEGTRANSFTR = 4
};

#define CRFO_LAB_BAD 0xffffffffu   /* CRF_LAB_BAD, CRF/src/io/CRF_FeatureStream.h:15 */
#define CRFO_NO_IDX 0xffffffffu    /* illegal N-state transition, CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:389,401 */

/* Mirrors CRF_FeatureMap_config (CRF/src/ftrmaps/CRF_FeatureMap.h:24-47) + the CRF_Model geometry
 * (CRFTrain/src/Main.cpp:539-575) + window-stream options (CRFTrain/src/Main.cpp:508-515).
 * Same field order as crfref_config in ref_driver.cpp and crfgpu_config in include/crfgpu.h. */
typedef struct crforacle_config {
	uint32_t model_type;
	uint32_t n_labs;
	uint32_t n_base_ftrs;
	uint32_t n_states;
	uint32_t max_dur;
	uint32_t n_actual_labs;
	uint32_t extract_seg_ftrs;
	uint32_t use_state_ftrs, state_fidx_start, state_fidx_end;
	uint32_t use_trans_ftrs, trans_fidx_start, trans_fidx_end;
	uint32_t use_state_bias, use_trans_bias;
	double state_bias_val, trans_bias_val;
	/* context frames of feature stream 1 (ftr1_left_context_len / ftr1_right_context_len / ftr1_use_boundary_delta_ftr,
	 * CRFTrain/src/Main.cpp:508-515): the stream then carries left_ctx + right_ctx more frames per utterance than there are labels */
	uint32_t left_ctx, right_ctx, boundary_delta;
	/* optional second feature stream joined behind the first (ftr2_*, CRFTrain/src/Main.cpp:516-526); n_base_ftrs2 == 0: absent */
	uint32_t n_base_ftrs2, extract_seg_ftrs2, left_ctx2, right_ctx2, boundary_delta2;
} crforacle_config;

const char* crforacle_last_error(void);

uint32_t crforacle_window_width(const crforacle_config* c);
int crforacle_lambda_len(const crforacle_config* c, uint32_t* out);

/* lambda index maps: state_idx[n_labs], trans_idx[n_labs*n_labs] indexed [plab*n_labs+clab]
 * (CRF_StdFeatureMap::recalc, CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:472-517) */
int crforacle_index_maps(const crforacle_config* c, uint32_t* state_idx, uint32_t* trans_idx);

int crforacle_window_ftrs(const crforacle_config* c, uint32_t n_frames, const float* base_ftrs, float* out);
/* joined / context windows: stream s holds T + left_ctx_s + right_ctx_s frames for an utterance of T labelled frames (base2 may be NULL) */
int crforacle_window_ftrs2(const crforacle_config* c, uint32_t n_frames, const float* base_ftrs, const float* base_ftrs2, float* out);
int crforacle_window_labs(const crforacle_config* c, uint32_t n_frames, const uint32_t* frame_labs, uint32_t* out4);

/* grad accumulated into (caller zeroes); numer/logZ per utterance.  n_threads shards utterances
 * contiguously like CRF_FeatureStreamManager.cpp:425-464. */
int crforacle_fwdbwd_mt(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                        uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const uint32_t* frame_labs,
                        double* grad, double* numer, double* logZ, uint32_t n_threads);

/* The same with the second feature stream: utterance u's rows of stream s start at row frame_off[u] + u * (left_ctx_s + right_ctx_s)
 * and number T_u + left_ctx_s + right_ctx_s (the padded pfile of the TIMIT recipe, demo/segmental-timit-demo.cfg.in:16-33) */
int crforacle_fwdbwd_mt2(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                         uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2, const uint32_t* frame_labs,
                         double* grad, double* numer, double* logZ, uint32_t n_threads);
int crforacle_viterbi2(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                       uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                       uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                       float* path_cost, double* logZ);

/* Decoding against a phone-bigram LM (one state per phone): lm_start[P], lm_bigram[P*P] at [from*P + to] (diagonal unused),
 * lm_final[P] (+inf: not a final state); path_cost then includes the final weight.  All three NULL = crforacle_viterbi2. */
int crforacle_viterbi_lm(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                         uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                         const float* lm_start, const float* lm_bigram, const float* lm_final,
                         uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                         float* path_cost, double* logZ);

/* ... with beam pruning (nStateDecode's input_beam > 0; one state per phone): hypotheses of a node with weight >= min + beam are dropped */
int crforacle_viterbi_beam(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                           uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                           const float* lm_start, const float* lm_bigram, const float* lm_final, double beam,
                           uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                           float* path_cost, double* logZ);

/* Same as crforacle_fwdbwd_mt(…,1) for one utterance, additionally returning alpha/beta
 * ([T][n_labs] doubles, entries the reference never computes are set to -DBL_MAX = LOG0). */
int crforacle_fwdbwd_dump(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                          uint32_t n_frames, const float* base_ftrs, const uint32_t* frame_labs,
                          double* grad, double* numer, double* logZ, double* alpha, double* beta);

int crforacle_viterbi(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                      uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs,
                      uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                      float* path_cost, double* logZ);

#ifdef __cplusplus
}
#endif
#endif
