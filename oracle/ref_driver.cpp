/*
 * TEST INFRASTRUCTURE ONLY.  Driver that links the UNMODIFIED ASR-CRaFT hot-path sources
 * (compiled in place from /root/reference/CRF/src by oracle/Makefile, against the stub
 * QuickNet/OpenFst headers under oracle/stubs/) and exposes them through a small C ABI so
 * Python tests (ctypes) and bench.py's cpu_baseline / --impl reference legs can call the
 * reference's own implementation:
 *
 *   crfref_fwdbwd      -> CRF_GradBuilder::create()->buildGradient()
 *                         (CRF/src/trainers/gradbuilders/CRF_GradBuilder.cpp:97-162,
 *                          CRF_NewGradBuilder.cpp:48-382, CRF_NewGradBuilder_StdSeg.cpp:31-377,
 *                          CRF_NewGradBuilder_StdSeg_NoDur_NoTrans.cpp:65-492)
 *   crfref_fwdbwd_mt   -> same, fanned out over pthreads with the contiguous-view sharding of
 *                         CRF/src/io/CRF_FeatureStreamManager.cpp:425-464 and the serial reduce
 *                         of CRF/src/trainers/accumulators/CRF_Minibatch_GradAccumulator.cpp:277-298
 *                         (without the /N_active, which the caller applies)
 *   crfref_viterbi     -> CRF_ViterbiDecoder_StdSeg_NoSegTransFtr<CRF_ViterbiNode>::nStateDecode()
 *                         (CRF/src/decoders/CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:1369-2398)
 *   crfref_viterbi_old -> CRF_ViterbiDecoder::nStateDecode() (CRF/src/decoders/CRF_ViterbiDecoder.cpp:398-1035)
 *   crfref_window_ftrs -> CRF_InFtrStream_SeqMultiWindow::read_ftrs (CRF/src/io/CRF_InFtrStream_SeqMultiWindow.cpp:209-328)
 *   crfref_window_labs -> CRF_InLabStream_SeqMultiWindow::read_labs (CRF/src/io/CRF_InLabStream_SeqMultiWindow.cpp:246-306)
 *
 * Frames are served to the reference through in-memory QN_InFtrStream / QN_InLabStream
 * subclasses wrapped in the reference's own SeqMultiWindow streams, exactly as
 * CRF_FeatureStreamManager does (CRF/src/io/CRF_FeatureStreamManager.cpp:202-233,344-378).
 *
 * Nothing in the product (asr-craft_b200/) links or calls this file.
 */
#include <pthread.h>
#include <unistd.h>
#include <fcntl.h>
#include <vector>
#include <string>
#include <stdexcept>

#include "CRF.h"
#include "CRF_Model.h"
#include "ftrmaps/CRF_FeatureMap.h"
#include "io/CRF_FeatureStream.h"
#include "io/CRF_InFtrStream_SeqMultiWindow.h"
#include "io/CRF_InLabStream_SeqMultiWindow.h"
#include "trainers/gradbuilders/CRF_GradBuilder.h"
#include "decoders/CRF_ViterbiDecoder.h"
#include "decoders/CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.h"

extern "C" {

/* Mirrors CRF_FeatureMap_config (CRF/src/ftrmaps/CRF_FeatureMap.h:24-47) plus the CRF_Model
 * geometry set by CRFTrain (CRFTrain/src/Main.cpp:539-575) and the window-stream options
 * (CRFTrain/src/Main.cpp:508-515).  Field order is shared with oracle/crf_oracle.h. */
struct crfref_config {
	uint32_t model_type;      /* modeltype enum, CRF/src/CRF.h:50: 0 stdframe,1 stdseg,2 no_dur,3 no_dur_no_transftr,4 no_dur_no_segtransftr */
	uint32_t n_labs;          /* crf_label_size */
	uint32_t n_base_ftrs;     /* width of the un-windowed feature stream */
	uint32_t n_states;        /* crf_states */
	uint32_t max_dur;         /* label_maximum_duration == ftr window_len */
	uint32_t n_actual_labs;   /* num_actual_labs */
	uint32_t extract_seg_ftrs;/* ftr1_extract_seg_ftr */
	uint32_t use_state_ftrs, state_fidx_start, state_fidx_end;
	uint32_t use_trans_ftrs, trans_fidx_start, trans_fidx_end;
	uint32_t use_state_bias, use_trans_bias;
	double state_bias_val, trans_bias_val;
	/* ftr1 context options and the optional joined second stream (CRFTrain/src/Main.cpp:508-526); same order as crforacle_config */
	uint32_t left_ctx, right_ctx, boundary_delta;
	uint32_t n_base_ftrs2, extract_seg_ftrs2, left_ctx2, right_ctx2, boundary_delta2;
};

}  // extern "C"

namespace {

thread_local std::string g_err;

/* In-memory utterance store served as a QuickNet feature stream. */
class MemFtrStream : public QN_InFtrStream {
	const float* data_; const uint32_t* off_; size_t nutt_, width_;
	long seg_; size_t pos_;
public:
	MemFtrStream(const float* d, const uint32_t* off, size_t nutt, size_t width)
		: data_(d), off_(off), nutt_(nutt), width_(width), seg_(-1), pos_(0) {}
	size_t num_ftrs() { return width_; }
	QN_SegID nextseg() {
		if (seg_ + 1 >= (long)nutt_) { seg_ = (long)nutt_; return QN_SEGID_BAD; }
		seg_++; pos_ = 0; return seg_;
	}
	size_t read_ftrs(size_t cnt, float* ftrs) {
		if (seg_ < 0 || seg_ >= (long)nutt_) return 0;
		size_t len = off_[seg_ + 1] - off_[seg_];
		size_t n = (pos_ + cnt <= len) ? cnt : len - pos_;
		memcpy(ftrs, data_ + ((size_t)off_[seg_] + pos_) * width_, n * width_ * sizeof(float));
		pos_ += n; return n;
	}
	int rewind() { seg_ = -1; pos_ = 0; return QN_OK; }
	size_t num_segs() { return nutt_; }
	size_t num_frames(size_t segno = QN_ALL) {
		if (segno == QN_ALL) return off_[nutt_] - off_[0];
		return off_[segno + 1] - off_[segno];
	}
	int get_pos(size_t* s, size_t* f) { *s = seg_; *f = pos_; return QN_OK; }
	QN_SegID set_pos(size_t s, size_t f) { seg_ = (long)s; pos_ = f; return seg_; }
};

class MemLabStream : public QN_InLabStream {
	const uint32_t* data_; const uint32_t* off_; size_t nutt_;
	long seg_; size_t pos_;
public:
	MemLabStream(const uint32_t* d, const uint32_t* off, size_t nutt)
		: data_(d), off_(off), nutt_(nutt), seg_(-1), pos_(0) {}
	size_t num_labs() { return 1; }
	QN_SegID nextseg() {
		if (seg_ + 1 >= (long)nutt_) { seg_ = (long)nutt_; return QN_SEGID_BAD; }
		seg_++; pos_ = 0; return seg_;
	}
	size_t read_labs(size_t cnt, QNUInt32* labs) {
		if (seg_ < 0 || seg_ >= (long)nutt_) return 0;
		size_t len = off_[seg_ + 1] - off_[seg_];
		size_t n = (pos_ + cnt <= len) ? cnt : len - pos_;
		memcpy(labs, data_ + (size_t)off_[seg_] + pos_, n * sizeof(QNUInt32));
		pos_ += n; return n;
	}
	int rewind() { seg_ = -1; pos_ = 0; return QN_OK; }
	size_t num_segs() { return nutt_; }
	size_t num_frames(size_t segno = QN_ALL) {
		if (segno == QN_ALL) return off_[nutt_] - off_[0];
		return off_[segno + 1] - off_[segno];
	}
	int get_pos(size_t* s, size_t* f) { *s = seg_; *f = pos_; return QN_OK; }
	QN_SegID set_pos(size_t s, size_t f) { seg_ = (long)s; pos_ = f; return seg_; }
};

/* The reference prints progress to stdout/stderr from constructors, decoders and destructors;
 * silence fd 1/2 while it runs so test output stays readable. */
class Quiet {
	int so_, se_;
public:
	Quiet() {
		fflush(stdout); fflush(stderr); cout.flush(); cerr.flush();
		so_ = dup(1); se_ = dup(2);
		int nul = open("/dev/null", O_WRONLY);
		if (!getenv("CRFREF_VERBOSE")) { dup2(nul, 1); dup2(nul, 2); }
		close(nul);
	}
	~Quiet() {
		fflush(stdout); fflush(stderr); cout.flush(); cerr.flush();
		dup2(so_, 1); dup2(se_, 2); close(so_); close(se_);
	}
};

struct ModelBundle {
	CRF_FeatureMap_config* fcfg;   /* heap: the map keeps the pointer (CRF_FeatureMap.cpp:31-35) */
	CRF_Model* crf;
	ModelBundle(const crfref_config* c, size_t win_ftrs, const double* lambda, uint32_t lambda_len) {
		fcfg = new CRF_FeatureMap_config();
		fcfg->map_type = c->use_trans_ftrs ? STDTRANS : STDSTATE;
		fcfg->numLabs = c->n_labs;
		fcfg->numFeas = (QNUInt32)win_ftrs;
		fcfg->numStates = c->n_states;
		fcfg->useStateFtrs = c->use_state_ftrs != 0;
		fcfg->stateFidxStart = c->state_fidx_start;
		fcfg->stateFidxEnd = c->state_fidx_end;
		fcfg->useTransFtrs = c->use_trans_ftrs != 0;
		fcfg->transFidxStart = c->trans_fidx_start;
		fcfg->transFidxEnd = c->trans_fidx_end;
		fcfg->useStateBias = c->use_state_bias != 0;
		fcfg->useTransBias = c->use_trans_bias != 0;
		fcfg->stateBiasVal = c->state_bias_val;
		fcfg->transBiasVal = c->trans_bias_val;
		fcfg->maxDur = c->max_dur;
		fcfg->durFtrStart = 0;
		fcfg->nActualLabs = c->n_actual_labs;
		crf = new CRF_Model(c->n_labs);
		crf->setLabMaxDur(c->max_dur);
		crf->setNActualLabs(c->n_actual_labs);
		crf->setModelType((modeltype)c->model_type);
		crf->setFeatureMap(CRF_FeatureMap::createFeatureMap(fcfg));
		if (lambda) {
			if (crf->getLambdaLen() != lambda_len)
				throw std::runtime_error("lambda length mismatch: reference map has " +
					std::to_string(crf->getLambdaLen()) + ", caller passed " + std::to_string(lambda_len));
			memcpy(crf->getLambda(), lambda, sizeof(double) * lambda_len);
		}
	}
	~ModelBundle() { delete crf; delete fcfg; }
};

/* Offsets of a stream that carries lc + rc context frames per utterance: utterance u (absolute index u0 + u) starts at row
 * off[u] + (u0 + u) * (lc + rc) and has T_u + lc + rc rows. */
std::vector<uint32_t> padded_offsets(const uint32_t* off, size_t nutt, size_t u0, uint32_t ctx) {
	std::vector<uint32_t> o(nutt + 1);
	for (size_t u = 0; u <= nutt; u++) o[u] = off[u] + (uint32_t)((u0 + u) * ctx);
	return o;
}

struct Streams {
	std::vector<uint32_t> off1, off2;
	MemFtrStream memF; MemFtrStream* memF2; MemLabStream* memL;
	CRF_InFtrStream_SeqMultiWindow winF; CRF_InFtrStream_SeqMultiWindow* winF2; CRF_InLabStream_SeqMultiWindow* winL;
	CRF_FeatureStream* fs; CRF_FeatureStream* fs1; CRF_FeatureStream* fs2;
	/* u0 = absolute index of the first utterance (the shards of crfref_fwdbwd_mt index the shared frame arrays) */
	Streams(const crfref_config* c, const float* ftrs, const float* ftrs2, const uint32_t* labs, const uint32_t* off, size_t nutt, size_t u0 = 0)
		: off1(padded_offsets(off, nutt, u0, c->left_ctx + c->right_ctx)), off2(padded_offsets(off, nutt, u0, c->left_ctx2 + c->right_ctx2)),
		  memF(ftrs, off1.data(), nutt, c->n_base_ftrs), memF2(NULL), memL(NULL),
		  winF(0, "f", memF, c->max_dur, 0, 0, c->left_ctx, c->right_ctx, c->extract_seg_ftrs != 0, c->boundary_delta != 0), winF2(NULL), winL(NULL),
		  fs(NULL), fs1(NULL), fs2(NULL) {
		if (labs) {
			memL = new MemLabStream(labs, off, nutt);
			/* input buffer large enough to hold a whole utterance, so no chunk boundary is ever hit */
			winL = new CRF_InLabStream_SeqMultiWindow(0, "l", *memL, c->max_dur, 0, 0, 65536, 65536);
			fs = new CRF_FeatureStream(&winF, winL);
		} else {
			fs = new CRF_FeatureStream(&winF);
		}
		if (c->n_base_ftrs2) {
			/* the second stream is windowed on its own and joined behind the first: CRF_FeatureStreamManager::join
			 * (CRF/src/io/CRF_FeatureStreamManager.cpp:482-500) -> CRF_FeatureStream::join (CRF/src/io/CRF_FeatureStream.cpp:172-184) */
			if (!ftrs2) throw std::runtime_error("the configuration joins a second feature stream but none was passed");
			memF2 = new MemFtrStream(ftrs2, off2.data(), nutt, c->n_base_ftrs2);
			winF2 = new CRF_InFtrStream_SeqMultiWindow(0, "f2", *memF2, c->max_dur, 0, 0, c->left_ctx2, c->right_ctx2, c->extract_seg_ftrs2 != 0, c->boundary_delta2 != 0);
			fs1 = fs; fs2 = new CRF_FeatureStream(winF2);
			fs = fs1->join(fs2);
		}
	}
	~Streams() { delete fs; if (fs1) { delete fs1; delete fs2; } delete winF2; delete memF2; delete winL; delete memL; }
};

struct ShardArgs {
	const crfref_config* cfg; CRF_Model* crf;
	const float* ftrs; const float* ftrs2; const uint32_t* labs; const uint32_t* off; size_t nutt, u0;
	double* grad; double* numer; double* logZ; std::string err;
};

void run_shard(ShardArgs* a) {
	try {
		Streams st(a->cfg, a->ftrs, a->ftrs2, a->labs, a->off, a->nutt, a->u0);
		CRF_GradBuilder* gb = CRF_GradBuilder::create(a->crf, EXPF);
		st.fs->rewind();
		size_t u = 0;
		while (st.fs->nextseg() != QN_SEGID_BAD) {
			double zx = 0.0;
			double num = gb->buildGradient(st.fs, a->grad, &zx);
			a->numer[u] = num; a->logZ[u] = zx; u++;
		}
		delete gb;
		if (u != a->nutt) throw std::runtime_error("stream ended early");
	} catch (std::exception& e) { a->err = e.what(); }
}

void* shard_thread(void* p) { run_shard((ShardArgs*)p); return NULL; }

size_t part_width(size_t F, size_t D, bool seg, size_t lc, size_t rc, bool bdelta) {
	/* CRF_InFtrStream_SeqMultiWindow ctor (CRF/src/io/CRF_InFtrStream_SeqMultiWindow.cpp:47-117); cross-checked against num_ftrs() of the live streams below */
	if (D == 1) return (lc + 1 + rc) * F;
	if (seg) return 8 * F + D + (lc + rc) * F;
	if (bdelta) return (lc < rc + 1 ? lc : rc + 1) * F;
	return (lc + 1 + rc) * F;
}
size_t window_width(const crfref_config* c) {
	size_t w = part_width(c->n_base_ftrs, c->max_dur, c->extract_seg_ftrs != 0, c->left_ctx, c->right_ctx, c->boundary_delta != 0);
	if (c->n_base_ftrs2) w += part_width(c->n_base_ftrs2, c->max_dur, c->extract_seg_ftrs2 != 0, c->left_ctx2, c->right_ctx2, c->boundary_delta2 != 0);
	return w;
}

/* Subclass used only to observe the decoder's per-frame survivors, so the (label,duration)
 * sequence of the best path can be reported: the reference's linear output FST carries one arc
 * per segment but no duration, and nStateDecode clears viterbiDurs while tracing back
 * (CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:2340-2346).  pruning() and expandFinalNode() are
 * virtual in the reference; the overrides call the base implementation first and only copy state. */
class ObservedDecoder : public CRF_ViterbiDecoder_StdSeg_NoSegTransFtr<CRF_ViterbiNode> {
public:
	std::vector<std::vector<uint> > phnIds; std::vector<std::vector<int> > ptrs; std::vector<std::vector<uint> > durs;
	std::vector<float> finalWts; std::vector<uint> finalStates;
	ObservedDecoder(CRF_FeatureStream* f, CRF_Model* m) : CRF_ViterbiDecoder_StdSeg_NoSegTransFtr<CRF_ViterbiNode>(f, m) {}
	void pruning(uint nodeCnt, double beam) {
		CRF_ViterbiDecoder_StdSeg_NoSegTransFtr<CRF_ViterbiNode>::pruning(nodeCnt, beam);
		if (phnIds.size() <= nodeCnt) { phnIds.resize(nodeCnt + 1); ptrs.resize(nodeCnt + 1); durs.resize(nodeCnt + 1); }
		phnIds[nodeCnt] = this->nodeList->at(nodeCnt)->viterbiPhnIds;
		ptrs[nodeCnt] = this->nodeList->at(nodeCnt)->viterbiPointers;
		durs[nodeCnt] = this->nodeList->at(nodeCnt)->viterbiDurs;
	}
	void expandFinalNode(uint finalNodeCnt, VectorFst<StdArc>* lm, double beam) {
		finalWts = *this->prevViterbiWts_nStates;
		finalStates = *this->prevViterbiStateIds;
		CRF_ViterbiDecoder_StdSeg_NoSegTransFtr<CRF_ViterbiNode>::expandFinalNode(finalNodeCnt, lm, beam);
	}
	uint nStatesPerPhone() { return this->nStates; }
};

}  // namespace

extern "C" {

const char* crfref_last_error() { return g_err.c_str(); }

uint32_t crfref_window_width(const crfref_config* c) { return (uint32_t)window_width(c); }

int crfref_lambda_len(const crfref_config* c, uint32_t* out) {
	try {
		Quiet q;
		ModelBundle mb(c, window_width(c), NULL, 0);
		*out = mb.crf->getLambdaLen();
		return 0;
	} catch (std::exception& e) { g_err = e.what(); return 1; }
}

/* grad is accumulated into (caller zeroes), numer/logZ are per utterance. */
int crfref_fwdbwd_mt2(const crfref_config* c, const double* lambda, uint32_t lambda_len,
                      uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2, const uint32_t* frame_labs,
                      double* grad, double* numer, double* logZ, uint32_t n_threads) {
	try {
		Quiet q;
		ModelBundle mb(c, window_width(c), lambda, lambda_len);
		if (n_threads < 1) n_threads = 1;
		if (n_threads > n_utt) n_threads = n_utt ? n_utt : 1;
		std::vector<ShardArgs> sh(n_threads);
		std::vector<std::vector<double> > sgrad(n_threads);
		size_t per = n_utt / n_threads;
		for (uint32_t i = 0; i < n_threads; i++) {
			size_t start = i * per, cnt = (i == n_threads - 1) ? n_utt - start : per;
			sh[i].cfg = c; sh[i].crf = mb.crf;
			/* offsets stay absolute: the shard's streams index the shared frame arrays */
			sh[i].ftrs = base_ftrs; sh[i].ftrs2 = base_ftrs2; sh[i].labs = frame_labs; sh[i].off = frame_off + start; sh[i].nutt = cnt; sh[i].u0 = start;
			sh[i].numer = numer + start; sh[i].logZ = logZ + start;
			if (n_threads == 1) sh[i].grad = grad;
			else { sgrad[i].assign(lambda_len, 0.0); sh[i].grad = sgrad[i].data(); }
		}
		if (n_threads == 1) run_shard(&sh[0]);
		else {
			std::vector<pthread_t> th(n_threads);
			for (uint32_t i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, shard_thread, &sh[i]);
			for (uint32_t i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
			for (uint32_t i = 0; i < n_threads; i++)
				for (uint32_t k = 0; k < lambda_len; k++) grad[k] += sgrad[i][k];
		}
		for (uint32_t i = 0; i < n_threads; i++)
			if (!sh[i].err.empty()) { g_err = sh[i].err; return 2; }
		return 0;
	} catch (std::exception& e) { g_err = e.what(); return 1; }
}

int crfref_fwdbwd_mt(const crfref_config* c, const double* lambda, uint32_t lambda_len,
                     uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const uint32_t* frame_labs,
                     double* grad, double* numer, double* logZ, uint32_t n_threads) {
	return crfref_fwdbwd_mt2(c, lambda, lambda_len, n_utt, frame_off, base_ftrs, NULL, frame_labs, grad, numer, logZ, n_threads);
}

int crfref_fwdbwd(const crfref_config* c, const double* lambda, uint32_t lambda_len,
                  uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const uint32_t* frame_labs,
                  double* grad, double* numer, double* logZ) {
	return crfref_fwdbwd_mt(c, lambda, lambda_len, n_utt, frame_off, base_ftrs, frame_labs, grad, numer, logZ, 1);
}

/* Best path per utterance as segments.  out_lab/out_dur/out_phn are [sum T] (worst case one
 * segment per frame), segments of utterance u start at frame_off[u]; n_seg[u] segments are valid.
 * out_lab = sub-state label (ilabel-1), out_phn = phone emitted on that arc (olabel-1, or
 * 0xffffffff when none), path_cost = float cost of the winning hypothesis, logZ = final weight. */
/* Phone-bigram language model handed to nStateDecode as its lm_fst (one state per phone only): the topology of the decoder's own
 * free-phone LM (createFreePhoneLmFst, .cpp:1270-1348: state 0 = start, state p + 1 = "the last phone was p", an arc to every OTHER phone,
 * every phone state final) with a weight on every arc -- lm_start[q], lm_bigram[p*P + q] (the diagonal is unused) -- and a final weight
 * lm_final[p].  N states per phone: lm_start = unigram costs, lm_bigram = P exit costs on the epsilon arcs back to the start state.
 * All three NULL: lm_fst == NULL. */
static int viterbi_impl(const crfref_config* c, const double* lambda, uint32_t lambda_len,
                        uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                        const float* lm_start, const float* lm_bigram, const float* lm_final, double beam,
                        uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                        float* path_cost, double* logZ) {
	try {
		Quiet q;
		ModelBundle mb(c, window_width(c), lambda, lambda_len);
		Streams st(c, base_ftrs, base_ftrs2, NULL, frame_off, n_utt);
		VectorFst<StdArc> lm; const bool have_lm = lm_start != NULL;
		if (have_lm) {
			const int P = (int)(c->n_labs / c->n_states);
			int s0 = lm.AddState(); lm.SetStart(s0);
			for (int p = 0; p < P; p++) { int sp = lm.AddState(); lm.AddArc(s0, StdArc(p + 1, p + 1, lm_start[p], sp)); lm.SetFinal(sp, lm_final[p]); }
			if (c->n_states == 1) {
				for (int p = 0; p < P; p++) for (int r = 0; r < P; r++) if (r != p) lm.AddArc(p + 1, StdArc(r + 1, r + 1, lm_bigram[(size_t)p * P + r], r + 1));
			} else {
				/* N states per phone: the free-phone LM's own topology (.cpp:1313-1330) -- every phone state returns to the start state through
				 * an epsilon arc, here with the phone's EXIT cost lm_bigram[p] (P values), and leaves it again on the unigram arcs lm_start[q] */
				for (int p = 0; p < P; p++) lm.AddArc(p + 1, StdArc(0, 0, lm_bigram[p], s0));
			}
		}
		st.fs->rewind();
		for (uint32_t u = 0; u < n_utt; u++) {
			if (st.fs->nextseg() == QN_SEGID_BAD) throw std::runtime_error("stream ended early");
			ObservedDecoder vd(st.fs, mb.crf);   /* one decoder per utterance, CRFDecode/src/Main.cpp:1064-1112 */
			VectorFst<StdArc> best, full;
			int T = vd.nStateDecode(&best, have_lm ? &lm : NULL, &full, beam);
			if ((uint32_t)T != frame_off[u + 1] - frame_off[u]) throw std::runtime_error("decoder frame count mismatch");
			/* walk the reference's own linear best-path FST */
			std::vector<uint32_t> labs, phns; double zx = 0.0;
			int s = best.Start();
			while (s >= 0 && best.NumArcs(s) > 0) {
				const StdArc& a = best.arcs_[s][0];
				labs.push_back((uint32_t)(a.ilabel - 1));
				phns.push_back(a.olabel > 0 ? (uint32_t)(a.olabel - 1) : 0xffffffffu);
				s = a.nextstate;
			}
			if (s >= 0) zx = best.Final(s).Value();
			/* independent traceback over the observed survivors to recover durations
			 * (same rule as CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:2156-2171, 2204-2349) */
			uint nS = vd.nStatesPerPhone();
			float minw = 99999.0; int min_idx = -1;
			if (!have_lm) {
				for (size_t idx = 0; idx * nS < vd.finalWts.size(); idx++) {
					int e = (int)(idx * nS + nS - 1);
					if (vd.finalWts[e] < minw) { minw = vd.finalWts[e]; min_idx = e; }
				}
			} else {
				/* with an input LM the decoder takes the minimum over its finalStateSet -- ordered by LM state id, weight = hypothesis weight
				 * + final weight of its LM state (expandFinalNode :746-758, addToFinalSet :929-946, selection :2138-2153) */
				for (int sid = 0; sid < lm.NumStates(); sid++)
					for (size_t idx = 0; idx < vd.finalStates.size(); idx++) {
						if ((int)vd.finalStates[idx] != sid || lm.Final(sid) == TropicalWeight::Zero()) continue;
						int e = (int)(idx * nS + nS - 1);
						float w = vd.finalWts[e] + lm.Final(sid).Value();
						if (w < minw) { minw = w; min_idx = e; }
					}
			}
			std::vector<uint32_t> tl, td;
			int end = T - 1;
			while (min_idx >= 0 && end >= 0) {
				uint lab = (vd.phnIds[end][min_idx / nS] - 1) * nS + (min_idx % nS);
				uint d = vd.durs[end][min_idx];
				tl.push_back(lab); td.push_back(d);
				int startf = end + 1 - (int)d;
				if (startf > 0) min_idx = vd.ptrs[end][min_idx];
				end = startf - 1;
			}
			if (min_idx < 0 && tl.empty()) {
				/* "Could not reach end of utterance" (.cpp:2188-2195): the reference emits one dummy arc */
				n_seg[u] = 0; path_cost[u] = minw; logZ[u] = zx;
				continue;
			}
			if (tl.size() != labs.size()) throw std::runtime_error("traceback/FST segment count mismatch");
			uint32_t base = frame_off[u];
			n_seg[u] = (uint32_t)labs.size();
			for (size_t k = 0; k < labs.size(); k++) {
				size_t r = labs.size() - 1 - k;   /* tl is end-to-start */
				if (tl[r] != labs[k]) throw std::runtime_error("traceback/FST label mismatch");
				out_lab[base + k] = labs[k]; out_dur[base + k] = td[r]; out_phn[base + k] = phns[k];
			}
			path_cost[u] = minw; logZ[u] = zx;
		}
		return 0;
	} catch (std::exception& e) { g_err = e.what(); return 1; }
}

int crfref_viterbi(const crfref_config* c, const double* lambda, uint32_t lambda_len,
                   uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs,
                   uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                   float* path_cost, double* logZ) {
	return viterbi_impl(c, lambda, lambda_len, n_utt, frame_off, base_ftrs, NULL, NULL, NULL, NULL, 0.0, out_lab, out_dur, out_phn, n_seg, path_cost, logZ);
}

int crfref_viterbi2(const crfref_config* c, const double* lambda, uint32_t lambda_len,
                    uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                    uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                    float* path_cost, double* logZ) {
	return viterbi_impl(c, lambda, lambda_len, n_utt, frame_off, base_ftrs, base_ftrs2, NULL, NULL, NULL, 0.0, out_lab, out_dur, out_phn, n_seg, path_cost, logZ);
}

int crfref_viterbi_lm(const crfref_config* c, const double* lambda, uint32_t lambda_len,
                      uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                      const float* lm_start, const float* lm_bigram, const float* lm_final,
                      uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                      float* path_cost, double* logZ) {
	return viterbi_impl(c, lambda, lambda_len, n_utt, frame_off, base_ftrs, base_ftrs2, lm_start, lm_bigram, lm_final, 0.0, out_lab, out_dur, out_phn, n_seg, path_cost, logZ);
}

int crfref_viterbi_beam(const crfref_config* c, const double* lambda, uint32_t lambda_len,
                        uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, const float* base_ftrs2,
                        const float* lm_start, const float* lm_bigram, const float* lm_final, double beam,
                        uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                        float* path_cost, double* logZ) {
	return viterbi_impl(c, lambda, lambda_len, n_utt, frame_off, base_ftrs, base_ftrs2, lm_start, lm_bigram, lm_final, beam, out_lab, out_dur, out_phn, n_seg, path_cost, logZ);
}

/* Older frame-level decoder; one label per frame. */
int crfref_viterbi_old(const crfref_config* c, const double* lambda, uint32_t lambda_len,
                       uint32_t n_utt, const uint32_t* frame_off, const float* base_ftrs, uint32_t* out_lab) {
	try {
		Quiet q;
		crfref_config cc = *c; cc.model_type = STDFRAME;
		ModelBundle mb(&cc, window_width(&cc), lambda, lambda_len);
		Streams st(&cc, base_ftrs, NULL, NULL, frame_off, n_utt);
		st.fs->rewind();
		for (uint32_t u = 0; u < n_utt; u++) {
			if (st.fs->nextseg() == QN_SEGID_BAD) throw std::runtime_error("stream ended early");
			CRF_ViterbiDecoder vd(st.fs, mb.crf);
			VectorFst<StdArc> best;
			vd.nStateDecode(&best, NULL, 0.0);
			uint32_t k = frame_off[u];
			int s = best.Start();
			while (s >= 0 && best.NumArcs(s) > 0 && k < frame_off[u + 1]) {
				const StdArc& a = best.arcs_[s][0];
				out_lab[k++] = (uint32_t)(a.ilabel - 1);
				s = a.nextstate;
			}
			if (k != frame_off[u + 1]) throw std::runtime_error("old decoder: arc count != frame count");
		}
		return 0;
	} catch (std::exception& e) { g_err = e.what(); return 1; }
}

/* Expanded window features exactly as buildGradient receives them: for frame t the windows of
 * duration 1..min(t+1,max_dur) ending at t, each window_width floats, written at
 * out[(frame_off_in_utt*max_dur + (d-1)) * width]; unused (t,d) slots are left untouched. */
int crfref_window_ftrs2(const crfref_config* c, uint32_t n_frames, const float* base_ftrs, const float* base_ftrs2, float* out) {
	try {
		Quiet q;
		uint32_t off[2] = {0, n_frames};
		Streams st(c, base_ftrs, base_ftrs2, NULL, off, 1);
		size_t w = st.fs->num_ftrs();
		if (w != window_width(c)) throw std::runtime_error("window width formula disagrees with the reference streams");
		if (st.fs->nextseg() == QN_SEGID_BAD) throw std::runtime_error("no segment");
		std::vector<float> buf(w * c->max_dur);
		size_t bunch = 1, t = 0;
		for (;;) {
			size_t n = st.fs->read(bunch, buf.data(), NULL);
			if (n == 0) break;
			memcpy(out + (size_t)t * c->max_dur * w, buf.data(), n * w * sizeof(float));
			t++;
			if (bunch < c->max_dur) bunch++;
		}
		if (t != n_frames) throw std::runtime_error("window stream frame count mismatch");
		return 0;
	} catch (std::exception& e) { g_err = e.what(); return 1; }
}
int crfref_window_ftrs(const crfref_config* c, uint32_t n_frames, const float* base_ftrs, float* out) { return crfref_window_ftrs2(c, n_frames, base_ftrs, NULL, out); }

/* Per-frame 4-word label records (lab,start,end,broken) or CRF_LAB_BAD x4. */
int crfref_window_labs(const crfref_config* c, uint32_t n_frames, const uint32_t* frame_labs, uint32_t* out4) {
	try {
		Quiet q;
		uint32_t off[2] = {0, n_frames};
		MemLabStream memL(frame_labs, off, 1);
		CRF_InLabStream_SeqMultiWindow winL(0, "l", memL, c->max_dur, 0, 0, 65536, 65536);
		if (winL.nextseg() == QN_SEGID_BAD) throw std::runtime_error("no segment");
		for (uint32_t t = 0; t < n_frames; t++)
			if (winL.read_labs(1, out4 + 4 * (size_t)t) != 1) throw std::runtime_error("label stream ended early");
		return 0;
	} catch (std::exception& e) { g_err = e.what(); return 1; }
}

}  // extern "C"
