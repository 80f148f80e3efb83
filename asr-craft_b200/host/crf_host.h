// crf_host.h -- C++ host layer above the C ABI (include/crfgpu.h), mirroring the plug-in surface of ASR-CRaFT for
// the lattice hot path: same class and member names, argument meaning and error behaviour (std::runtime_error), so
// that a reference user finds the seams where they expect them.  Reference interfaces (ASR-CRaFT tree):
//   CRF_FeatureMap_config                      CRF/src/ftrmaps/CRF_FeatureMap.h:24-47
//   CRF_Model                                  CRF/src/CRF_Model.h:23-84, CRF_Model.cpp:75-87,205-235,290-308
//   CRF_FeatureStream (nextseg/read/rewind)    CRF/src/io/CRF_FeatureStream.h:35-67, .cpp:71-136
//   CRF_GradBuilder::create / buildGradient    CRF/src/trainers/gradbuilders/CRF_GradBuilder.h:40-44, .cpp:97-162
//   CRF_Minibatch_GradAccumulator              CRF/src/trainers/accumulators/CRF_Minibatch_GradAccumulator.h:51-75, .cpp:201-322
//   CRF_ViterbiDecoder_StdSeg_NoSegTransFtr    CRF/src/decoders/CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:1369-2398
//   CRF_SGTrainer (train loop, update, files)  CRF/src/trainers/CRF_SGTrainer.cpp:99-430, CRF_Trainer.h:35-42
// Differences that follow from the device design (documented in INTEGRATION.md):
//   * the feature stream handed to the builders is the UN-windowed one (what the pfile / ilab hold): one frame of
//     num_ftrs() floats and one phone label per read(); the segment windows are expanded on the device;
//   * OpenFst is not available here: the decoder returns the best path as a vector of arcs carrying exactly the fields
//     the reference puts on its linear FST (ilabel = sub-state label + 1, olabel = phone + 1 or 0) plus the duration.
#ifndef CRF_HOST_H
#define CRF_HOST_H

#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/crfgpu.h"

typedef uint32_t QNUInt32;
typedef long QN_SegID;
enum { QN_SEGID_BAD = -1 };
#define CRF_LAB_BAD 0xffffffffu

enum ftrmaptype { STDSTATE, STDTRANS, STDSPARSE, STDSPARSETRANS, INFILE };
enum objfunctype { EXPF, EXPFSOFT, FERR };
enum modeltype { STDFRAME, STDSEG, STDSEG_NO_DUR, STDSEG_NO_DUR_NO_TRANSFTR, STDSEG_NO_DUR_NO_SEGTRANSFTR };

struct CRF_FeatureMap_config {
	ftrmaptype map_type;
	QNUInt32 numLabs;
	QNUInt32 numFeas;          // width of the WINDOW feature vector (state/trans index ranges refer to it)
	QNUInt32 numStates;
	bool useStateFtrs;
	QNUInt32 stateFidxStart;
	QNUInt32 stateFidxEnd;
	bool useTransFtrs;
	QNUInt32 transFidxStart;
	QNUInt32 transFidxEnd;
	bool useStateBias;
	bool useTransBias;
	double stateBiasVal;
	double transBiasVal;
	QNUInt32 maxDur;
	QNUInt32 durFtrStart;
	QNUInt32 nActualLabs;
};

// Un-windowed frame stream: nextseg() advances to the next utterance (QN_SEGID_BAD at the end), read() returns up to `bs`
// frames (num_ftrs() floats and one label word each) of the current utterance, 0 when it is exhausted.
class CRF_FeatureStream {
public:
	virtual ~CRF_FeatureStream() {}
	virtual QN_SegID nextseg() = 0;
	virtual size_t read(size_t bs, float* fb, QNUInt32* lb) = 0;
	virtual void rewind() = 0;
	virtual QNUInt32 num_ftrs() = 0;
	virtual QNUInt32 num_labs() { return 1; }
	virtual QNUInt32 num_segs() = 0;
	// A new stream over the contiguous utterance range [startseg, startseg + nsegs) of this one -- what
	// CRF_FeatureStreamManager::getChild(i)->trn_stream->view(...) hands every training thread (CRF_FeatureStreamManager.cpp:425-464).
	// The caller owns the result; nullptr = this stream cannot be split (then only nStreams == 1 is possible).
	virtual CRF_FeatureStream* newView(QNUInt32 /*startseg*/, QNUInt32 /*nsegs*/) { return nullptr; }
	// CRF_FeatureStream::join (CRF/src/io/CRF_FeatureStream.cpp:172-184): a new stream whose frames are this stream's followed by
	// in_stream's ("the streams must match segment-wise and frame-wise"; here in_stream may carry the left + right context frames the
	// model's second window stream is configured with, CRF_Model::setSecondStream).  The result reads like this stream (features,
	// labels, segments) and hands the second one out through joinedStream(); it advances and rewinds both.  The caller owns it.
	virtual CRF_FeatureStream* join(CRF_FeatureStream* in_stream);
	virtual CRF_FeatureStream* joinedStream() { return nullptr; }
};

class CRF_JoinedFeatureStream : public CRF_FeatureStream {
	CRF_FeatureStream* a; CRF_FeatureStream* b; bool own;
public:
	CRF_JoinedFeatureStream(CRF_FeatureStream* first, CRF_FeatureStream* second, bool owns = false) : a(first), b(second), own(owns) {}
	~CRF_JoinedFeatureStream() { if (own) { delete a; delete b; } }
	QN_SegID nextseg() { const QN_SegID x = a->nextseg(), y = b->nextseg(); if ((x == QN_SEGID_BAD) != (y == QN_SEGID_BAD)) throw std::runtime_error("joined feature streams disagree on the number of segments"); return x; }
	size_t read(size_t bs, float* fb, QNUInt32* lb) { return a->read(bs, fb, lb); }
	void rewind() { a->rewind(); b->rewind(); }
	QNUInt32 num_ftrs() { return a->num_ftrs(); }
	QNUInt32 num_segs() { return a->num_segs(); }
	CRF_FeatureStream* newView(QNUInt32 startseg, QNUInt32 nsegs) {
		CRF_FeatureStream* va = a->newView(startseg, nsegs); CRF_FeatureStream* vb = b->newView(startseg, nsegs);
		if (!va || !vb) { delete va; delete vb; return nullptr; }
		return new CRF_JoinedFeatureStream(va, vb, true);
	}
	CRF_FeatureStream* joinedStream() { return b; }
};

// In-memory ragged batch (what tests and the self-test feed); view(start, n) restricts it to a contiguous utterance range,
// the rule CRF_FeatureStreamManager uses to shard a corpus over streams (CRF_FeatureStreamManager.cpp:425-464).
class CRF_MemFeatureStream : public CRF_FeatureStream {
	struct Data { std::vector<uint32_t> off; std::vector<float> ftrs; std::vector<QNUInt32> labs; };
	std::shared_ptr<const Data> d; QNUInt32 nf;
	QNUInt32 first, count; long seg; uint32_t pos;
public:
	CRF_MemFeatureStream(const std::vector<uint32_t>& frame_off, const std::vector<float>& f, const std::vector<QNUInt32>& l, QNUInt32 n_ftrs);
	QN_SegID nextseg();
	size_t read(size_t bs, float* fb, QNUInt32* lb);
	void rewind();
	QNUInt32 num_ftrs() { return nf; }
	QNUInt32 num_segs() { return count; }
	void view(QNUInt32 startseg, QNUInt32 nsegs);
	CRF_FeatureStream* newView(QNUInt32 startseg, QNUInt32 nsegs);     // shares the data, own position
};

class CRF_Model {
protected:
	QNUInt32 nlabs;
	std::vector<double> lambda, lambdaAcc, gradSqrAcc;
	QNUInt32 lab_max_dur, nActualLabs;
	modeltype model_type;
	CRF_FeatureMap_config fmap;
	bool have_map;
	QNUInt32 n_base_ftrs; bool extract_seg_ftrs;
	QNUInt32 n_base_ftrs2, left_ctx2, right_ctx2; bool extract_seg_ftrs2, boundary_delta2;     // second window stream (0 features: none)
	std::vector<crfgpu_handle> handles;       // one device context per GPU (handles[0] created by setFeatureMap)
	crfgpu_config dev_cfg;
	QNUInt32 init_present, init_iter;         // resume state (CRF_Model.cpp:29,323,487-507)
public:
	CRF_Model(QNUInt32 num_labs);
	virtual ~CRF_Model();
	QNUInt32 getNLabs() { return nlabs; }
	// createFeatureMap(cfg) + setFeatureMap: fixes the lambda layout and allocates lambda (CRF_Model.cpp:75-87).
	// n_base_ftrs / extract_seg_ftrs are the window-stream options of CRFTrain (ftr1_window_len == maxDur, ftr1_extract_seg_ftr).
	virtual void setFeatureMap(const CRF_FeatureMap_config& cfg, QNUInt32 n_base_ftrs, bool extract_seg_ftrs, int device = 0);
	virtual double* getLambda() { return lambda.data(); }
	virtual QNUInt32 getLambdaLen() { return (QNUInt32)lambda.size(); }
	virtual double* getLambdaAcc() { return lambdaAcc.data(); }
	virtual double* getGradSqrAcc() { return gradSqrAcc.data(); }      // AdaGrad accumulator (CRF_Model.h:42)
	virtual void setLambda(double* lam, QNUInt32 len);
	virtual void resetLambda();
	virtual bool writeToFile(const char* fname);                       // ASCII, one value per line, default ostream precision
	virtual bool writeToFile(const char* fname, double* lam, QNUInt32 ll);   // any vector in the same format (CRF_Model.cpp:248-279)
	virtual bool readFromFile(const char* fname);
	// resume (CRF_Model.cpp:323-372, CRFTrain/src/Main.cpp:611-630): the average file times `present` refills lambdaAcc
	virtual bool readAverageFromFile(const char* fname, int present);
	virtual bool readGradSqrAccFromFile(const char* fname);
	virtual QNUInt32 getPresentations() { return init_present; }
	// resume from the state a finished train() left in this object (lambda, lambdaAcc, gradSqrAcc) instead of from the lossy ASCII files
	void resumeFromMemory(QNUInt32 presentations, QNUInt32 start_iter) { init_present = presentations; init_iter = start_iter; }
	virtual void setInitIter(QNUInt32 start_iter) { init_iter = start_iter; }
	virtual QNUInt32 getInitIter() { return init_iter; }
	virtual void setLabMaxDur(QNUInt32 d) { lab_max_dur = d; }
	virtual QNUInt32 getLabMaxDur() { return lab_max_dur; }
	// The second feature stream of CRFTrain / CRFDecode (ftr2_file with ftr2_window_len == maxDur, ftr2_left_context_len,
	// ftr2_right_context_len, ftr2_extract_seg_ftr, ftr2_use_boundary_delta_ftr; CRFTrain/src/Main.cpp:516-526), joined behind the first:
	// call BEFORE setFeatureMap, like setLabMaxDur.  The streams handed to the builders / decoder must then be joined ones
	// (CRF_FeatureStream::join) whose second stream carries left_ctx + right_ctx more frames per utterance (the padded pfile).
	virtual void setSecondStream(QNUInt32 n_base_ftrs2_in, bool extract_seg_ftrs2_in, QNUInt32 left_ctx, QNUInt32 right_ctx, bool boundary_delta = false) {
		n_base_ftrs2 = n_base_ftrs2_in; extract_seg_ftrs2 = extract_seg_ftrs2_in; left_ctx2 = left_ctx; right_ctx2 = right_ctx; boundary_delta2 = boundary_delta;
	}
	QNUInt32 secondStreamFtrs() { return n_base_ftrs2; }
	QNUInt32 numStates() { return have_map ? fmap.numStates : 1; }
	virtual void setNActualLabs(QNUInt32 n) { nActualLabs = n; }
	virtual QNUInt32 getNActualLabs() { return nActualLabs; }
	virtual void setModelType(modeltype m) { model_type = m; }
	virtual modeltype getModelType() { return model_type; }
	crfgpu_handle gpu(size_t i = 0);                                  // the device context(s) of this model (the first is created by setFeatureMap)
	// Data-parallel training over several GPUs of one box from ONE process: one more device context per listed device and an NCCL
	// communicator over all of them (crfgpu_comm_init_all).  The accumulator then maps stream s to device s % nDevices().
	void addDevices(const std::vector<int>& devices);
	size_t nDevices() { return handles.size(); }
	void pushLambdaToDevices();                                       // crfgpu_set_lambda on every device context
	// lambda lives on the device between minibatches (CRF_SGTrainer below): accumulateGradient then skips the upload and the
	// host copy is refreshed by syncLambdaFromDevice() (end of an iteration, before a checkpoint)
	bool lambdaOnDevice; void syncLambdaFromDevice();
	QNUInt32 baseFtrs() { return n_base_ftrs; }
};

// per-utterance seam, kept for API parity: grad is ACCUMULATED into, returns the numerator, *Zx_out = logZ
class CRF_GradBuilder {
protected:
	CRF_Model* crf;
	std::vector<float> ftr_buf, ftr2_buf; std::vector<QNUInt32> lab_buf; std::vector<double> tmp_grad;
public:
	CRF_GradBuilder(CRF_Model* crf_in) : crf(crf_in) {}
	virtual ~CRF_GradBuilder() {}
	virtual double buildGradient(CRF_FeatureStream* ftr_strm, double* grad, double* Zx_out);
	static CRF_GradBuilder* create(CRF_Model* crf_ptr, objfunctype ofunc);
};

// minibatch seam (CRF_Minibatch_GradAccumulator.cpp:112-322).  nStreams contiguous views of the corpus, one per training thread in the
// reference; per call stream i contributes floor(minibatch / nStreams) (+1 for the first minibatch % nStreams streams) utterances -- at
// least one, a stream stops early at the end of its view -- and the summed gradient is divided by the number of streams that were
// NOT yet exhausted when the call started.  Here stream s is computed on device s % nDevices (all streams of one device form ONE
// device batch), and with several devices the per-device sums are combined by ONE NCCL all-reduce (crfgpu_allreduce_grad).
#define CRF_UINT32_MAX 0xffffffffu
class CRF_Minibatch_GradAccumulator {
protected:
	CRF_Model* crf; QNUInt32 nStreams, minibatch;
	std::vector<CRF_FeatureStream*> ftrStrms; std::vector<bool> owned; std::vector<QN_SegID> strmsSegids;
	struct DevBatch { std::vector<uint32_t> off; std::vector<float> ftrs, ftrs2; std::vector<QNUInt32> labs; };
	std::vector<DevBatch> dev;
	double runBatch(double* grad, double* Zx_out, QNUInt32* uttCount, bool* isEndOfIter, QNUInt32* nActive);
public:
	// `stream` is the whole training corpus; for myNStreams > 1 it is split with newView() by the reference's view rule
	CRF_Minibatch_GradAccumulator(CRF_Model* myCrf, CRF_FeatureStream* stream, QNUInt32 myNStreams);
	virtual ~CRF_Minibatch_GradAccumulator();
	// grad is overwritten with sum(emp - exp) / nStreams_active, *Zx_out = sum logZ, returns sum of numerators
	virtual double accumulateGradient(double* grad, double* Zx_out, QNUInt32* uttCount, bool* isEndOfIter);
	// minibatch < nStreams is an error (the reference prints and exits, .cpp:175-181); the default is the whole view per call
	void setMinibatch(QNUInt32 mb);
	// the same batch, but the (all-reduced) gradient stays on the device(s) for crfgpu_sgd_update (no download, no division)
	virtual double accumulateGradientOnDevice(double* Zx_out, QNUInt32* uttCount, bool* isEndOfIter, QNUInt32* nActive);
	QNUInt32 getNStreams() { return nStreams; }
	void rewindAllAndNextSegs();
	// utterances each stream takes in the next call and the divisor, from the streams' end flags (pure host rule, exposed for tests)
	static QNUInt32 planShares(QNUInt32 minibatch, QNUInt32 nStreams, const std::vector<bool>& atEnd, std::vector<QNUInt32>* share);
};

struct CRF_BestPathArc {
	int ilabel;      // sub-state label + 1
	int olabel;      // phone + 1 when the segment starts a phone, else 0 (epsilon)
	QNUInt32 dur;    // frames covered (the reference keeps this in viterbiDurs)
};

// The language models the device decodes against (nStateDecode's lm_fst): complete phone-bigram LMs in the topology of the decoder's own
// free-phone LM (createFreePhoneLmFst, CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:1270-1348), one state per phone.  Costs = the arc
// weights an OpenFst LM of that shape carries: start[q] on start -> q, bigram[p*P + q] on p -> q (p != q), final_wt[p] (+inf: not final).
// With N > 1 states per phone the reference's free-phone LM returns from every phone state to the start state through an epsilon arc:
// start[q] = unigram cost of the arc start -> q, bigram[p] (P values) = exit cost of phone p's epsilon arc (a phone insertion penalty).
struct CRF_PhoneBigramLm {
	std::vector<float> start, bigram, final_wt;
};

class CRF_ViterbiDecoder_StdSeg_NoSegTransFtr {
	CRF_FeatureStream* strm; CRF_Model* crf;
public:
	CRF_ViterbiDecoder_StdSeg_NoSegTransFtr(CRF_FeatureStream* ftr_strm_in, CRF_Model* crf_in) : strm(ftr_strm_in), crf(crf_in) {}
	// decodes the CURRENT utterance of the stream (free-phone LM); returns the number of frames, like nStateDecode.  beam > 0: the
	// reference's beam pruning (one state per phone)
	int nStateDecode(std::vector<CRF_BestPathArc>* result, float* path_cost, double beam = 0.0);
	// the same against a language model (the reference's nStateDecode(result, lm_fst, ...), .cpp:1369-1372); path_cost includes the final weight
	int nStateDecode(std::vector<CRF_BestPathArc>* result, const CRF_PhoneBigramLm* lm_fst, float* path_cost, double beam = 0.0);
	// the LM later nStateDecode / nStateDecodeBatch calls decode against (nullptr: the free-phone LM)
	void setLm(const CRF_PhoneBigramLm* lm_fst);
	// The decode loop of CRFDecode (Main.cpp:1064-1112: one decoder object per utterance) as ONE device batch: decodes up to max_utts
	// utterances from the current one on (advancing the stream with nextseg()), lambda uploaded once; returns the utterances decoded.
	// results[u] / path_costs[u] / n_frames[u] are what nStateDecode returns for utterance u.
	// *stream_end is set when the stream has no further utterance (nextseg() returned QN_SEGID_BAD).
	size_t nStateDecodeBatch(size_t max_utts, std::vector<std::vector<CRF_BestPathArc>>* results, std::vector<float>* path_costs, std::vector<int>* n_frames,
	                         bool* stream_end = nullptr, double beam = 0.0);
};

// CRF_SGTrainer::train() (CRF/src/trainers/CRF_SGTrainer.cpp:99-430) over the device path: per minibatch one device batch and one
// crfgpu_sgd_update (lambda += lr * grad / nActiveStreams, lambdaAcc += lambda, optional AdaGrad / gvar), at the end of every iteration
// the reference's files -- <weights>.i<k>.out, <weights>.i<k>.avg.out (lambdaAcc / (float)accCnt), <weights>.done.train.i<k> -- and the
// learning-rate decay; finally <weights> and <weights>.avg.out.  lr / lr_decay_rate are floats promoted in the update (CRF_Trainer.h:35-42).
// Iterations are numbered from CRF_Model::getInitIter() (0 unless resuming) while < maxIters (CRF_SGTrainer.cpp:86,203); with AdaGrad
// every iteration also writes <weights>.i<k>.gradSqrAcc.out (:370-381); the done markers are <dir of weights>/.done.train.i<k> and
// <dir>/.done.train (CRF_Trainer.cpp:144-179).  Resume = readFromFile + readAverageFromFile(avg, presentations) [+ readGradSqrAccFromFile]
// + setInitIter(k) on the model before train(), exactly the CRFTrain options init_weight_file / avg_weight_file / avg_weight_present /
// grad_sqr_acc_file / init_iter (CRFTrain/src/Main.cpp:599-630).
class CRF_SGTrainer {
	CRF_Model* crf_ptr; CRF_FeatureStream* strm; std::string weight_fname;
	float lr, lr_decay_rate; int maxIters; QNUInt32 minibatch, nStreams; bool useAdagrad; double eta, eps; bool useGvar; float gvar;
public:
	CRF_SGTrainer(CRF_Model* crf_in, CRF_FeatureStream* stream, const char* wt_fname);
	void setLR(float v) { lr = v; }
	void setLRDecayRate(float v) { lr_decay_rate = v; }
	void setMaxIters(int n) { maxIters = n; }
	void setMinibatch(QNUInt32 mb, QNUInt32 n_streams) { minibatch = mb; nStreams = n_streams; }
	void setAdagrad(bool on, double eta_in, double eps_in) { useAdagrad = on; eta = eta_in; eps = eps_in; }
	void setGaussVar(float gvar_in) { gvar = gvar_in; if (gvar != 0.0) useGvar = true; }     // CRF_Trainer.cpp:106-111; the update uses float 1/gvar (CRF_SGTrainer.cpp:164-167)
	int presentations;                    // utterances presented so far (accCnt): what a resumed run passes as avg_weight_present
	std::vector<double> iterLogLi;        // per iteration: sum over the corpus of numerator - logZ (what the trainer prints as Iter-Avg LogLi * utts)
	void train();
};

#endif
