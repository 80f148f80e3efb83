// crf_host.cpp -- see crf_host.h.  Thin: every arithmetic step happens behind the C ABI on the device.
#include "crf_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

using std::runtime_error;
using std::string;

static void check(int rc, const char* what) {
	if (rc != CRFGPU_OK) throw runtime_error(string(what) + ": " + crfgpu_last_error());   // the reference throws; main() catches and exits
}

// ---------------------------------------------------------------------------------------------- stream
CRF_MemFeatureStream::CRF_MemFeatureStream(const std::vector<uint32_t>& frame_off, const std::vector<float>& f,
                                           const std::vector<QNUInt32>& l, QNUInt32 n_ftrs)
    : nf(n_ftrs), first(0), count((QNUInt32)frame_off.size() - 1), seg(-1), pos(0) {
	auto data = std::make_shared<Data>();
	data->off = frame_off; data->ftrs = f; data->labs = l;
	d = data;
}

void CRF_MemFeatureStream::view(QNUInt32 startseg, QNUInt32 nsegs) { first = startseg; count = nsegs; rewind(); }
CRF_FeatureStream* CRF_MemFeatureStream::newView(QNUInt32 startseg, QNUInt32 nsegs) {
	CRF_MemFeatureStream* v = new CRF_MemFeatureStream(*this);      // shares the data block
	v->view(first + startseg, nsegs);
	return v;
}
void CRF_MemFeatureStream::rewind() { seg = -1; pos = 0; }
QN_SegID CRF_MemFeatureStream::nextseg() {
	if (seg + 1 >= (long)count) { seg = count; return QN_SEGID_BAD; }
	seg++; pos = d->off[first + seg];
	return seg;
}
size_t CRF_MemFeatureStream::read(size_t bs, float* fb, QNUInt32* lb) {
	if (seg < 0 || seg >= (long)count) return 0;
	const uint32_t end = d->off[first + seg + 1];
	const size_t n = std::min<size_t>(bs, end - pos);
	if (n) {
		std::memcpy(fb, &d->ftrs[(size_t)pos * nf], n * nf * sizeof(float));
		if (lb) for (size_t i = 0; i < n; i++) lb[i] = d->labs.empty() ? CRF_LAB_BAD : d->labs[pos + i];
	}
	pos += (uint32_t)n;
	return n;
}

// ---------------------------------------------------------------------------------------------- model
CRF_Model::CRF_Model(QNUInt32 num_labs)
    : nlabs(num_labs), lab_max_dur(1), nActualLabs(num_labs), model_type(STDFRAME), have_map(false), n_base_ftrs(0),
      extract_seg_ftrs(false), n_base_ftrs2(0), left_ctx2(0), right_ctx2(0), extract_seg_ftrs2(false), boundary_delta2(false),
      init_present(0), init_iter(0), lambdaOnDevice(false) {}
CRF_Model::~CRF_Model() { for (crfgpu_handle h : handles) crfgpu_destroy(h); }

void CRF_Model::setFeatureMap(const CRF_FeatureMap_config& cfg, QNUInt32 base_ftrs, bool seg_ftrs, int device) {
	if (cfg.map_type != STDSTATE && cfg.map_type != STDTRANS) throw runtime_error("only the dense feature maps are implemented on the device (stdstate, stdtrans)");
	fmap = cfg; have_map = true; n_base_ftrs = base_ftrs; extract_seg_ftrs = seg_ftrs;
	crfgpu_config& c = dev_cfg;
	std::memset(&c, 0, sizeof(c));
	c.model_type = (uint32_t)model_type; c.n_labs = cfg.numLabs; c.n_base_ftrs = base_ftrs; c.n_states = cfg.numStates;
	c.max_dur = lab_max_dur; c.n_actual_labs = nActualLabs; c.extract_seg_ftrs = seg_ftrs ? 1 : 0;
	c.use_state_ftrs = cfg.useStateFtrs; c.state_fidx_start = cfg.stateFidxStart; c.state_fidx_end = cfg.stateFidxEnd;
	c.use_trans_ftrs = cfg.useTransFtrs; c.trans_fidx_start = cfg.transFidxStart; c.trans_fidx_end = cfg.transFidxEnd;
	c.use_state_bias = cfg.useStateBias; c.use_trans_bias = cfg.useTransBias;
	c.state_bias_val = cfg.stateBiasVal; c.trans_bias_val = cfg.transBiasVal;
	c.n_base_ftrs2 = n_base_ftrs2; c.extract_seg_ftrs2 = extract_seg_ftrs2 ? 1 : 0; c.left_ctx2 = left_ctx2; c.right_ctx2 = right_ctx2;
	c.boundary_delta2 = boundary_delta2 ? 1 : 0;
	if (crfgpu_window_width(&c) != cfg.numFeas)
		throw runtime_error("CRF_FeatureMap_config::numFeas does not match the window width of the feature stream");
	for (crfgpu_handle h : handles) crfgpu_destroy(h);
	handles.clear();
	crfgpu_handle h = nullptr;
	check(crfgpu_create(&c, device, &h), "crfgpu_create");
	handles.push_back(h);
	lambda.assign(crfgpu_lambda_len(h), 0.0);
	lambdaAcc.assign(lambda.size(), 0.0);
	gradSqrAcc.assign(lambda.size(), 0.0);
}
void CRF_Model::addDevices(const std::vector<int>& devices) {
	if (handles.empty()) throw runtime_error("CRF_Model: setFeatureMap has not been called");
	for (int dv : devices) {
		crfgpu_handle h = nullptr;
		check(crfgpu_create(&dev_cfg, dv, &h), "crfgpu_create");
		handles.push_back(h);
	}
	if (handles.size() > 1) check(crfgpu_comm_init_all(handles.data(), (int)handles.size()), "crfgpu_comm_init_all");
}
crfgpu_handle CRF_Model::gpu(size_t i) {
	if (i >= handles.size()) throw runtime_error("CRF_Model: setFeatureMap has not been called");
	return handles[i];
}
void CRF_Model::pushLambdaToDevices() {
	for (crfgpu_handle h : handles) check(crfgpu_set_lambda(h, lambda.data(), (uint32_t)lambda.size()), "crfgpu_set_lambda");
}
void CRF_Model::setLambda(double* lam, QNUInt32 len) {
	if (len != lambda.size()) throw runtime_error("CRF_Model::setLambda: length mismatch");
	std::copy(lam, lam + len, lambda.begin());
}
void CRF_Model::resetLambda() { std::fill(lambda.begin(), lambda.end(), 0.0); }
bool CRF_Model::writeToFile(const char* fname, double* lam, QNUInt32 ll) {
	std::ofstream ofile(fname);
	if (!ofile.is_open()) throw runtime_error(string("CRF_Model::writeToFile() caught exception: cannot open the file:\n") + fname);   // CRF_Model.cpp:270-276
	for (QNUInt32 i = 0; i < ll; i++) ofile << lam[i] << std::endl;      // default ostream precision: 6 significant digits (CRF_Model.cpp:253-255)
	if (ofile.bad()) throw runtime_error(string("CRF_Model::writeToFile() caught exception: errors when writing the weights to the file:\n") + fname);
	return true;
}
bool CRF_Model::writeToFile(const char* fname) { return writeToFile(fname, lambda.data(), (QNUInt32)lambda.size()); }
static bool read_values(const char* fname, std::vector<double>& v) {
	std::ifstream ifile(fname);
	if (!ifile.is_open()) return false;
	for (size_t i = 0; i < v.size(); i++) {      // one value per line; a short file leaves the rest untouched, like the reference's getline loop
		string s;
		std::getline(ifile, s);
		std::istringstream iss(s);
		iss >> std::dec >> v[i];
	}
	return true;
}
bool CRF_Model::readFromFile(const char* fname) { return read_values(fname, lambda); }
bool CRF_Model::readAverageFromFile(const char* fname, int present) {
	init_present = (QNUInt32)present;
	if (!read_values(fname, lambdaAcc)) return false;
	if (present > 0) for (double& v : lambdaAcc) v = v * present;
	return true;
}
bool CRF_Model::readGradSqrAccFromFile(const char* fname) { return read_values(fname, gradSqrAcc); }

// ---------------------------------------------------------------------------------------------- training
CRF_GradBuilder* CRF_GradBuilder::create(CRF_Model* crf_ptr, objfunctype ofunc) {
	if (ofunc != EXPF) throw runtime_error("only the EXPF objective is implemented (the reference's factory never selects the others either)");
	return new CRF_GradBuilder(crf_ptr);
}

// all frames of the current utterance of `s` appended to the batch; returns the frame count
static size_t read_utterance(CRF_FeatureStream* s, std::vector<float>& ftrs, std::vector<QNUInt32>* labs) {
	const QNUInt32 nf = s->num_ftrs();
	const size_t chunk = 256;
	size_t T = 0;
	for (;;) {
		const size_t f0 = ftrs.size(), l0 = labs ? labs->size() : 0;
		ftrs.resize(f0 + chunk * nf);
		if (labs) labs->resize(l0 + chunk);
		const size_t n = s->read(chunk, ftrs.data() + f0, labs ? labs->data() + l0 : nullptr);
		ftrs.resize(f0 + n * nf);
		if (labs) labs->resize(l0 + n);
		T += n;
		if (n < chunk) break;
	}
	return T;
}

CRF_FeatureStream* CRF_FeatureStream::join(CRF_FeatureStream* in_stream) { return new CRF_JoinedFeatureStream(this, in_stream); }

// the current utterance of the joined second stream (when the model has one) appended to ftrs2
static void read_second(CRF_Model* crf, CRF_FeatureStream* s, std::vector<float>& ftrs2) {
	if (!crf->secondStreamFtrs()) return;
	CRF_FeatureStream* b = s->joinedStream();
	if (!b) throw runtime_error("the model was configured with a second feature stream (setSecondStream): hand a joined stream (CRF_FeatureStream::join) to the builders / decoder");
	if (b->num_ftrs() != crf->secondStreamFtrs()) throw runtime_error("width of the joined second stream differs from setSecondStream's");
	if (!read_utterance(b, ftrs2, nullptr)) throw runtime_error("No features read from this sentence (second stream)");
}

double CRF_GradBuilder::buildGradient(CRF_FeatureStream* ftr_strm, double* grad, double* Zx_out) {
	ftr_buf.clear(); lab_buf.clear(); ftr2_buf.clear();
	if (!read_utterance(ftr_strm, ftr_buf, &lab_buf)) throw runtime_error("No features read from this sentence");       // CRF_NewGradBuilder.cpp
	read_second(crf, ftr_strm, ftr2_buf);
	const uint32_t off[2] = {0, (uint32_t)lab_buf.size()};
	tmp_grad.assign(crf->getLambdaLen(), 0.0);
	double numer = 0.0;
	if (!crf->lambdaOnDevice) check(crfgpu_set_lambda(crf->gpu(), crf->getLambda(), crf->getLambdaLen()), "crfgpu_set_lambda");
	check(crfgpu_fwdbwd_batch2(crf->gpu(), 1, off, ftr_buf.data(), ftr2_buf.empty() ? nullptr : ftr2_buf.data(), lab_buf.data(), tmp_grad.data(), &numer, Zx_out), "crfgpu_fwdbwd_batch");
	for (QNUInt32 i = 0; i < crf->getLambdaLen(); i++) grad[i] += tmp_grad[i];
	return numer;
}

CRF_Minibatch_GradAccumulator::CRF_Minibatch_GradAccumulator(CRF_Model* myCrf, CRF_FeatureStream* stream, QNUInt32 myNStreams)
    : crf(myCrf), nStreams(myNStreams), minibatch(CRF_UINT32_MAX) {
	if (!stream) throw runtime_error("CRF_Minibatch_GradAccumulator: Cannot open the feature stream");
	if (!nStreams) throw runtime_error("CRF_Minibatch_GradAccumulator: the number of streams must be positive");
	if (nStreams == 1) { ftrStrms.push_back(stream); owned.push_back(false); }
	else {
		// the contiguous views of CRF_FeatureStreamManager.cpp:425-464
		std::vector<uint32_t> first(nStreams), count(nStreams);
		check(crfgpu_shard_views(stream->num_segs(), nStreams, first.data(), count.data()), "crfgpu_shard_views");
		for (QNUInt32 s = 0; s < nStreams; s++) {
			CRF_FeatureStream* v = stream->newView(first[s], count[s]);
			if (!v) throw runtime_error("CRF_Minibatch_GradAccumulator: Cannot get the child feature stream (the stream does not implement newView)");
			ftrStrms.push_back(v); owned.push_back(true);
		}
	}
	strmsSegids.assign(nStreams, QN_SEGID_BAD);
	dev.resize(crf->nDevices());
}
CRF_Minibatch_GradAccumulator::~CRF_Minibatch_GradAccumulator() {
	for (size_t s = 0; s < ftrStrms.size(); s++) if (owned[s]) delete ftrStrms[s];
}

void CRF_Minibatch_GradAccumulator::setMinibatch(QNUInt32 mb) {
	if (mb < nStreams)      // also catches 0: the reference tests this first, so its "0 = totally batch" branch is unreachable (.cpp:175-187)
		throw runtime_error("CRF_Minibatch_GradAccumulator::setMinibatch() Error: minibatch size (" + std::to_string(mb) + ") is less than the number of threads (" + std::to_string(nStreams) + ").");
	minibatch = mb == 0 ? CRF_UINT32_MAX : mb;
}

void CRF_Minibatch_GradAccumulator::rewindAllAndNextSegs() {
	for (QNUInt32 s = 0; s < nStreams; s++) { ftrStrms[s]->rewind(); strmsSegids[s] = ftrStrms[s]->nextseg(); }
}

QNUInt32 CRF_Minibatch_GradAccumulator::planShares(QNUInt32 mb, QNUInt32 n, const std::vector<bool>& atEnd, std::vector<QNUInt32>* share) {
	share->assign(n, 0);
	QNUInt32 active = 0;
	for (QNUInt32 s = 0; s < n; s++) {
		if (atEnd[s]) continue;
		(*share)[s] = crfgpu_minibatch_share(mb, n, s);       // floor(mb/n) + (s < mb % n); CRF_UINT32_MAX = the rest of the view
		active++;
	}
	return active;
}

double CRF_Minibatch_GradAccumulator::runBatch(double* grad, double* Zx_out, QNUInt32* uttCount, bool* isEndOfIter, QNUInt32* nActiveOut) {
	const size_t nDev = crf->nDevices();
	if (minibatch < nStreams) throw runtime_error("CRF_Minibatch_GradAccumulator::accumulateGradient() Error: minibatch size is less than the number of threads.");
	std::vector<bool> atEnd(nStreams);
	for (QNUInt32 s = 0; s < nStreams; s++) atEnd[s] = strmsSegids[s] == QN_SEGID_BAD;
	std::vector<QNUInt32> share;
	const QNUInt32 nActive = planShares(minibatch, nStreams, atEnd, &share);
	if (!nActive)
		throw runtime_error("All feature streams are at the end! You don't have any utterances or you forget to rewind all the streams.\n"
		                    "For the latter case, run rewindAllAndNextSegs().");
	for (DevBatch& b : dev) { b.off.assign(1, 0); b.ftrs.clear(); b.ftrs2.clear(); b.labs.clear(); }
	QNUInt32 total = 0;
	for (QNUInt32 s = 0; s < nStreams; s++) {
		if (atEnd[s]) continue;
		DevBatch& b = dev[s % nDev];
		QNUInt32 cnt = 0;
		do {       // the thread's do-while (.cpp:46-96): at least one utterance, stop at the share or at the end of the view
			if (!read_utterance(ftrStrms[s], b.ftrs, &b.labs)) throw runtime_error("No features read from this sentence");
			read_second(crf, ftrStrms[s], b.ftrs2);
			b.off.push_back((uint32_t)b.labs.size());
			cnt++;
			strmsSegids[s] = ftrStrms[s]->nextseg();
			if (cnt >= share[s]) break;
		} while (strmsSegids[s] != QN_SEGID_BAD);
		total += cnt;
	}
	static const float no_f = 0.0f; static const uint32_t no_l = 0;
	for (size_t d = 0; d < nDev; d++) {
		DevBatch& b = dev[d];
		crfgpu_handle h = crf->gpu(d);
		if (!crf->lambdaOnDevice) check(crfgpu_set_lambda(h, crf->getLambda(), crf->getLambdaLen()), "crfgpu_set_lambda");
		const uint32_t n = (uint32_t)b.off.size() - 1;      // may be 0 on a device whose streams are exhausted: it still joins the all-reduce
		check(crfgpu_stage_batch2(h, n, b.off.data(), n ? b.ftrs.data() : &no_f, (n && !b.ftrs2.empty()) ? b.ftrs2.data() : &no_f, n ? b.labs.data() : &no_l), "crfgpu_stage_batch");
		check(crfgpu_fwdbwd_staged(h), "crfgpu_fwdbwd_staged");          // asynchronous: the devices compute side by side
	}
	if (nDev > 1) {
		check(crfgpu_group_start(), "crfgpu_group_start");
		for (size_t d = 0; d < nDev; d++) check(crfgpu_allreduce_grad(crf->gpu(d)), "crfgpu_allreduce_grad");
		check(crfgpu_group_end(), "crfgpu_group_end");
	}
	double tail[4];
	check(crfgpu_fetch_tail(crf->gpu(0), tail), "crfgpu_fetch_tail");       // [sum numer, sum logZ, n_utt, 0] over ALL devices
	if (!std::isfinite(tail[1]) || !std::isfinite(tail[0]))
		throw std::overflow_error("non-finite log partition function in this minibatch (the reference throws from CRF_LogMath)");
	if ((QNUInt32)tail[2] != total) throw runtime_error("CRF_Minibatch_GradAccumulator: utterance count of the device batches differs from the host's");
	if (grad) {
		check(crfgpu_fetch_fwdbwd(crf->gpu(0), grad, nullptr, nullptr), "crfgpu_fetch_fwdbwd");
		// averaging the gradient over streams, not utterances (.cpp:306-308)
		for (QNUInt32 i = 0; i < crf->getLambdaLen(); i++) grad[i] /= nActive;
	}
	QNUInt32 nEnd = 0;
	for (QNUInt32 s = 0; s < nStreams; s++) if (strmsSegids[s] == QN_SEGID_BAD) nEnd++;
	*isEndOfIter = nEnd == nStreams;
	*uttCount = total; *Zx_out = tail[1];
	if (nActiveOut) *nActiveOut = nActive;
	return tail[0];
}

double CRF_Minibatch_GradAccumulator::accumulateGradient(double* grad, double* Zx_out, QNUInt32* uttCount, bool* isEndOfIter) {
	return runBatch(grad, Zx_out, uttCount, isEndOfIter, nullptr);
}
double CRF_Minibatch_GradAccumulator::accumulateGradientOnDevice(double* Zx_out, QNUInt32* uttCount, bool* isEndOfIter, QNUInt32* nActive) {
	return runBatch(nullptr, Zx_out, uttCount, isEndOfIter, nActive);
}

void CRF_Model::syncLambdaFromDevice() {
	check(crfgpu_get_lambda(gpu(), lambda.data(), lambdaOnDevice ? lambdaAcc.data() : nullptr, nullptr, lambdaOnDevice ? gradSqrAcc.data() : nullptr), "crfgpu_get_lambda");
}

// ---------------------------------------------------------------------------------------------- trainer
CRF_SGTrainer::CRF_SGTrainer(CRF_Model* crf_in, CRF_FeatureStream* stream, const char* wt_fname)
    : crf_ptr(crf_in), strm(stream), weight_fname(wt_fname), lr(0.008f), lr_decay_rate(1.0f), maxIters(1), minibatch(1), nStreams(1),
      useAdagrad(false), eta(1.0), eps(1e-6), useGvar(false), gvar(0.0f), presentations(0) {}

static void touch(const std::string& f) { std::ofstream o(f.c_str()); if (!o.is_open()) std::fprintf(stderr, "ERROR: cannot touch the done file %s\n", f.c_str()); }

void CRF_SGTrainer::train() {
	const QNUInt32 len = crf_ptr->getLambdaLen();
	const size_t found = weight_fname.find_last_of('/');
	const std::string weight_dir = found == std::string::npos ? "." : weight_fname.substr(0, found);     // CRF_Trainer.cpp:24-30
	int iCounter = (int)crf_ptr->getInitIter();
	for (int i = 0; i < iCounter; ++i) strm->rewind();      // keeps a random-order generator in step with the interrupted run (:88-93); a no-op for seq order
	CRF_Minibatch_GradAccumulator gaccum(crf_ptr, strm, nStreams);
	gaccum.setMinibatch(minibatch);
	// lambda and the trainer's accumulators move to every device once; a resumed run carries lambdaAcc (average x presentations) and the
	// AdaGrad sums in (CRF_SGTrainer.cpp:133-145), lambdaSqrAcc restarts at zero as in the reference (:131,149)
	crf_ptr->pushLambdaToDevices();
	for (size_t d = 0; d < crf_ptr->nDevices(); d++)
		check(crfgpu_set_train_state(crf_ptr->gpu(d), crf_ptr->getLambdaAcc(), nullptr, crf_ptr->getGradSqrAcc()), "crfgpu_set_train_state");
	crf_ptr->lambdaOnDevice = true;
	iterLogLi.clear();
	int accCnt = (int)crf_ptr->getPresentations();
	std::vector<double> avg(len, 0.0);
	float invSquareVar = 0.0f;
	if (useGvar) invSquareVar = 1 / gvar;                   // float, and 1/gvar not 1/gvar^2: the reference's own arithmetic (:164-167)
	gaccum.rewindAllAndNextSegs();
	double totLogLi = 0.0;
	while (iCounter < maxIters) {
		double Zx = 0.0; QNUInt32 cnt = 0, nActive = 0; bool eoi = false;
		const double num = gaccum.accumulateGradientOnDevice(&Zx, &cnt, &eoi, &nActive);
		totLogLi += num - Zx;
		crfgpu_sgd opt = {(double)lr, useGvar ? 1u : 0u, (double)invSquareVar, useAdagrad ? 1u : 0u, eta, eps};
		for (size_t d = 0; d < crf_ptr->nDevices(); d++)      // every device applies the identical update to its copy of lambda
			check(crfgpu_sgd_update(crf_ptr->gpu(d), &opt, (double)nActive), "crfgpu_sgd_update");
		accCnt += (int)cnt; presentations = accCnt;
		if (!eoi) continue;
		iterLogLi.push_back(totLogLi);
		crf_ptr->syncLambdaFromDevice();
		const std::string base = weight_fname + ".i" + std::to_string(iCounter);
		crf_ptr->writeToFile((base + ".out").c_str());
		for (QNUInt32 i = 0; i < len; i++) avg[i] = crf_ptr->getLambdaAcc()[i] / (float)accCnt;      // CRF_SGTrainer.cpp:357
		crf_ptr->writeToFile((base + ".avg.out").c_str(), avg.data(), len);
		if (useAdagrad) crf_ptr->writeToFile((base + ".gradSqrAcc.out").c_str(), crf_ptr->getGradSqrAcc(), len);      // :370-381
		gaccum.rewindAllAndNextSegs();
		touch(weight_dir + "/.done.train.i" + std::to_string(iCounter));
		iCounter++;
		totLogLi = 0.0;
		if (!useAdagrad) lr *= lr_decay_rate;
	}
	crf_ptr->writeToFile(weight_fname.c_str());
	crf_ptr->writeToFile((weight_fname + ".avg.out").c_str(), avg.data(), len);
	touch(weight_dir + "/.done.train");
	crf_ptr->lambdaOnDevice = false;
}

// ---------------------------------------------------------------------------------------------- decoding
static void arcs_from_segments(const uint32_t* lab, const uint32_t* dur, const uint32_t* phn, uint32_t nseg, std::vector<CRF_BestPathArc>* result) {
	result->clear();
	for (uint32_t k = 0; k < nseg; k++)
		result->push_back(CRF_BestPathArc{(int)lab[k] + 1, phn[k] == CRFGPU_LAB_BAD ? 0 : (int)phn[k] + 1, dur[k]});
}

void CRF_ViterbiDecoder_StdSeg_NoSegTransFtr::setLm(const CRF_PhoneBigramLm* lm_fst) {
	if (!lm_fst) { check(crfgpu_set_phone_lm(crf->gpu(), nullptr, nullptr, nullptr), "crfgpu_set_phone_lm"); return; }
	const size_t P = lm_fst->start.size();
	if (crf->numStates() > 1) {      // N states per phone: unigram costs + the exit costs of the epsilon arcs back to the LM's start state
		if (lm_fst->bigram.size() != P || lm_fst->final_wt.size() != P) throw runtime_error("CRF_PhoneBigramLm (N states per phone): start[P] = unigram costs, bigram[P] = exit costs, final_wt[P] expected");
		check(crfgpu_set_phone_unigram_lm(crf->gpu(), lm_fst->start.data(), lm_fst->bigram.data(), lm_fst->final_wt.data()), "crfgpu_set_phone_unigram_lm");
		return;
	}
	if (lm_fst->bigram.size() != P * P || lm_fst->final_wt.size() != P) throw runtime_error("CRF_PhoneBigramLm: start[P], bigram[P*P], final_wt[P] expected");
	check(crfgpu_set_phone_lm(crf->gpu(), lm_fst->start.data(), lm_fst->bigram.data(), lm_fst->final_wt.data()), "crfgpu_set_phone_lm");
}

int CRF_ViterbiDecoder_StdSeg_NoSegTransFtr::nStateDecode(std::vector<CRF_BestPathArc>* result, const CRF_PhoneBigramLm* lm_fst, float* path_cost, double beam) {
	setLm(lm_fst);
	struct Drop { CRF_ViterbiDecoder_StdSeg_NoSegTransFtr* d; ~Drop() { try { d->setLm(nullptr); } catch (...) {} } } drop{this};   // the LM belongs to this call, as in the reference
	return nStateDecode(result, path_cost, beam);
}

int CRF_ViterbiDecoder_StdSeg_NoSegTransFtr::nStateDecode(std::vector<CRF_BestPathArc>* result, float* path_cost, double beam) {
	check(crfgpu_set_beam(crf->gpu(), beam), "crfgpu_set_beam");      // input_beam of nStateDecode (one state per phone; 0 = no pruning)
	std::vector<float> f;
	const uint32_t T = (uint32_t)read_utterance(strm, f, nullptr);
	if (!T) throw runtime_error("No features read from this sentence");
	std::vector<float> f2; read_second(crf, strm, f2);
	const uint32_t off[2] = {0, T};
	std::vector<uint32_t> lab(T), dur(T), phn(T); uint32_t nseg = 0; float cost = 0.0f;
	check(crfgpu_set_lambda(crf->gpu(), crf->getLambda(), crf->getLambdaLen()), "crfgpu_set_lambda");
	check(crfgpu_viterbi_batch2(crf->gpu(), 1, off, f.data(), f2.empty() ? nullptr : f2.data(), lab.data(), dur.data(), phn.data(), &nseg, &cost), "crfgpu_viterbi_batch");
	arcs_from_segments(lab.data(), dur.data(), phn.data(), nseg, result);
	if (path_cost) *path_cost = cost;
	return (int)T;
}

size_t CRF_ViterbiDecoder_StdSeg_NoSegTransFtr::nStateDecodeBatch(size_t max_utts, std::vector<std::vector<CRF_BestPathArc>>* results,
                                                                   std::vector<float>* path_costs, std::vector<int>* n_frames, bool* stream_end, double beam) {
	check(crfgpu_set_beam(crf->gpu(), beam), "crfgpu_set_beam");
	std::vector<float> f, f2; std::vector<uint32_t> off(1, 0);
	if (stream_end) *stream_end = false;
	while (off.size() - 1 < max_utts) {
		const size_t T = read_utterance(strm, f, nullptr);
		if (!T) throw runtime_error("No features read from this sentence");
		read_second(crf, strm, f2);
		off.push_back(off.back() + (uint32_t)T);
		if (strm->nextseg() == QN_SEGID_BAD) { if (stream_end) *stream_end = true; break; }
	}
	const uint32_t n = (uint32_t)off.size() - 1, N = off.back();
	std::vector<uint32_t> lab(N), dur(N), phn(N), nseg(n); std::vector<float> cost(n);
	check(crfgpu_set_lambda(crf->gpu(), crf->getLambda(), crf->getLambdaLen()), "crfgpu_set_lambda");      // once per batch, not per utterance
	check(crfgpu_viterbi_batch2(crf->gpu(), n, off.data(), f.data(), f2.empty() ? nullptr : f2.data(), lab.data(), dur.data(), phn.data(), nseg.data(), cost.data()), "crfgpu_viterbi_batch");
	results->assign(n, std::vector<CRF_BestPathArc>());
	if (n_frames) n_frames->assign(n, 0);
	for (uint32_t u = 0; u < n; u++) {
		arcs_from_segments(&lab[off[u]], &dur[off[u]], &phn[off[u]], nseg[u], &(*results)[u]);
		if (n_frames) (*n_frames)[u] = (int)(off[u + 1] - off[u]);
	}
	if (path_costs) *path_costs = cost;
	return n;
}
