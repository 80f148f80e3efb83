// crf_host.cpp -- see crf_host.h.  Thin: every arithmetic step happens behind the C ABI on the device.
#include "crf_host.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>

using std::runtime_error;
using std::string;

static void check(int rc, const char* what) {
	if (rc != CRFGPU_OK) throw runtime_error(string(what) + ": " + crfgpu_last_error());   // the reference throws; main() catches and exits
}

// ---------------------------------------------------------------------------------------------- stream
CRF_MemFeatureStream::CRF_MemFeatureStream(const std::vector<uint32_t>& frame_off, const std::vector<float>& f,
                                           const std::vector<QNUInt32>& l, QNUInt32 n_ftrs)
    : off(frame_off), ftrs(f), labs(l), nf(n_ftrs), first(0), count((QNUInt32)frame_off.size() - 1), seg(-1), pos(0) {}

void CRF_MemFeatureStream::view(QNUInt32 startseg, QNUInt32 nsegs) { first = startseg; count = nsegs; rewind(); }
void CRF_MemFeatureStream::rewind() { seg = -1; pos = 0; }
QN_SegID CRF_MemFeatureStream::nextseg() {
	if (seg + 1 >= (long)count) { seg = count; return QN_SEGID_BAD; }
	seg++; pos = off[first + seg];
	return seg;
}
size_t CRF_MemFeatureStream::read(size_t bs, float* fb, QNUInt32* lb) {
	if (seg < 0 || seg >= (long)count) return 0;
	const uint32_t end = off[first + seg + 1];
	const size_t n = std::min<size_t>(bs, end - pos);
	if (n) {
		std::memcpy(fb, &ftrs[(size_t)pos * nf], n * nf * sizeof(float));
		if (lb) for (size_t i = 0; i < n; i++) lb[i] = labs.empty() ? CRF_LAB_BAD : labs[pos + i];
	}
	pos += (uint32_t)n;
	return n;
}

// ---------------------------------------------------------------------------------------------- model
CRF_Model::CRF_Model(QNUInt32 num_labs)
    : nlabs(num_labs), lab_max_dur(1), nActualLabs(num_labs), model_type(STDFRAME), have_map(false), n_base_ftrs(0),
      extract_seg_ftrs(false), handle(nullptr), lambdaOnDevice(false) {}
CRF_Model::~CRF_Model() { if (handle) crfgpu_destroy(handle); }

void CRF_Model::setFeatureMap(const CRF_FeatureMap_config& cfg, QNUInt32 base_ftrs, bool seg_ftrs, int device) {
	if (cfg.map_type != STDSTATE && cfg.map_type != STDTRANS) throw runtime_error("only the dense feature maps are implemented on the device (stdstate, stdtrans)");
	fmap = cfg; have_map = true; n_base_ftrs = base_ftrs; extract_seg_ftrs = seg_ftrs;
	crfgpu_config c;
	std::memset(&c, 0, sizeof(c));
	c.model_type = (uint32_t)model_type; c.n_labs = cfg.numLabs; c.n_base_ftrs = base_ftrs; c.n_states = cfg.numStates;
	c.max_dur = lab_max_dur; c.n_actual_labs = nActualLabs; c.extract_seg_ftrs = seg_ftrs ? 1 : 0;
	c.use_state_ftrs = cfg.useStateFtrs; c.state_fidx_start = cfg.stateFidxStart; c.state_fidx_end = cfg.stateFidxEnd;
	c.use_trans_ftrs = cfg.useTransFtrs; c.trans_fidx_start = cfg.transFidxStart; c.trans_fidx_end = cfg.transFidxEnd;
	c.use_state_bias = cfg.useStateBias; c.use_trans_bias = cfg.useTransBias;
	c.state_bias_val = cfg.stateBiasVal; c.trans_bias_val = cfg.transBiasVal;
	if (crfgpu_window_width(&c) != cfg.numFeas)
		throw runtime_error("CRF_FeatureMap_config::numFeas does not match the window width of the feature stream");
	if (handle) { crfgpu_destroy(handle); handle = nullptr; }
	check(crfgpu_create(&c, device, &handle), "crfgpu_create");
	lambda.assign(crfgpu_lambda_len(handle), 0.0);
	lambdaAcc.assign(lambda.size(), 0.0);
}
crfgpu_handle CRF_Model::gpu() {
	if (!handle) throw runtime_error("CRF_Model: setFeatureMap has not been called");
	return handle;
}
void CRF_Model::setLambda(double* lam, QNUInt32 len) {
	if (len != lambda.size()) throw runtime_error("CRF_Model::setLambda: length mismatch");
	std::copy(lam, lam + len, lambda.begin());
}
void CRF_Model::resetLambda() { std::fill(lambda.begin(), lambda.end(), 0.0); }
bool CRF_Model::writeToFile(const char* fname) {
	std::ofstream ofile(fname);
	if (!ofile.good()) return false;
	for (double v : lambda) ofile << v << std::endl;      // default precision, as CRF_Model.cpp:210-212
	return true;
}
bool CRF_Model::readFromFile(const char* fname) {
	std::ifstream ifile(fname);
	if (!ifile.good()) return false;
	size_t i = 0; double v;
	while (i < lambda.size() && (ifile >> v)) lambda[i++] = v;
	return i == lambda.size();
}

// ---------------------------------------------------------------------------------------------- training
CRF_GradBuilder* CRF_GradBuilder::create(CRF_Model* crf_ptr, objfunctype ofunc) {
	if (ofunc != EXPF) throw runtime_error("only the EXPF objective is implemented (the reference's factory never selects the others either)");
	return new CRF_GradBuilder(crf_ptr);
}

double CRF_GradBuilder::buildGradient(CRF_FeatureStream* ftr_strm, double* grad, double* Zx_out) {
	const QNUInt32 nf = ftr_strm->num_ftrs();
	ftr_buf.clear(); lab_buf.clear();
	std::vector<float> fb(nf); QNUInt32 lb = 0;
	while (ftr_strm->read(1, fb.data(), &lb) == 1) { ftr_buf.insert(ftr_buf.end(), fb.begin(), fb.end()); lab_buf.push_back(lb); }
	if (lab_buf.empty()) throw runtime_error("No features read from this sentence");       // CRF_NewGradBuilder.cpp
	const uint32_t off[2] = {0, (uint32_t)lab_buf.size()};
	tmp_grad.assign(crf->getLambdaLen(), 0.0);
	double numer = 0.0;
	if (!crf->lambdaOnDevice) check(crfgpu_set_lambda(crf->gpu(), crf->getLambda(), crf->getLambdaLen()), "crfgpu_set_lambda");
	check(crfgpu_fwdbwd_batch(crf->gpu(), 1, off, ftr_buf.data(), lab_buf.data(), tmp_grad.data(), &numer, Zx_out), "crfgpu_fwdbwd_batch");
	for (QNUInt32 i = 0; i < crf->getLambdaLen(); i++) grad[i] += tmp_grad[i];
	return numer;
}

CRF_Minibatch_GradAccumulator::CRF_Minibatch_GradAccumulator(CRF_Model* myCrf, CRF_FeatureStream* stream, QNUInt32 myNStreams)
    : crf(myCrf), strm(stream), nStreams(myNStreams), minibatch(myNStreams), started(false) {}

void CRF_Minibatch_GradAccumulator::rewindAllAndNextSegs() { strm->rewind(); started = strm->nextseg() != QN_SEGID_BAD; }

double CRF_Minibatch_GradAccumulator::accumulateGradient(double* grad, double* Zx_out, QNUInt32* uttCount, bool* isEndOfIter) {
	const QNUInt32 nf = strm->num_ftrs();
	off.assign(1, 0); ftrs.clear(); labs.clear();
	std::vector<float> fb(nf); QNUInt32 lb = 0;
	*isEndOfIter = false;
	if (!started) rewindAllAndNextSegs();
	QNUInt32 n = 0;
	while (n < minibatch && started) {
		size_t T = 0;
		while (strm->read(1, fb.data(), &lb) == 1) { ftrs.insert(ftrs.end(), fb.begin(), fb.end()); labs.push_back(lb); T++; }
		if (!T) throw runtime_error("No features read from this sentence");
		off.push_back((uint32_t)labs.size()); n++;
		if (strm->nextseg() == QN_SEGID_BAD) { started = false; *isEndOfIter = true; }
	}
	if (!n) { *isEndOfIter = true; *uttCount = 0; *Zx_out = 0.0; std::fill(grad, grad + crf->getLambdaLen(), 0.0); return 0.0; }
	numer.assign(n, 0.0); logZ.assign(n, 0.0);
	if (!crf->lambdaOnDevice) check(crfgpu_set_lambda(crf->gpu(), crf->getLambda(), crf->getLambdaLen()), "crfgpu_set_lambda");
	check(crfgpu_fwdbwd_batch(crf->gpu(), n, off.data(), ftrs.data(), labs.data(), grad, numer.data(), logZ.data()), "crfgpu_fwdbwd_batch");
	double num = 0.0; *Zx_out = 0.0;
	for (QNUInt32 u = 0; u < n; u++) { num += numer[u]; *Zx_out += logZ[u]; }
	*uttCount = n;
	// the reference divides the summed gradient by the number of ACTIVE streams, not by the utterance count (.cpp:306-308)
	const QNUInt32 nActive = std::min(n, nStreams);
	for (QNUInt32 i = 0; i < crf->getLambdaLen(); i++) grad[i] /= (double)nActive;
	return num;
}

double CRF_Minibatch_GradAccumulator::accumulateGradientOnDevice(double* Zx_out, QNUInt32* uttCount, bool* isEndOfIter, QNUInt32* nActive) {
	const QNUInt32 nf = strm->num_ftrs();
	off.assign(1, 0); ftrs.clear(); labs.clear();
	std::vector<float> fb(nf); QNUInt32 lb = 0;
	*isEndOfIter = false;
	if (!started) rewindAllAndNextSegs();
	QNUInt32 n = 0;
	while (n < minibatch && started) {
		size_t T = 0;
		while (strm->read(1, fb.data(), &lb) == 1) { ftrs.insert(ftrs.end(), fb.begin(), fb.end()); labs.push_back(lb); T++; }
		if (!T) throw runtime_error("No features read from this sentence");
		off.push_back((uint32_t)labs.size()); n++;
		if (strm->nextseg() == QN_SEGID_BAD) { started = false; *isEndOfIter = true; }
	}
	*uttCount = n; *Zx_out = 0.0; *nActive = std::min(n, nStreams);
	if (!n) { *isEndOfIter = true; return 0.0; }
	numer.assign(n, 0.0); logZ.assign(n, 0.0);
	if (!crf->lambdaOnDevice) check(crfgpu_set_lambda(crf->gpu(), crf->getLambda(), crf->getLambdaLen()), "crfgpu_set_lambda");
	check(crfgpu_stage_batch(crf->gpu(), n, off.data(), ftrs.data(), labs.data()), "crfgpu_stage_batch");
	check(crfgpu_fwdbwd_staged(crf->gpu()), "crfgpu_fwdbwd_staged");
	check(crfgpu_fetch_fwdbwd(crf->gpu(), nullptr, numer.data(), logZ.data()), "crfgpu_fetch_fwdbwd");   // scalars only: the gradient stays in HBM
	double num = 0.0;
	for (QNUInt32 u = 0; u < n; u++) { num += numer[u]; *Zx_out += logZ[u]; }
	return num;
}

void CRF_Model::syncLambdaFromDevice() {
	check(crfgpu_get_lambda(gpu(), lambda.data(), lambdaOnDevice ? lambdaAcc.data() : nullptr, nullptr, nullptr), "crfgpu_get_lambda");
}

// ---------------------------------------------------------------------------------------------- trainer
CRF_SGTrainer::CRF_SGTrainer(CRF_Model* crf_in, CRF_FeatureStream* stream, const char* wt_fname)
    : crf_ptr(crf_in), strm(stream), weight_fname(wt_fname), lr(0.008f), lr_decay_rate(1.0f), maxIters(1), minibatch(1), nStreams(1),
      useAdagrad(false), eta(1.0), eps(1e-6), useGvar(false), invSquareVar(0.0) {}

static bool write_values(const std::string& fname, const double* v, size_t n) {
	FILE* f = std::fopen(fname.c_str(), "w");
	if (!f) return false;
	for (size_t i = 0; i < n; i++) std::fprintf(f, "%g\n", v[i]);      // default ostream precision: 6 significant digits (CRF_Model.cpp:210-212)
	return std::fclose(f) == 0;
}

void CRF_SGTrainer::train() {
	const QNUInt32 len = crf_ptr->getLambdaLen();
	CRF_Minibatch_GradAccumulator gaccum(crf_ptr, strm, nStreams);
	gaccum.setMinibatch(minibatch);
	check(crfgpu_set_lambda(crf_ptr->gpu(), crf_ptr->getLambda(), len), "crfgpu_set_lambda");
	check(crfgpu_set_train_state(crf_ptr->gpu(), nullptr, nullptr, nullptr), "crfgpu_set_train_state");
	crf_ptr->lambdaOnDevice = true;
	iterLogLi.clear();
	QNUInt32 accCnt = 0;
	std::vector<double> avg(len);
	for (int iCounter = 1; iCounter <= maxIters; iCounter++) {
		double totLogLi = 0.0; bool eoi = false;
		gaccum.rewindAllAndNextSegs();
		while (!eoi) {
			double Zx = 0.0; QNUInt32 cnt = 0, nActive = 0;
			const double num = gaccum.accumulateGradientOnDevice(&Zx, &cnt, &eoi, &nActive);
			if (!cnt) break;
			totLogLi += num - Zx;
			crfgpu_sgd opt = {(double)lr, useGvar ? 1u : 0u, invSquareVar, useAdagrad ? 1u : 0u, eta, eps};
			check(crfgpu_sgd_update(crf_ptr->gpu(), &opt, (double)nActive), "crfgpu_sgd_update");
			accCnt += cnt;
		}
		iterLogLi.push_back(totLogLi);
		crf_ptr->syncLambdaFromDevice();
		const std::string base = weight_fname + ".i" + std::to_string(iCounter);
		if (!crf_ptr->writeToFile((base + ".out").c_str())) throw runtime_error("ERROR! File " + base + ".out unable to be opened for writing.");
		for (QNUInt32 i = 0; i < len; i++) avg[i] = crf_ptr->getLambdaAcc()[i] / (float)accCnt;      // CRF_SGTrainer.cpp:357
		if (!write_values(base + ".avg.out", avg.data(), len)) throw runtime_error("ERROR! File " + base + ".avg.out unable to be opened for writing.");
		{ FILE* f = std::fopen((weight_fname + ".done.train.i" + std::to_string(iCounter)).c_str(), "w"); if (f) std::fclose(f); }
		if (!useAdagrad) lr *= lr_decay_rate;
	}
	if (!crf_ptr->writeToFile(weight_fname.c_str())) throw runtime_error("ERROR! File " + weight_fname + " unable to be opened for writing.");
	if (accCnt) { for (QNUInt32 i = 0; i < len; i++) avg[i] = crf_ptr->getLambdaAcc()[i] / (float)accCnt; write_values(weight_fname + ".avg.out", avg.data(), len); }
	crf_ptr->lambdaOnDevice = false;
}

// ---------------------------------------------------------------------------------------------- decoding
int CRF_ViterbiDecoder_StdSeg_NoSegTransFtr::nStateDecode(std::vector<CRF_BestPathArc>* result, float* path_cost, double beam) {
	if (beam > 0.0) throw runtime_error("beam pruning / LM-constrained decoding is not implemented on the device (free-phone LM, beam 0 only)");
	const QNUInt32 nf = strm->num_ftrs();
	std::vector<float> f, fb(nf);
	while (strm->read(1, fb.data(), nullptr) == 1) f.insert(f.end(), fb.begin(), fb.end());
	const uint32_t T = (uint32_t)(f.size() / nf);
	if (!T) throw runtime_error("No features read from this sentence");
	const uint32_t off[2] = {0, T};
	std::vector<uint32_t> lab(T), dur(T), phn(T); uint32_t nseg = 0; float cost = 0.0f;
	check(crfgpu_set_lambda(crf->gpu(), crf->getLambda(), crf->getLambdaLen()), "crfgpu_set_lambda");
	check(crfgpu_viterbi_batch(crf->gpu(), 1, off, f.data(), lab.data(), dur.data(), phn.data(), &nseg, &cost), "crfgpu_viterbi_batch");
	result->clear();
	for (uint32_t k = 0; k < nseg; k++)
		result->push_back(CRF_BestPathArc{(int)lab[k] + 1, phn[k] == CRFGPU_LAB_BAD ? 0 : (int)phn[k] + 1, dur[k]});
	if (path_cost) *path_cost = cost;
	return (int)T;
}
