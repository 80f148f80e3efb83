// host_selftest -- drives the C++ host layer (crf_host.h) the way CRFTrain / CRFDecode drive the reference classes and checks the
// results against expected values from a case file written by tests/test_host_cpp.py (goldens of the unmodified reference).
//   host_selftest <case.txt>      run training (+ Viterbi if the case has a path) on cuda:0, exit 0 on agreement
//   host_selftest <case.txt> acc <nStreams> <minibatch> <nDevices>
//                                 one epoch through CRF_Minibatch_GradAccumulator::accumulateGradient with nStreams corpus views on
//                                 nDevices GPUs; every minibatch's [uttCount, endOfIter, numerator, Zx, grad...] is appended (doubles)
//                                 to <case.txt>.acc.bin for the caller to compare with the reference's rule
//   host_selftest <case.txt> resume
//                                 AdaGrad training: 2 iterations in one run == 1 iteration, then a resumed second one
//   host_selftest --plan <n_utt> <nStreams> <minibatch>
//                                 no device: prints the minibatch composition of one epoch ("stream:first+count ..." per call, "/nActive")
//   host_selftest --no-device     verify that the layer fails loudly (std::runtime_error) when there is no CUDA device
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>

#include "crf_host.h"

static CRF_FeatureMap_config fmap_config(QNUInt32 n_labs, QNUInt32 n_states, QNUInt32 width, QNUInt32 max_dur, QNUInt32 n_act) {
	// mirrors set_fmap_config (CRFTrain/src/Main.cpp:372-430) for crf_featuremap=stdstate with both biases
	CRF_FeatureMap_config c;
	std::memset(&c, 0, sizeof(c));
	c.map_type = STDSTATE; c.numLabs = n_labs; c.numFeas = width; c.numStates = n_states;
	c.useStateFtrs = true; c.stateFidxStart = 0; c.stateFidxEnd = width - 1;
	c.useTransFtrs = false; c.transFidxStart = 0; c.transFidxEnd = width - 1;
	c.useStateBias = true; c.useTransBias = true; c.stateBiasVal = 1.0; c.transBiasVal = 1.0;
	c.maxDur = max_dur; c.durFtrStart = 0; c.nActualLabs = n_act;
	return c;
}

template <class T> static void read_vec(std::istream& in, std::vector<T>& v, size_t n) {
	v.resize(n);
	for (size_t i = 0; i < n; i++) { double x; in >> x; v[i] = (T)x; }
}

int main(int argc, char** argv) {
	if (argc < 2) { std::fprintf(stderr, "usage: host_selftest <case.txt> | --no-device\n"); return 2; }
	try {
		if (!std::strcmp(argv[1], "--no-device")) {
			CRF_Model m(4);
			try { m.setFeatureMap(fmap_config(4, 1, 3, 1, 4), 3, false); }
			catch (const std::runtime_error& e) { std::printf("failed loudly as expected: %s\n", e.what()); return 0; }
			std::printf("a CUDA device is present: nothing to check\n");
			return 0;
		}
		if (!std::strcmp(argv[1], "--plan")) {
			if (argc < 5) throw std::runtime_error("--plan <n_utt> <nStreams> <minibatch>");
			const QNUInt32 n = (QNUInt32)std::atoi(argv[2]), ns = (QNUInt32)std::atoi(argv[3]), mb = (QNUInt32)std::atoi(argv[4]);
			if (mb < ns) throw std::runtime_error("minibatch size is less than the number of threads");
			std::vector<uint32_t> first(ns), count(ns), pos(ns, 0);
			if (crfgpu_shard_views(n, ns, first.data(), count.data()) != CRFGPU_OK) throw std::runtime_error(crfgpu_last_error());
			for (;;) {
				std::vector<bool> atEnd(ns);
				for (QNUInt32 s = 0; s < ns; s++) atEnd[s] = pos[s] >= count[s];
				std::vector<QNUInt32> share;
				const QNUInt32 act = CRF_Minibatch_GradAccumulator::planShares(mb, ns, atEnd, &share);
				if (!act) break;
				for (QNUInt32 s = 0; s < ns; s++) {
					if (atEnd[s]) continue;
					const QNUInt32 take = std::min<QNUInt32>(std::max<QNUInt32>(share[s], 1), count[s] - pos[s]);
					std::printf("%u:%u+%u ", s, first[s] + pos[s], take);
					pos[s] += take;
				}
				std::printf("/%u\n", act);
			}
			return 0;
		}
		std::ifstream in(argv[1]);
		if (!in.good()) throw std::runtime_error("cannot open case file");
		QNUInt32 mode, mt, n_labs, n_base, n_states, max_dur, n_act, segf, n_utt, N, len, n_arcs;   // mode 0: training case, 1: decoding case
		in >> mode >> mt >> n_labs >> n_base >> n_states >> max_dur >> n_act >> segf >> n_utt >> N >> len >> n_arcs;
		std::vector<uint32_t> off; std::vector<double> lam, logZ, numer, grad; std::vector<float> ftrs; std::vector<QNUInt32> labs;
		std::vector<int> arcs;
		read_vec(in, off, n_utt + 1); read_vec(in, lam, len); read_vec(in, ftrs, (size_t)N * n_base); read_vec(in, labs, N);
		read_vec(in, logZ, n_utt); read_vec(in, numer, n_utt); read_vec(in, grad, len); read_vec(in, arcs, (size_t)n_arcs * 3);
		if (!in.good() && !in.eof()) throw std::runtime_error("short case file");
		QNUInt32 width = max_dur == 1 ? n_base : (segf ? 8 * n_base + max_dur : n_base);
		// optional trailer: a second feature stream joined behind the first (the TIMIT recipe's layout) and the feature ranges of the map
		QNUInt32 n_base2 = 0, segf2 = 0, lc2 = 0, rc2 = 0, bd2 = 0, use_tf = 0, sf0 = 0, sf1 = 0, tf0 = 0, tf1 = 0, N2 = 0;
		std::vector<float> ftrs2; std::vector<uint32_t> off2;
		if (in >> n_base2 >> segf2 >> lc2 >> rc2 >> bd2 >> use_tf >> sf0 >> sf1 >> tf0 >> tf1 >> N2) {
			read_vec(in, ftrs2, (size_t)N2 * n_base2);
			for (QNUInt32 u = 0; u <= n_utt; u++) off2.push_back(off[u] + u * (lc2 + rc2));
			if (off2.back() != N2) throw std::runtime_error("second stream: frame count does not match the context lengths");
			width += max_dur == 1 ? (lc2 + 1 + rc2) * n_base2 : (segf2 ? 8 * n_base2 + max_dur + (lc2 + rc2) * n_base2 : (bd2 ? std::min(lc2, rc2 + 1) : lc2 + 1 + rc2) * n_base2);
		}

		CRF_Model my_crf(n_labs);
		my_crf.setLabMaxDur(max_dur); my_crf.setNActualLabs(n_act); my_crf.setModelType((modeltype)mt);
		CRF_FeatureMap_config fcfg = fmap_config(n_labs, n_states, width, max_dur, n_act);
		if (n_base2) {
			my_crf.setSecondStream(n_base2, segf2 != 0, lc2, rc2, bd2 != 0);
			fcfg.stateFidxStart = sf0; fcfg.stateFidxEnd = sf1;
			if (use_tf) { fcfg.map_type = STDTRANS; fcfg.useTransFtrs = true; fcfg.transFidxStart = tf0; fcfg.transFidxEnd = tf1; }
		}
		my_crf.setFeatureMap(fcfg, n_base, segf != 0);
		if (my_crf.getLambdaLen() != len) throw std::runtime_error("lambda length differs from the reference's");
		my_crf.setLambda(lam.data(), len);

		int bad = 0;
		auto close = [&](double a, double b, double tol, const char* what) {
			if (std::fabs(a - b) > tol * std::fmax(1.0, std::fabs(b))) { std::printf("MISMATCH %s: %.12g vs %.12g\n", what, a, b); bad++; }
		};
		const std::string sub = argc > 2 ? argv[2] : "";
		if (n_base2) {
			// ---- joined second stream: minibatch seam over two views of the JOINED stream, per-utterance seam, decode seam ----
			CRF_MemFeatureStream s1(off, ftrs, labs.empty() ? std::vector<QNUInt32>(N, 0) : labs, n_base);
			CRF_MemFeatureStream s2(off2, ftrs2, std::vector<QNUInt32>(N2, 0), n_base2);
			std::unique_ptr<CRF_FeatureStream> js(s1.join(&s2));
			if (mode == 0) {
				CRF_Minibatch_GradAccumulator gaccum(&my_crf, js.get(), 2);
				std::vector<double> g(len, 0.0), gsum(len, 0.0); double Zx = 0.0, zacc = 0.0, nacc = 0.0; QNUInt32 cnt = 0, tot = 0; bool eoi = false;
				gaccum.setMinibatch(4);
				gaccum.rewindAllAndNextSegs();
				while (!eoi) {
					const double num = gaccum.accumulateGradient(g.data(), &Zx, &cnt, &eoi);
					// undo the division by the active streams so that the minibatches add up to the golden's batch gradient
					const double nact = std::min<QNUInt32>(2, cnt);       // two views whose sizes differ by at most one, shares of two utterances: both active unless ONE utterance is left
					for (QNUInt32 i = 0; i < len; i++) gsum[i] += g[i] * nact;
					zacc += Zx; nacc += num; tot += cnt;
				}
				double zsum = 0, nsum = 0, gmax = 0;
				for (QNUInt32 u = 0; u < n_utt; u++) { zsum += logZ[u]; nsum += numer[u]; }
				for (double v : grad) gmax = std::fmax(gmax, std::fabs(v));
				close(zacc, zsum, 1e-5, "sum logZ (joined)"); close(nacc, nsum, 1e-5, "sum numerator (joined)");
				if (tot != n_utt) { std::printf("MISMATCH joined uttCount %u\n", tot); bad++; }
				for (QNUInt32 i = 0; i < len; i++)
					if (std::fabs(gsum[i] - grad[i]) > 1e-4 * gmax + 1e-4 * std::fabs(grad[i])) { if (bad < 5) std::printf("MISMATCH joined grad[%u]: %.9g vs %.9g\n", i, gsum[i], grad[i]); bad++; }
				std::unique_ptr<CRF_GradBuilder> gb(CRF_GradBuilder::create(&my_crf, EXPF));
				std::vector<double> g2(len, 0.0);
				js->rewind();
				QNUInt32 u = 0;
				while (js->nextseg() != QN_SEGID_BAD) {
					double z = 0.0; const double nu = gb->buildGradient(js.get(), g2.data(), &z);
					close(z, logZ[u], 1e-5, "logZ (joined)"); close(nu, numer[u], 1e-5, "numerator (joined)"); u++;
				}
				for (QNUInt32 i = 0; i < len; i++)
					if (std::fabs(g2[i] - grad[i]) > 1e-4 * gmax + 1e-4 * std::fabs(grad[i])) { if (bad < 5) std::printf("MISMATCH joined grad2[%u]\n", i); bad++; }
			} else {
				js->rewind(); js->nextseg();
				CRF_ViterbiDecoder_StdSeg_NoSegTransFtr vd(js.get(), &my_crf);
				std::vector<CRF_BestPathArc> path; float cost = 0.0f;
				const int frames = vd.nStateDecode(&path, &cost);
				if (frames != (int)(off[1] - off[0]) || path.size() != n_arcs) { std::printf("MISMATCH joined decode: %d frames, %zu arcs\n", frames, path.size()); bad++; }
				else for (QNUInt32 k = 0; k < n_arcs; k++)
					if (path[k].ilabel != arcs[3 * k] || path[k].olabel != arcs[3 * k + 1] || (int)path[k].dur != arcs[3 * k + 2]) { std::printf("MISMATCH joined arc %u\n", k); bad++; }
				js->rewind(); js->nextseg();
				std::vector<std::vector<CRF_BestPathArc>> res; std::vector<float> costs; std::vector<int> nfr; bool end = false;
				const size_t done = vd.nStateDecodeBatch(n_utt, &res, &costs, &nfr, &end);
				if (done != n_utt || !end || res[0].size() != path.size() || costs[0] != cost) { std::printf("MISMATCH joined batch decode\n"); bad++; }
			}
			if (bad) { std::printf("host_selftest: %d mismatches\n", bad); return 1; }
			std::printf("host_selftest ok (joined streams)\n");
			return 0;
		}
		if (sub == "acc") {
			if (argc < 6) throw std::runtime_error("acc <nStreams> <minibatch> <nDevices>");
			const QNUInt32 ns = (QNUInt32)std::atoi(argv[3]), mb = (QNUInt32)std::atoi(argv[4]); const int nd = std::atoi(argv[5]);
			if (nd > 1) { std::vector<int> more; for (int d = 1; d < nd; d++) more.push_back(d); my_crf.addDevices(more); }
			CRF_MemFeatureStream strm(off, ftrs, labs, n_base);
			CRF_Minibatch_GradAccumulator gaccum(&my_crf, &strm, ns);
			gaccum.setMinibatch(mb);
			gaccum.rewindAllAndNextSegs();
			FILE* out = std::fopen((std::string(argv[1]) + ".acc.bin").c_str(), "wb");
			if (!out) throw std::runtime_error("cannot write the .acc.bin file");
			std::vector<double> g(len); bool eoi = false; int calls = 0;
			while (!eoi) {
				double Zx = 0.0; QNUInt32 cnt = 0;
				const double num = gaccum.accumulateGradient(g.data(), &Zx, &cnt, &eoi);
				const double head[4] = {(double)cnt, eoi ? 1.0 : 0.0, num, Zx};
				std::fwrite(head, sizeof(double), 4, out); std::fwrite(g.data(), sizeof(double), len, out);
				calls++;
			}
			std::fclose(out);
			// one more call without a rewind must fail loudly (the reference prints "All feature streams are at the end!" and exits)
			bool threw = false;
			try { double Zx; QNUInt32 cnt; gaccum.accumulateGradient(g.data(), &Zx, &cnt, &eoi); } catch (const std::runtime_error&) { threw = true; }
			if (!threw) { std::printf("MISMATCH: accumulateGradient past the end of every stream did not throw\n"); return 1; }
			std::printf("host_selftest ok (%d minibatches)\n", calls);
			return 0;
		}
		if (sub == "resume") {
			// AdaGrad, 2 iterations: (A) one run; (B) one iteration, then a second trainer resumed from the state left in the model
			// (CRFTrain's init_iter / avg_weight_present / grad_sqr_acc_file resume, CRFTrain/src/Main.cpp:599-630); (C) resumed from the FILES
			const std::string wname = std::string(argv[1]) + ".resume.weights";
			auto run = [&](CRF_Model& m, int from, int to, int present) {
				CRF_MemFeatureStream tstrm(off, ftrs, labs, n_base);
				m.resumeFromMemory((QNUInt32)present, (QNUInt32)from);
				CRF_SGTrainer t(&m, &tstrm, wname.c_str());
				t.setMaxIters(to); t.setMinibatch(3, 2); t.setAdagrad(true, 0.05, 1e-6); t.setGaussVar(50.0f);
				t.train();
				return t.presentations;
			};
			std::vector<double> lam0(lam);
			my_crf.setLambda(lam0.data(), len);
			std::fill(my_crf.getLambdaAcc(), my_crf.getLambdaAcc() + len, 0.0); std::fill(my_crf.getGradSqrAcc(), my_crf.getGradSqrAcc() + len, 0.0);
			run(my_crf, 0, 2, 0);
			std::vector<double> A(my_crf.getLambda(), my_crf.getLambda() + len), Aacc(my_crf.getLambdaAcc(), my_crf.getLambdaAcc() + len);
			my_crf.setLambda(lam0.data(), len);
			std::fill(my_crf.getLambdaAcc(), my_crf.getLambdaAcc() + len, 0.0); std::fill(my_crf.getGradSqrAcc(), my_crf.getGradSqrAcc() + len, 0.0);
			const int present = run(my_crf, 0, 1, 0);
			if (present != (int)n_utt) { std::printf("MISMATCH presentations %d\n", present); bad++; }
			{ FILE* fp = std::fopen((wname + ".i0.gradSqrAcc.out").c_str(), "r"); if (!fp) { std::printf("MISMATCH missing .i0.gradSqrAcc.out\n"); bad++; } else std::fclose(fp); }
			run(my_crf, 1, 2, present);
			double lmax = 0.0;
			for (QNUInt32 i = 0; i < len; i++) lmax = std::fmax(lmax, std::fabs(A[i]));
			for (QNUInt32 i = 0; i < len; i++) {
				if (std::fabs(my_crf.getLambda()[i] - A[i]) > 1e-13 * lmax) { if (bad < 5) std::printf("MISMATCH resumed lambda[%u]: %.15g vs %.15g\n", i, my_crf.getLambda()[i], A[i]); bad++; }
				if (std::fabs(my_crf.getLambdaAcc()[i] - Aacc[i]) > 1e-12 * std::fmax(1.0, std::fabs(Aacc[i]))) { if (bad < 5) std::printf("MISMATCH resumed lambdaAcc[%u]\n", i); bad++; }
			}
			// (C) through the 6-digit ASCII files of iteration 0, as CRFTrain resumes: same trajectory up to the files' rounding
			CRF_Model m2(n_labs);
			m2.setLabMaxDur(max_dur); m2.setNActualLabs(n_act); m2.setModelType((modeltype)mt);
			m2.setFeatureMap(fmap_config(n_labs, n_states, width, max_dur, n_act), n_base, segf != 0);
			if (!m2.readFromFile((wname + ".i0.out").c_str()) || !m2.readAverageFromFile((wname + ".i0.avg.out").c_str(), present) ||
			    !m2.readGradSqrAccFromFile((wname + ".i0.gradSqrAcc.out").c_str())) throw std::runtime_error("cannot read the iteration-0 files back");
			m2.setInitIter(1);
			{
				CRF_MemFeatureStream tstrm(off, ftrs, labs, n_base);
				CRF_SGTrainer t(&m2, &tstrm, wname.c_str());
				t.setMaxIters(2); t.setMinibatch(3, 2); t.setAdagrad(true, 0.05, 1e-6); t.setGaussVar(50.0f);
				t.train();
			}
			for (QNUInt32 i = 0; i < len; i++)
				if (std::fabs(m2.getLambda()[i] - A[i]) > 1e-3 * lmax) { if (bad < 5) std::printf("MISMATCH file-resumed lambda[%u]: %.9g vs %.9g\n", i, m2.getLambda()[i], A[i]); bad++; }
			const std::string wdir = wname.find_last_of('/') == std::string::npos ? "." : wname.substr(0, wname.find_last_of('/'));
			for (const std::string& f : {wname + ".i0.out", wname + ".i0.avg.out", wname + ".i0.gradSqrAcc.out", wdir + "/.done.train.i0", wname + ".i1.out", wname + ".i1.avg.out",
			                             wname + ".i1.gradSqrAcc.out", wdir + "/.done.train.i1", wname, wname + ".avg.out", wdir + "/.done.train"}) std::remove(f.c_str());
			if (bad) { std::printf("host_selftest: %d mismatches\n", bad); return 1; }
			std::printf("host_selftest ok\n");
			return 0;
		}
		if (mode == 0) {
			// ---- minibatch seam ----
			CRF_MemFeatureStream strm(off, ftrs, labs, n_base);
			CRF_Minibatch_GradAccumulator gaccum(&my_crf, &strm, 1);
			gaccum.setMinibatch(n_utt);
			std::vector<double> g(len, 0.0); double Zx = 0.0; QNUInt32 cnt = 0; bool eoi = false;
			gaccum.rewindAllAndNextSegs();
			const double num = gaccum.accumulateGradient(g.data(), &Zx, &cnt, &eoi);
			double zsum = 0, nsum = 0, gmax = 0;
			for (QNUInt32 u = 0; u < n_utt; u++) { zsum += logZ[u]; nsum += numer[u]; }
			for (double v : grad) gmax = std::fmax(gmax, std::fabs(v));
			close(Zx, zsum, 1e-5, "sum logZ"); close(num, nsum, 1e-5, "sum numerator");
			if (cnt != n_utt || !eoi) { std::printf("MISMATCH uttCount %u eoi %d\n", cnt, (int)eoi); bad++; }
			for (QNUInt32 i = 0; i < len; i++)
				if (std::fabs(g[i] - grad[i]) > 1e-4 * gmax + 1e-4 * std::fabs(grad[i])) { if (bad < 5) std::printf("MISMATCH grad[%u]: %.9g vs %.9g\n", i, g[i], grad[i]); bad++; }
			// ---- per-utterance seam (CRF_GradBuilder::buildGradient accumulates into grad) ----
			std::unique_ptr<CRF_GradBuilder> gb(CRF_GradBuilder::create(&my_crf, EXPF));
			std::vector<double> g2(len, 0.0);
			strm.rewind();
			QNUInt32 u = 0;
			while (strm.nextseg() != QN_SEGID_BAD) {
				double z = 0.0; const double nu = gb->buildGradient(&strm, g2.data(), &z);
				close(z, logZ[u], 1e-5, "logZ"); close(nu, numer[u], 1e-5, "numerator"); u++;
			}
			for (QNUInt32 i = 0; i < len; i++)
				if (std::fabs(g2[i] - grad[i]) > 1e-4 * gmax + 1e-4 * std::fabs(grad[i])) { if (bad < 5) std::printf("MISMATCH grad2[%u]\n", i); bad++; }
			// ---- one SGD step as CRF_SGTrainer.cpp:309 and the lossy ASCII checkpoint round trip ----
			const float lr = 0.008f;
			for (QNUInt32 i = 0; i < len; i++) my_crf.getLambda()[i] += lr * g[i];
			std::string path = std::string(argv[1]) + ".weights.out";
			if (!my_crf.writeToFile(path.c_str())) throw std::runtime_error("writeToFile failed");
			std::vector<double> before(my_crf.getLambda(), my_crf.getLambda() + len);
			if (!my_crf.readFromFile(path.c_str())) throw std::runtime_error("readFromFile failed");
			for (QNUInt32 i = 0; i < len; i++) close(my_crf.getLambda()[i], before[i], 1e-5, "checkpoint round trip");
			std::remove(path.c_str());
			// ---- CRF_SGTrainer::train() with lambda resident on the device: 2 iterations, minibatches of 2 utterances, against the same
			//      loop run through the host seam (accumulateGradient + `lambda += lr * grad`, CRF_SGTrainer.cpp:309) ----
			std::vector<double> lam0(len);
			for (QNUInt32 i = 0; i < len; i++) lam0[i] = lam[i];
			my_crf.setLambda(lam0.data(), len);
			const std::string wname = std::string(argv[1]) + ".sgd.weights";
			CRF_MemFeatureStream tstrm(off, ftrs, labs, n_base);
			CRF_SGTrainer trainer(&my_crf, &tstrm, wname.c_str());
			trainer.setLR(0.01f); trainer.setLRDecayRate(0.5f); trainer.setMaxIters(2); trainer.setMinibatch(2, 2);
			trainer.train();
			std::vector<double> dev_lam(my_crf.getLambda(), my_crf.getLambda() + len);
			my_crf.setLambda(lam0.data(), len);
			CRF_MemFeatureStream hstrm(off, ftrs, labs, n_base);
			CRF_Minibatch_GradAccumulator hacc(&my_crf, &hstrm, 2);
			hacc.setMinibatch(2);
			float hlr = 0.01f; double lmax = 0.0;
			for (int it = 0; it < 2; it++) {
				bool end = false;
				hacc.rewindAllAndNextSegs();
				while (!end) {
					double z = 0.0; QNUInt32 c2 = 0;
					hacc.accumulateGradient(g.data(), &z, &c2, &end);
					if (!c2) break;
					for (QNUInt32 i = 0; i < len; i++) my_crf.getLambda()[i] += hlr * g[i];
				}
				hlr *= 0.5f;
			}
			for (QNUInt32 i = 0; i < len; i++) lmax = std::fmax(lmax, std::fabs(my_crf.getLambda()[i]));
			for (QNUInt32 i = 0; i < len; i++)
				if (std::fabs(dev_lam[i] - my_crf.getLambda()[i]) > 1e-9 * lmax) { if (bad < 5) std::printf("MISMATCH sgd lambda[%u]: %.12g vs %.12g\n", i, dev_lam[i], my_crf.getLambda()[i]); bad++; }
			// iterations are numbered from init_iter = 0 (CRF_SGTrainer.cpp:86,203); done markers live in the weight file's directory
			const std::string wdir = wname.find_last_of('/') == std::string::npos ? "." : wname.substr(0, wname.find_last_of('/'));
			for (const std::string& f : {wname + ".i0.out", wname + ".i0.avg.out", wdir + "/.done.train.i0", wname + ".i1.out", wname + ".i1.avg.out",
			                             wdir + "/.done.train.i1", wname, wname + ".avg.out", wdir + "/.done.train"}) {
				FILE* fp = std::fopen(f.c_str(), "r");
				if (!fp) { std::printf("MISMATCH missing trainer file %s\n", f.c_str()); bad++; } else { std::fclose(fp); std::remove(f.c_str()); }
			}
			if (trainer.iterLogLi.size() != 2 || !(trainer.iterLogLi[1] > trainer.iterLogLi[0])) { std::printf("MISMATCH log-likelihood did not improve\n"); bad++; }
		}
		if (mode == 1) {
			// ---- decode seam: first utterance, free-phone LM, beam 0 ----
			CRF_MemFeatureStream strm(off, ftrs, std::vector<QNUInt32>(), n_base);
			strm.nextseg();
			CRF_ViterbiDecoder_StdSeg_NoSegTransFtr vd(&strm, &my_crf);
			std::vector<CRF_BestPathArc> path; float cost = 0.0f;
			const int frames = vd.nStateDecode(&path, &cost);
			if (frames != (int)(off[1] - off[0]) || path.size() != n_arcs) { std::printf("MISMATCH decode: %d frames, %zu arcs\n", frames, path.size()); bad++; }
			else for (QNUInt32 k = 0; k < n_arcs; k++)
				if (path[k].ilabel != arcs[3 * k] || path[k].olabel != arcs[3 * k + 1] || (int)path[k].dur != arcs[3 * k + 2]) { std::printf("MISMATCH arc %u\n", k); bad++; }
			// ---- LM seam (one state per phone): nStateDecode(result, lm_fst, ...).  A bigram LM whose every weight is 0 is the free-phone LM
			//      with explicit weights: same arcs and cost; a prohibitive cost on every arc INTO the first phone of the free path (and on
			//      starting with it) must change the path; the LM belongs to the call, so a plain decode afterwards is the free one again ----
			if (n_states == 1 && n_labs > 2) {
				CRF_PhoneBigramLm lm; lm.start.assign(n_labs, 0.0f); lm.bigram.assign((size_t)n_labs * n_labs, 0.0f); lm.final_wt.assign(n_labs, 0.0f);
				std::vector<CRF_BestPathArc> p0; float c0 = 0.0f;
				strm.rewind(); strm.nextseg();
				vd.nStateDecode(&p0, &lm, &c0);
				if (p0.size() != path.size() || c0 != cost) { std::printf("MISMATCH zero-weight LM: %zu arcs cost %.9g vs %zu arcs cost %.9g\n", p0.size(), c0, path.size(), cost); bad++; }
				const int first = path[0].ilabel - 1;
				lm.start[first] = 5000.0f;
				for (QNUInt32 q = 0; q < n_labs; q++) lm.bigram[(size_t)q * n_labs + first] = 5000.0f;
				std::vector<CRF_BestPathArc> p1; float c1 = 0.0f;
				strm.rewind(); strm.nextseg();
				vd.nStateDecode(&p1, &lm, &c1);
				bool uses = false;
				for (const CRF_BestPathArc& a : p1) uses = uses || a.ilabel - 1 == first;
				if (uses || !(c1 > cost)) { std::printf("MISMATCH penalised LM: phone %d still on the path or cost %.9g not above %.9g\n", first, c1, cost); bad++; }
				std::vector<CRF_BestPathArc> p2; float c2 = 0.0f;
				strm.rewind(); strm.nextseg();
				vd.nStateDecode(&p2, &c2);
				if (p2.size() != path.size() || c2 != cost) { std::printf("MISMATCH the LM outlived its call\n"); bad++; }
				// beam: a huge beam prunes nothing, a tiny one can only lose paths, beam 0 afterwards is the unpruned decode again
				std::vector<CRF_BestPathArc> p3; float c3 = 0.0f, c4 = 0.0f, c5 = 0.0f;
				strm.rewind(); strm.nextseg(); vd.nStateDecode(&p3, &c3, 1e6);
				if (p3.size() != path.size() || c3 != cost) { std::printf("MISMATCH beam 1e6 changed the path\n"); bad++; }
				strm.rewind(); strm.nextseg(); p3.clear(); vd.nStateDecode(&p3, &c4, 1e-3);
				if (c4 < cost) { std::printf("MISMATCH a pruned decode is cheaper than the full one\n"); bad++; }
				strm.rewind(); strm.nextseg(); p3.clear(); vd.nStateDecode(&p3, &c5);
				if (c5 != cost) { std::printf("MISMATCH the beam outlived its call\n"); bad++; }
			}
			// ---- the whole stream as device batches of 3 utterances: utterance 0 must decode as above, the stream must end where it ends ----
			strm.rewind(); strm.nextseg();
			std::vector<std::vector<CRF_BestPathArc>> all; bool end = false; size_t done = 0;
			while (!end) {
				std::vector<std::vector<CRF_BestPathArc>> res; std::vector<float> costs; std::vector<int> nfr;
				done += vd.nStateDecodeBatch(3, &res, &costs, &nfr, &end);
				for (size_t u = 0; u < res.size(); u++) {
					if (nfr[u] != (int)(off[all.size() + 1] - off[all.size()])) { std::printf("MISMATCH batch decode frame count\n"); bad++; }
					if (all.empty() && (res[u].size() != path.size() || costs[u] != cost)) { std::printf("MISMATCH batch decode of utterance 0\n"); bad++; }
					all.push_back(res[u]);
				}
			}
			if (done != n_utt) { std::printf("MISMATCH batch decode: %zu of %u utterances\n", done, n_utt); bad++; }
			for (size_t k = 0; k < path.size() && k < all[0].size(); k++)
				if (all[0][k].ilabel != path[k].ilabel || all[0][k].olabel != path[k].olabel || all[0][k].dur != path[k].dur) { std::printf("MISMATCH batch arc %zu\n", k); bad++; }
		}
		if (bad) { std::printf("host_selftest: %d mismatches\n", bad); return 1; }
		std::printf("host_selftest ok\n");
		return 0;
	} catch (const std::exception& e) {
		std::fprintf(stderr, "host_selftest: %s\n", e.what());
		return 3;
	}
}
