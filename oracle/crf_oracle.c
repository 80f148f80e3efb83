/*
 * TEST INFRASTRUCTURE ONLY -- see crf_oracle.h.  Plain-C (fp64) restatement of the ASR-CRaFT
 * CRF lattice hot path.  Compiled with -ffp-contract=off so that `acc += x*w` is a separate
 * multiply and add, as in the reference's x86-64 -O2 build (no FMA on baseline x86-64).
 * All citations are relative to /root/reference/.
 */
#include "crf_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define LOG0 (-DBL_MAX) /* CRF/src/utils/CRF_LogMath.h:26 */

static __thread char g_err[512];
const char* crforacle_last_error(void) { return g_err; }
#define FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return 1; } while (0)

/* ------------------------------------------------------------------------------------------
 * log-space sum, CRF/src/utils/CRF_LogMath.cpp:69-125 (two-pass: max, sum of exp, log)
 * ---------------------------------------------------------------------------------------- */
static double log_add_n(const double* r, int n) {
	double mx = r[0];
	for (int i = 1; i < n; i++) if (r[i] > mx) mx = r[i];
	double s = 0.0;
	for (int i = 0; i < n; i++) s += exp(r[i] - mx);
	return mx + log(s);
}
/* CRF/src/utils/CRF_LogMath.cpp:41-64 */
static double log_add2(double a, double b) {
	double x = a, y = b;
	if (y > x) { y = a; x = b; }
	return x + log(1.0 + exp(y - x));
}

/* ------------------------------------------------------------------------------------------
 * lambda layout: CRF_StdFeatureMap::recalc / computeStateFeatureIdx / computeTransFeatureIdx
 * (CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:472-517, 280-320, 355-410)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
	uint32_t L, nAct, nStates;
	uint32_t nSf, nTf;     /* feature counts without bias */
	uint32_t nS, nT;       /* function counts incl. bias */
	uint32_t len;
	uint32_t* sidx;        /* [L] */
	uint32_t* tidx;        /* [L*L], [p*L+c] */
} fmap_t;

static int fmap_build(const crforacle_config* c, fmap_t* m) {
	memset(m, 0, sizeof *m);
	if (c->n_states == 0 || c->n_labs % c->n_states != 0) FAIL("invalid state/label combination");
	m->L = c->n_labs; m->nStates = c->n_states; m->nAct = c->n_labs / c->n_states;
	m->nSf = c->use_state_ftrs ? c->state_fidx_end - c->state_fidx_start + 1 : 0;
	m->nTf = c->use_trans_ftrs ? c->trans_fidx_end - c->trans_fidx_start + 1 : 0;
	m->nS = m->nSf + (c->use_state_bias ? 1 : 0);
	m->nT = m->nTf + (c->use_trans_bias ? 1 : 0);
	uint64_t transMult = (c->n_states == 1) ? (uint64_t)m->nAct * m->nAct
	                                        : (uint64_t)m->nAct * m->nAct + m->L + m->L - m->nAct;
	uint64_t len = (uint64_t)m->nS * m->L + (uint64_t)m->nT * transMult;
	if (len > 0xfffffff0u) FAIL("lambda vector too long");
	m->len = (uint32_t)len;
	m->sidx = (uint32_t*)malloc(sizeof(uint32_t) * m->L);
	m->tidx = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)m->L * m->L);
	uint32_t L = m->L, N = m->nStates;
	uint32_t run = 0;
	for (uint32_t cl = 0; cl < L; cl++) {
		if (N == 1) m->sidx[cl] = cl * (m->nS + L * m->nT);
		else {
			m->sidx[cl] = run;
			run += m->nS + ((cl % N == 0) ? (m->nAct + 1) * m->nT : 2 * m->nT);
		}
	}
	for (uint32_t cl = 0; cl < L; cl++)
		for (uint32_t p = 0; p < L; p++) {
			uint32_t v;
			if (N == 1) v = cl * (m->nS + L * m->nT) + m->nS + p * m->nT;
			else {
				v = m->sidx[cl] + m->nS;
				if (p != cl) {
					v += m->nT;
					if (cl % N == 0) {
						if ((p + 1) % N != 0) v = CRFO_NO_IDX;
						else v += (p / N) * m->nT;
					} else if (p != cl - 1) v = CRFO_NO_IDX;
				}
			}
			m->tidx[(size_t)p * L + cl] = v;
		}
	return 0;
}
static void fmap_free(fmap_t* m) { free(m->sidx); free(m->tidx); }

/* CRF_StdFeatureMap::computeStateArrayValue, CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:65-81 */
static double state_value(const crforacle_config* c, const fmap_t* m, const float* x, const double* lam, uint32_t cl) {
	double v = 0.0;
	uint32_t lc = m->sidx[cl];
	if (c->use_state_ftrs)
		for (uint32_t f = c->state_fidx_start; f <= c->state_fidx_end; f++) { v += x[f] * lam[lc]; lc++; }
	if (c->use_state_bias) { v += lam[lc] * c->state_bias_val; lc++; }
	return v;
}
/* CRF_StdFeatureMap::computeTransMatrixValue, CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:94-110 */
static double trans_value(const crforacle_config* c, const fmap_t* m, const float* x, const double* lam, uint32_t p, uint32_t cl) {
	double v = 0.0;
	uint32_t lc = m->tidx[(size_t)p * m->L + cl];
	if (c->use_trans_ftrs)
		for (uint32_t f = c->trans_fidx_start; f <= c->trans_fidx_end; f++) { v += x[f] * lam[lc]; lc++; }
	if (c->use_trans_bias) { v += lam[lc] * c->trans_bias_val; lc++; }
	return v;
}
/* CRF_StdFeatureMap::computeStateExpF, CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:130-175 */
static double state_expf(const crforacle_config* c, const fmap_t* m, const float* x, const double* lam,
                         double* ExpF, double* grad, double ab, uint32_t t_clab, uint32_t cl) {
	double ll = 0.0;
	uint32_t lc = m->sidx[cl];
	int match = (t_clab == cl);
	if (c->use_state_ftrs)
		for (uint32_t f = c->state_fidx_start; f <= c->state_fidx_end; f++) {
			ExpF[lc] += ab * x[f];
			if (match) { grad[lc] += x[f]; ll += lam[lc] * x[f]; }
			lc++;
		}
	if (c->use_state_bias) {
		ExpF[lc] += ab * c->state_bias_val;
		if (match) { grad[lc] += c->state_bias_val; ll += lam[lc] * c->state_bias_val; }
		lc++;
	}
	return ll;
}
/* CRF_StdFeatureMap::computeTransExpF, CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:197-223 */
static double trans_expf(const crforacle_config* c, const fmap_t* m, const float* x, const double* lam,
                         double* ExpF, double* grad, double ab, uint32_t t_plab, uint32_t t_clab, uint32_t p, uint32_t cl) {
	double ll = 0.0;
	uint32_t lc = m->tidx[(size_t)p * m->L + cl];
	int match = (cl == t_clab) && (p == t_plab);
	if (c->use_trans_ftrs)
		for (uint32_t f = c->trans_fidx_start; f <= c->trans_fidx_end; f++) {
			ExpF[lc] += ab * x[f];
			if (match) { grad[lc] += x[f]; ll += lam[lc] * x[f]; }
			lc++;
		}
	if (c->use_trans_bias) {
		ExpF[lc] += ab * c->trans_bias_val;
		if (match) { grad[lc] += c->trans_bias_val; ll += lam[lc] * c->trans_bias_val; }
		lc++;
	}
	return ll;
}

/* ------------------------------------------------------------------------------------------
 * window streams
 * ---------------------------------------------------------------------------------------- */
/* width of one stream's window vector: CRF_InFtrStream_SeqMultiWindow ctor, CRF/src/io/CRF_InFtrStream_SeqMultiWindow.cpp:47-117 */
static uint32_t part_width(uint32_t F, uint32_t D, uint32_t seg, uint32_t lc, uint32_t rc, uint32_t bdelta) {
	if (D == 1) return (lc + 1 + rc) * F;
	if (seg) return 8 * F + D + (lc + rc) * F;
	if (bdelta) { uint32_t both = lc < rc + 1 ? lc : rc + 1; return both * F; }
	return (lc + 1 + rc) * F;
}
uint32_t crforacle_window_width(const crforacle_config* c) {
	uint32_t w = part_width(c->n_base_ftrs, c->max_dur, c->extract_seg_ftrs, c->left_ctx, c->right_ctx, c->boundary_delta);
	/* QN_InFtrStream_JoinFtrs (QuickNet3, version unpinned by the reference): the frames of the second stream behind those of the first */
	if (c->n_base_ftrs2) w += part_width(c->n_base_ftrs2, c->max_dur, c->extract_seg_ftrs2, c->left_ctx2, c->right_ctx2, c->boundary_delta2);
	return w;
}
static int has_ext(const crforacle_config* c) { return c->left_ctx || c->right_ctx || c->boundary_delta || c->n_base_ftrs2; }

/* One (t,d) window of one stream.  x points at the utterance's frames of that stream, [lc + T + rc][F]: labelled frame t is row lc + t
 * (nextseg: cur_line = top_margin + left_context_len, :163-199); the window covers labelled frames t-d+1..t.
 * CRF/src/io/CRF_InFtrStream_SeqMultiWindow.cpp: read_ftrs :325-471 = [left context of the FIRST frame | body | right context], body =
 * sample_ftrs :556-607 | avg_ftrs :609-655 | max_ftrs :657-706 | min_ftrs :708-757 | dur_ftrs :850-895 (segment features) or
 * first_frame_ftrs :1017-1048; the right context hangs on the LAST frame with segment features (last_frame_right_ctx_ftrs :977-1015)
 * and on the first frame otherwise (first_frame_right_ctx_ftrs :937-975); boundary_delta_ftrs :1050-1110 replaces all of it. */
static void window_part(uint32_t F, uint32_t D, uint32_t seg, uint32_t lc, uint32_t rc, uint32_t bdelta,
                        const float* x, uint32_t t, uint32_t d, float* out) {
	const float* first = x + (size_t)(lc + t - d + 1) * F;     /* first frame of the window */
	const float* last = x + (size_t)(lc + t) * F;
	if (bdelta && !(D > 1 && seg)) {
		uint32_t both = lc < rc + 1 ? lc : rc + 1;
		for (uint32_t i = 0; i < both; i++) {
			const float* l = first - (size_t)(i + 1) * F; const float* r = first + (size_t)i * F;
			for (uint32_t f = 0; f < F; f++) out[(size_t)i * F + f] = l[f] >= r[f] ? l[f] - r[f] : r[f] - l[f];
		}
		return;
	}
	memcpy(out, first - (size_t)lc * F, (size_t)lc * F * sizeof(float));      /* first_frame_left_ctx_ftrs :897-935 */
	out += (size_t)lc * F;
	if (D == 1 || !seg) {
		memcpy(out, first, F * sizeof(float)); out += F;
		memcpy(out, first + F, (size_t)rc * F * sizeof(float));
		return;
	}
	/* 5 sampled frames at the 10/30/50/70/90 % points; float arithmetic as in the reference */
	float one_tenth = d * 0.1;
	for (int i = 1, k = 0; i < 10; i += 2, k++) {
		uint32_t step = (uint32_t)ceil(one_tenth * i) - 1;
		memcpy(out + (size_t)k * F, first + (size_t)step * F, F * sizeof(float));
	}
	float* avg = out + 5 * (size_t)F; float* mx = avg + F; float* mn = mx + F; float* du = mn + F;
	/* running statistics accumulate from the LAST frame of the window back to the first */
	for (uint32_t f = 0; f < F; f++) {
		const float* p = last + f;
		float acc = 0.0f, amax = *p, amin = *p;
		for (uint32_t k = 1; k <= d; k++) {
			float v = *p;
			acc += v;
			if (v > amax) amax = v;
			if (v < amin) amin = v;
			p -= F;
		}
		avg[f] = acc / d; mx[f] = amax; mn[f] = amin;
	}
	for (uint32_t k = 1; k <= D; k++) du[k - 1] = (k == d) ? 1.0f : 0.0f;
	memcpy(du + D, last + F, (size_t)rc * F * sizeof(float));
}

int crforacle_window_ftrs2(const crforacle_config* c, uint32_t T, const float* base, const float* base2, float* out) {
	uint32_t W = crforacle_window_width(c), D = c->max_dur;
	uint32_t W1 = part_width(c->n_base_ftrs, D, c->extract_seg_ftrs, c->left_ctx, c->right_ctx, c->boundary_delta);
	if (c->n_base_ftrs2 && !base2) FAIL("the configuration joins a second feature stream but none was passed");
	for (uint32_t t = 0; t < T; t++)
		for (uint32_t d = 1; d <= D && d <= t + 1; d++) {
			float* o = out + ((size_t)t * D + (d - 1)) * W;
			window_part(c->n_base_ftrs, D, c->extract_seg_ftrs, c->left_ctx, c->right_ctx, c->boundary_delta, base, t, d, o);
			if (c->n_base_ftrs2) window_part(c->n_base_ftrs2, D, c->extract_seg_ftrs2, c->left_ctx2, c->right_ctx2, c->boundary_delta2, base2, t, d, o + W1);
		}
	return 0;
}
int crforacle_window_ftrs(const crforacle_config* c, uint32_t T, const float* base, float* out) { return crforacle_window_ftrs2(c, T, base, NULL, out); }

/* CRF_InLabStream_SeqMultiWindow::groupLabels / nextseg / read_labs,
 * CRF/src/io/CRF_InLabStream_SeqMultiWindow.cpp:51-110, 119-196, 246-306 */
int crforacle_window_labs(const crforacle_config* c, uint32_t T, const uint32_t* labs, uint32_t* out4) {
	uint32_t D = c->max_dur;
	for (size_t i = 0; i < (size_t)T * 4; i++) out4[i] = CRFO_LAB_BAD;
	uint32_t s = 0;
	while (s < T) {
		uint32_t e = s;
		while (e + 1 < T && labs[e + 1] == labs[s]) e++;
		uint32_t lab = labs[s], dur = e - s + 1;
		if (lab != CRFO_LAB_BAD) {   /* runs of CRF_LAB_BAD are never grouped (:151-157) */
			if (dur <= D) {
				uint32_t* o = out4 + 4 * (size_t)e; o[0] = lab; o[1] = s; o[2] = e; o[3] = 0;
			} else {
				uint32_t np = (dur % D == 0) ? dur / D : dur / D + 1;
				uint32_t pd = dur / np, rem = dur % np, ps = s;
				for (uint32_t r = 0; r < np; r++) {
					uint32_t len = pd + (r < rem ? 1 : 0), pe = ps + len - 1;
					uint32_t* o = out4 + 4 * (size_t)pe; o[0] = lab; o[1] = ps; o[2] = pe; o[3] = 1;
					ps = pe + 1;
				}
			}
		}
		s = e + 1;
	}
	return 0;
}

int crforacle_lambda_len(const crforacle_config* c, uint32_t* out) {
	fmap_t m; if (fmap_build(c, &m)) return 1;
	*out = m.len; fmap_free(&m); return 0;
}
int crforacle_index_maps(const crforacle_config* c, uint32_t* sidx, uint32_t* tidx) {
	fmap_t m; if (fmap_build(c, &m)) return 1;
	memcpy(sidx, m.sidx, sizeof(uint32_t) * m.L);
	memcpy(tidx, m.tidx, sizeof(uint32_t) * (size_t)m.L * m.L);
	fmap_free(&m); return 0;
}

/* ------------------------------------------------------------------------------------------
 * forward-backward + gradient for one utterance
 * ---------------------------------------------------------------------------------------- */
typedef struct {
	const crforacle_config* c; const fmap_t* m; const double* lam;
	double* ExpF;     /* [len] scratch */
	double* Mconst;   /* [L*L] transition scores when they do not depend on the frame, else NULL */
	const float* x2;  /* the utterance's frames of the joined second stream (or NULL) */
} ctx_t;

/* Frame-level, 1 state/label: CRF_NewGradBuilder::buildGradient (CRF/src/trainers/gradbuilders/CRF_NewGradBuilder.cpp:48-382)
 * over CRF_StdStateNode (CRF/src/nodes/CRF_StdStateNode.cpp:58-299). */
static int fb_frame_1state(ctx_t* k, uint32_t T, const float* x, const uint32_t* lab4,
                           double* grad, double* numer, double* logZ, double* A_out, double* B_out) {
	const crforacle_config* c = k->c; const fmap_t* m = k->m; const double* lam = k->lam;
	uint32_t L = m->L, F = crforacle_window_width(c);   /* = n_base_ftrs without context frames / joined streams; else x is the expanded window matrix (fb_one) */
	double* S = (double*)malloc(sizeof(double) * (size_t)T * L);
	double* A = (double*)malloc(sizeof(double) * (size_t)T * L);
	double* B = (double*)malloc(sizeof(double) * (size_t)T * L);
	double* Mt = k->Mconst ? NULL : (double*)malloc(sizeof(double) * (size_t)T * L * L);
	double* acc = (double*)malloc(sizeof(double) * L);
	double* tmpB = (double*)malloc(sizeof(double) * L);
	memset(k->ExpF, 0, sizeof(double) * m->len);
	for (uint32_t t = 0; t < T; t++) {
		const float* xt = x + (size_t)t * F;
		const double* M = k->Mconst;
		if (!M) {
			double* Mw = Mt + (size_t)t * L * L;
			for (uint32_t cl = 0; cl < L; cl++) for (uint32_t p = 0; p < L; p++) Mw[(size_t)p * L + cl] = trans_value(c, m, xt, lam, p, cl);
			M = Mw;
		}
		for (uint32_t cl = 0; cl < L; cl++) S[(size_t)t * L + cl] = state_value(c, m, xt, lam, cl);
		for (uint32_t cl = 0; cl < L; cl++) {
			if (t == 0) { A[cl] = S[cl]; continue; }
			for (uint32_t p = 0; p < L; p++) acc[p] = A[(size_t)(t - 1) * L + p] + M[(size_t)p * L + cl];
			A[(size_t)t * L + cl] = log_add_n(acc, (int)L) + S[(size_t)t * L + cl];
		}
	}
	double Zx = log_add_n(A + (size_t)(T - 1) * L, (int)L);
	double ll = 0.0;
	for (uint32_t t = T; t-- > 0;) {
		const float* xt = x + (size_t)t * F;
		double* Bt = B + (size_t)t * L;
		if (t == T - 1) for (uint32_t cl = 0; cl < L; cl++) Bt[cl] = 0.0;
		else {
			const double* M = k->Mconst ? k->Mconst : Mt + (size_t)(t + 1) * L * L;
			for (uint32_t cl = 0; cl < L; cl++) tmpB[cl] = B[(size_t)(t + 1) * L + cl] + S[(size_t)(t + 1) * L + cl];
			for (uint32_t p = 0; p < L; p++) {
				for (uint32_t cl = 0; cl < L; cl++) acc[cl] = M[(size_t)p * L + cl] + tmpB[cl];
				Bt[p] = log_add_n(acc, (int)L);
			}
		}
		uint32_t label = lab4[4 * (size_t)t];
		uint32_t prev_lab = t > 0 ? lab4[4 * (size_t)(t - 1)] : L + 1;
		const double* M = k->Mconst ? k->Mconst : Mt + (size_t)t * L * L;
		double tot = 0.0, ttot = 0.0;
		for (uint32_t cl = 0; cl < L; cl++) {
			double ab = exp(A[(size_t)t * L + cl] + Bt[cl] - Zx);
			tot += ab;
			ll += state_expf(c, m, xt, lam, k->ExpF, grad, ab, label, cl);
			if (t == 0) ttot = 1.0;
			else for (uint32_t p = 0; p < L; p++) {
				double xi = exp(A[(size_t)(t - 1) * L + p] + M[(size_t)p * L + cl] + S[(size_t)t * L + cl] + Bt[cl] - Zx);
				ttot += xi;
				ll += trans_expf(c, m, xt, lam, k->ExpF, grad, xi, prev_lab, label, p, cl);
			}
		}
		if (tot > 1.1 || tot < 0.9 || ttot > 1.1 || ttot < 0.9) {   /* CRF_StdStateNode.cpp:252-275 */
			free(S); free(A); free(B); free(Mt); free(acc); free(tmpB);
			FAIL("posterior mass check failed at frame %u: %g / %g", t, tot, ttot);
		}
	}
	for (uint32_t i = 0; i < m->len; i++) grad[i] -= k->ExpF[i];
	*numer = ll; *logZ = Zx;
	if (A_out) memcpy(A_out, A, sizeof(double) * (size_t)T * L);
	if (B_out) memcpy(B_out, B, sizeof(double) * (size_t)T * L);
	free(S); free(A); free(B); free(Mt); free(acc); free(tmpB);
	return 0;
}

/* Frame-level, N states/label: CRF_StdNStateNode (CRF/src/nodes/CRF_StdNStateNode.cpp:65-409).
 * Transition scores are kept as diag[L], offDiag[L] (indexed by previous state) and dense[P*P]. */
static int fb_frame_nstate(ctx_t* k, uint32_t T, const float* x, const uint32_t* lab4,
                           double* grad, double* numer, double* logZ, double* A_out, double* B_out) {
	const crforacle_config* c = k->c; const fmap_t* m = k->m; const double* lam = k->lam;
	uint32_t L = m->L, F = crforacle_window_width(c), N = m->nStates, P = m->nAct;
	size_t per = (size_t)2 * L + (size_t)P * P;
	double* S = (double*)malloc(sizeof(double) * (size_t)T * L);
	double* A = (double*)malloc(sizeof(double) * (size_t)T * L);
	double* B = (double*)malloc(sizeof(double) * (size_t)T * L);
	double* TR = (double*)malloc(sizeof(double) * (size_t)T * per);
	double* acc = (double*)malloc(sizeof(double) * L);
	double* tmpB = (double*)malloc(sizeof(double) * L);
	memset(k->ExpF, 0, sizeof(double) * m->len);
	for (uint32_t t = 0; t < T; t++) {
		const float* xt = x + (size_t)t * F;
		double* diag = TR + (size_t)t * per; double* off = diag + L; double* dense = off + L;
		for (uint32_t cl = 0; cl < L; cl++) {
			S[(size_t)t * L + cl] = state_value(c, m, xt, lam, cl);
			diag[cl] = trans_value(c, m, xt, lam, cl, cl);
			if (cl % N == 0) for (uint32_t q = 0; q < P; q++) dense[(size_t)q * P + cl / N] = trans_value(c, m, xt, lam, q * N + N - 1, cl);
			else off[cl - 1] = trans_value(c, m, xt, lam, cl - 1, cl);
		}
		double* At = A + (size_t)t * L; const double* Ap = At - L;
		for (uint32_t cl = 0; cl < L; cl++) {
			if (t == 0) { At[cl] = S[cl]; continue; }
			double a = Ap[cl] + diag[cl];
			if (cl % N == 0) {
				for (uint32_t q = 0; q < P; q++) acc[q] = Ap[q * N + N - 1] + dense[(size_t)q * P + cl / N];
				a = log_add2(a, log_add_n(acc, (int)P));
			} else a = log_add2(a, Ap[cl - 1] + off[cl - 1]);
			At[cl] = a + S[(size_t)t * L + cl];
		}
	}
	double Zx = log_add_n(A + (size_t)(T - 1) * L, (int)L);
	double ll = 0.0;
	int bad = 0;
	for (uint32_t t = T; t-- > 0 && !bad;) {
		const float* xt = x + (size_t)t * F;
		double* Bt = B + (size_t)t * L;
		if (t == T - 1) for (uint32_t cl = 0; cl < L; cl++) Bt[cl] = 0.0;
		else {
			const double* diag = TR + (size_t)(t + 1) * per; const double* off = diag + L; const double* dense = off + L;
			for (uint32_t cl = 0; cl < L; cl++) tmpB[cl] = B[(size_t)(t + 1) * L + cl] + S[(size_t)(t + 1) * L + cl];
			for (uint32_t p = 0; p < L; p++) {
				double b = tmpB[p] + diag[p];
				if ((p + 1) % N == 0) {
					uint32_t ip = (p + 1) / N - 1;
					for (uint32_t q = 0; q < P; q++) acc[q] = dense[(size_t)ip * P + q] + tmpB[q * N];
					b = log_add2(b, log_add_n(acc, (int)P));
				} else b = log_add2(b, off[p] + tmpB[p + 1]);
				Bt[p] = b;
			}
		}
		uint32_t label = lab4[4 * (size_t)t];
		uint32_t prev_lab = t > 0 ? lab4[4 * (size_t)(t - 1)] : L + 1;
		const double* diag = TR + (size_t)t * per; const double* off = diag + L; const double* dense = off + L;
		const double* Ap = A + (size_t)(t ? t - 1 : 0) * L;
		double tot = 0.0, ttot = 0.0;
		for (uint32_t cl = 0; cl < L; cl++) {
			double sb = S[(size_t)t * L + cl] + Bt[cl] - Zx;
			double ab = exp(A[(size_t)t * L + cl] + Bt[cl] - Zx);
			tot += ab;
			ll += state_expf(c, m, xt, lam, k->ExpF, grad, ab, label, cl);
			if (t == 0) { ttot = 1.0; continue; }
			double xi = exp(Ap[cl] + diag[cl] + sb);
			ttot += xi; ll += trans_expf(c, m, xt, lam, k->ExpF, grad, xi, prev_lab, label, cl, cl);
			if (cl % N == 0) {
				for (uint32_t q = 0; q < P; q++) {
					uint32_t rp = q * N + N - 1;
					xi = exp(Ap[rp] + dense[(size_t)q * P + cl / N] + sb);
					ttot += xi; ll += trans_expf(c, m, xt, lam, k->ExpF, grad, xi, prev_lab, label, rp, cl);
				}
			} else {
				xi = exp(Ap[cl - 1] + off[cl - 1] + sb);
				ttot += xi; ll += trans_expf(c, m, xt, lam, k->ExpF, grad, xi, prev_lab, label, cl - 1, cl);
			}
		}
		if (tot > 1.1 || tot < 0.9 || ttot > 1.1 || ttot < 0.9) bad = 1;
	}
	if (!bad) {
		for (uint32_t i = 0; i < m->len; i++) grad[i] -= k->ExpF[i];
		*numer = ll; *logZ = Zx;
		if (A_out) memcpy(A_out, A, sizeof(double) * (size_t)T * L);
		if (B_out) memcpy(B_out, B, sizeof(double) * (size_t)T * L);
	}
	free(S); free(A); free(B); free(TR); free(acc); free(tmpB);
	if (bad) FAIL("posterior mass check failed (N-state)");
	return 0;
}

/* Segmental `stdseg`: CRF_NewGradBuilder_StdSeg::buildGradient (CRF/src/trainers/gradbuilders/CRF_NewGradBuilder_StdSeg.cpp:31-377)
 * over CRF_StdSegStateNode (CRF/src/nodes/CRF_StdSegStateNode.cpp:83-462).  Label id = (dur-1)*P + phone. */
static int fb_stdseg(ctx_t* k, uint32_t T, const float* x, const uint32_t* lab4,
                     double* grad, double* numer, double* logZ, double* A_out, double* B_out) {
	const crforacle_config* c = k->c; const fmap_t* m = k->m; const double* lam = k->lam;
	uint32_t L = m->L, D = c->max_dur, P = L / D, W = crforacle_window_width(c);
	if (c->n_states != 1) FAIL("stdseg with n_states>1 is not implemented by the reference (CRF_StateNode.cpp:497-502)");
	if (c->use_trans_ftrs && !k->Mconst) FAIL("oracle: stdseg with transition features not restated");
	const double* M = k->Mconst;
	float* X = (float*)malloc(sizeof(float) * (size_t)T * D * W);
	double* S = (double*)malloc(sizeof(double) * (size_t)T * L);
	double* A = (double*)malloc(sizeof(double) * (size_t)T * L);
	double* B = (double*)malloc(sizeof(double) * (size_t)T * L);
	double* acc = (double*)malloc(sizeof(double) * (L > D * P ? L : D * P));
	double* tmpB = (double*)malloc(sizeof(double) * L);
	uint32_t* nodeLab = (uint32_t*)malloc(sizeof(uint32_t) * T);
	crforacle_window_ftrs2(c, T, x, k->x2, X);
	memset(k->ExpF, 0, sizeof(double) * m->len);
	for (size_t i = 0; i < (size_t)T * L; i++) { A[i] = LOG0; B[i] = LOG0; }
#define AVAIL(t) (P * (((t) + 1 < D) ? (t) + 1 : D))
	for (uint32_t t = 0; t < T; t++) {
		/* node label, CRF_NewGradBuilder_StdSeg.cpp:172-187 */
		const uint32_t* l4 = lab4 + 4 * (size_t)t;
		nodeLab[t] = (l4[0] != CRFO_LAB_BAD) ? c->n_actual_labs * (l4[2] - l4[1]) + l4[0] : CRFO_LAB_BAD;
		uint32_t nodeMaxDur = t + 1 < D ? t + 1 : D, numPrev = t < D ? t : D;
		for (uint32_t d = 1; d <= nodeMaxDur; d++)
			for (uint32_t y = 0; y < P; y++) {
				uint32_t cl = (d - 1) * P + y;
				double s = state_value(c, m, X + ((size_t)t * D + (d - 1)) * W, lam, cl);
				S[(size_t)t * L + cl] = s;
				if (d <= numPrev) {
					uint32_t av = AVAIL(t - d);
					const double* Ap = A + (size_t)(t - d) * L;
					for (uint32_t p = 0; p < av; p++) acc[p] = Ap[p] + M[(size_t)p * L + cl];
					A[(size_t)t * L + cl] = log_add_n(acc, (int)av) + s;
				} else A[(size_t)t * L + cl] = s;
			}
	}
	double Zx = log_add_n(A + (size_t)(T - 1) * L, (int)AVAIL(T - 1));
	double ll = 0.0;
	int bad = 0;
	for (uint32_t t = T; t-- > 0 && !bad;) {
		uint32_t avail = AVAIL(t), numNext = (T - 1 - t) < D ? (T - 1 - t) : D;
		double* Bt = B + (size_t)t * L;
		if (numNext == 0) for (uint32_t cl = 0; cl < L; cl++) Bt[cl] = 0.0;   /* setTailBeta covers all nLabs */
		else {
			for (uint32_t d = 1; d <= numNext; d++)
				for (uint32_t y = 0; y < P; y++) {
					uint32_t nl = (d - 1) * P + y;
					tmpB[nl] = B[(size_t)(t + d) * L + nl] + S[(size_t)(t + d) * L + nl];
				}
			for (uint32_t cl = 0; cl < avail; cl++) {
				for (uint32_t nl = 0; nl < numNext * P; nl++) acc[nl] = M[(size_t)cl * L + nl] + tmpB[nl];
				Bt[cl] = log_add_n(acc, (int)(numNext * P));
			}
		}
		/* prev_lab: label of the nearest earlier node that ends a reference segment (:343-351) */
		uint32_t prev_lab = CRFO_LAB_BAD;
		for (uint32_t u = t; u > 0; u--) { prev_lab = nodeLab[u - 1]; if (prev_lab != CRFO_LAB_BAD) break; }
		uint32_t label = nodeLab[t];
		uint32_t nodeMaxDur = t + 1 < D ? t + 1 : D, numPrev = t < D ? t : D;
		double tot = 0.0, ttot = 0.0;
		for (uint32_t d = 1; d <= nodeMaxDur; d++) {
			const float* xs = X + ((size_t)t * D + (d - 1)) * W;
			for (uint32_t y = 0; y < P; y++) {
				uint32_t cl = (d - 1) * P + y;
				double ab = exp(A[(size_t)t * L + cl] + Bt[cl] - Zx);
				tot += ab;
				ll += state_expf(c, m, xs, lam, k->ExpF, grad, ab, label, cl);
				if (d <= numPrev) {
					uint32_t av = AVAIL(t - d);
					const double* Ap = A + (size_t)(t - d) * L;
					for (uint32_t p = 0; p < av; p++) {
						double xi = exp(Ap[p] + M[(size_t)p * L + cl] + S[(size_t)t * L + cl] + Bt[cl] - Zx);
						ttot += xi;
						ll += trans_expf(c, m, xs, lam, k->ExpF, grad, xi, prev_lab, label, p, cl);
					}
				}
			}
		}
		if (numPrev == 0) ttot = 1.0;
		if (tot > 1.000001 || tot < -0.000001 || ttot > 1.000001 || ttot < -0.000001) bad = 1;   /* CRF_StdSegStateNode.cpp:417-436 */
	}
#undef AVAIL
	if (!bad) {
		for (uint32_t i = 0; i < m->len; i++) grad[i] -= k->ExpF[i];
		*numer = ll; *logZ = Zx;
		if (A_out) memcpy(A_out, A, sizeof(double) * (size_t)T * L);
		if (B_out) memcpy(B_out, B, sizeof(double) * (size_t)T * L);
	}
	free(X); free(S); free(A); free(B); free(acc); free(tmpB); free(nodeLab);
	if (bad) FAIL("posterior mass check failed (stdseg)");
	return 0;
}

/* Segmental `stdseg_no_dur`, `stdseg_no_dur_no_transftr`, `stdseg_no_dur_no_segtransftr` with one state per phone and no transition
 * FEATURES (labels = phones; the duration lives in the window features only):
 * CRF_NewGradBuilder_StdSeg_NoDur_NoTrans::buildGradient (CRF/src/trainers/gradbuilders/CRF_NewGradBuilder_StdSeg_NoDur_NoTrans.cpp:65-492)
 * over CRF_StdSegStateNode_WithoutDurLab_WithoutSegTransFtr (CRF/src/nodes/CRF_StdSegStateNode_WithoutDurLab_WithoutSegTransFtr.cpp:
 * computeTransMatrix :39-121, computeAlpha :123-333, computeAlphaPlusTrans :1077-1116, computeBeta :395-614, computeExpF :616-1066),
 * in its native O(P^2 + D*P) form:
 *   A_t[y]      = logsum_y' (alpha_t[y'] + M[y',y])                       alpha_t[d,y] = S_t[d,y] + A_{t-d}[y]   (= S_t[d,y] when d == t+1)
 *   alpha_t[y]  = logsum_d alpha_t[d,y]                                    logZ = logsum_y alpha_{T-1}[y]
 *   B_t[y]      = logsum_d (S_{t+d}[d,y] + beta_{t+d}[y])                  beta_t[y'] = logsum_y (M[y',y] + B_t[y]),  beta_{T-1} = 0
 *   gamma_t[d,y] = exp(alpha_t[d,y] + beta_t[y] - logZ)                    xi_t[y',y]  = exp(alpha_t[y'] + M[y',y] + B_t[y] - logZ)
 * Without transition features the three node families give identical results and equal `stdseg` on the (duration, phone) label set
 * with tied weights (oracle/binding.py keeps that restatement; tests/test_oracle.py checks both against goldens made by the
 * reference's own no_dur node classes).  alpha/beta dumps: [T][P] (sum over durations). */
static int fb_nodur(ctx_t* k, uint32_t T, const float* x, const uint32_t* lab4,
                    double* grad, double* numer, double* logZ, double* A_out, double* B_out) {
	const crforacle_config* c = k->c; const fmap_t* m = k->m; const double* lam = k->lam;
	const uint32_t P = m->L, D = c->max_dur, W = crforacle_window_width(c);
	/* N states per phone (CRF_StdSegNStateNode_WithoutDurLab_WithoutSegTransFtr): every sub-state is a segment of its own and only the
	 * legal pairs of the N-state map exist -- self (diagTransMatrix), end state -> start state (denseTransMatrix) and k-1 -> k
	 * (offDiagTransMatrix), computeAlphaPlusTrans :1161-1268, computeBeta :450-718, computeExpF :719-1160 */
#define LEGAL(q, y) (m->tidx[(size_t)(q) * P + (y)] != CRFO_NO_IDX)
	/* transition FEATURES are restated for stdseg_no_dur_no_segtransftr only: M_t[y'][y] comes from the duration-1 window of the frame
	 * the new segment starts in (CRF_StdSegStateNode_WithoutDurLab_WithoutSegTransFtr::computeTransMatrix :39-121) */
	if (c->use_trans_ftrs && c->model_type != CRFO_STDSEG_NO_DUR_NO_SEGTRANSFTR) FAIL("oracle: transition features restated for stdseg_no_dur_no_segtransftr only");
	float* X = (float*)malloc(sizeof(float) * (size_t)T * D * W);
	double* S = (double*)malloc(sizeof(double) * (size_t)T * D * P);      /* [t][d-1][y] */
	double* AD = (double*)malloc(sizeof(double) * (size_t)T * D * P);     /* alpha_t[d,y] */
	double* A = (double*)malloc(sizeof(double) * (size_t)T * P);          /* alpha_t[y] */
	double* AT = (double*)malloc(sizeof(double) * (size_t)T * P);         /* A_t[y] */
	double* B = (double*)malloc(sizeof(double) * (size_t)T * P);          /* beta_t[y] */
	double* BV = (double*)malloc(sizeof(double) * (size_t)T * P);         /* B_t[y] */
	double* Mt = (k->Mconst || !c->use_trans_ftrs) ? NULL : (double*)malloc(sizeof(double) * (size_t)T * P * P);   /* M_t[y'][y] of the frame the segment starts in */
	double* acc = (double*)malloc(sizeof(double) * (P > D ? P : D));
	double* Mn = NULL;
	crforacle_window_ftrs2(c, T, x, k->x2, X);
	memset(k->ExpF, 0, sizeof(double) * m->len);
	if (!k->Mconst && !c->use_trans_ftrs) {   /* N-state, bias only: one matrix, illegal pairs never read */
		Mn = (double*)malloc(sizeof(double) * (size_t)P * P);
		for (uint32_t q = 0; q < P; q++) for (uint32_t y = 0; y < P; y++) Mn[(size_t)q * P + y] = LEGAL(q, y) ? trans_value(c, m, NULL, lam, q, y) : 0.0;
	}
#define MAT(t) (k->Mconst ? k->Mconst : (Mn ? Mn : Mt + (size_t)(t) * P * P))
	if (Mt)
		for (uint32_t t = 0; t < T; t++)
			for (uint32_t q = 0; q < P; q++) for (uint32_t y = 0; y < P; y++)
				Mt[((size_t)t * P + q) * P + y] = LEGAL(q, y) ? trans_value(c, m, X + (size_t)t * D * W, lam, q, y) : 0.0;   /* N states: legal pairs only */
	for (uint32_t t = 0; t < T; t++) {
		const uint32_t dmax = t + 1 < D ? t + 1 : D;
		for (uint32_t y = 0; y < P; y++) {
			for (uint32_t d = 1; d <= dmax; d++) {
				const double s = state_value(c, m, X + ((size_t)t * D + (d - 1)) * W, lam, y);
				S[((size_t)t * D + d - 1) * P + y] = s;
				acc[d - 1] = AD[((size_t)t * D + d - 1) * P + y] = s + (d <= t ? AT[(size_t)(t - d) * P + y] : 0.0);
			}
			A[(size_t)t * P + y] = log_add_n(acc, (int)dmax);
		}
		if (t + 1 < T) {
			const double* M = MAT(t + 1);
			for (uint32_t y = 0; y < P; y++) {
				int na = 0;
				for (uint32_t q = 0; q < P; q++) if (LEGAL(q, y)) acc[na++] = A[(size_t)t * P + q] + M[(size_t)q * P + y];
				AT[(size_t)t * P + y] = log_add_n(acc, na);
			}
		}
	}
	const double Zx = log_add_n(A + (size_t)(T - 1) * P, (int)P);
	double ll = 0.0;
	int bad = 0;
	for (uint32_t t = T; t-- > 0;) {
		const uint32_t nn = (T - 1 - t) < D ? (T - 1 - t) : D;
		if (nn == 0) for (uint32_t y = 0; y < P; y++) { B[(size_t)t * P + y] = 0.0; BV[(size_t)t * P + y] = LOG0; }
		else {
			const double* M = MAT(t + 1);
			for (uint32_t y = 0; y < P; y++) {
				for (uint32_t d = 1; d <= nn; d++) acc[d - 1] = S[((size_t)(t + d) * D + d - 1) * P + y] + B[(size_t)(t + d) * P + y];
				BV[(size_t)t * P + y] = log_add_n(acc, (int)nn);
			}
			for (uint32_t q = 0; q < P; q++) {
				int na = 0;
				for (uint32_t y = 0; y < P; y++) if (LEGAL(q, y)) acc[na++] = M[(size_t)q * P + y] + BV[(size_t)t * P + y];
				B[(size_t)t * P + q] = log_add_n(acc, na);
			}
		}
	}
	for (uint32_t t = 0; t < T && !bad; t++) {
		const uint32_t* l4 = lab4 + 4 * (size_t)t;
		const uint32_t dmax = t + 1 < D ? t + 1 : D;
		const uint32_t lab = l4[0], ldur = (lab != CRFO_LAB_BAD) ? l4[2] - l4[1] + 1 : 0;
		double tot = 0.0;
		for (uint32_t d = 1; d <= dmax; d++) {
			const float* xs = X + ((size_t)t * D + (d - 1)) * W;
			for (uint32_t y = 0; y < P; y++) {
				const double ab = exp(AD[((size_t)t * D + d - 1) * P + y] + B[(size_t)t * P + y] - Zx);
				tot += ab;
				ll += state_expf(c, m, xs, lam, k->ExpF, grad, ab, (lab != CRFO_LAB_BAD && d == ldur) ? lab : CRFO_LAB_BAD, y);
			}
		}
		if (tot > 1.000001 || tot < -0.000001) bad = 1;
		/* transitions out of frame t into the segment that starts at t+1, with the features of frame t+1's duration-1 window; the
		 * reference pair is (segment ending here, next reference segment) -- computeExpF :737-792 and the builder's next_lab (:436-453) */
		if (t + 1 < T) {
			const double* M = MAT(t + 1);
			const float* xn = X + (size_t)(t + 1) * D * W;
			uint32_t next_lab = CRFO_LAB_BAD;
			for (uint32_t u = t + 1; u < T; u++) { next_lab = lab4[4 * (size_t)u]; if (next_lab != CRFO_LAB_BAD) break; }
			double ttot = 0.0;
			for (uint32_t q = 0; q < P; q++)
				for (uint32_t y = 0; y < P; y++) {
					if (!LEGAL(q, y)) continue;
					const double xi = exp(A[(size_t)t * P + q] + M[(size_t)q * P + y] + BV[(size_t)t * P + y] - Zx);
					ttot += xi;
					const int match = lab != CRFO_LAB_BAD && next_lab != CRFO_LAB_BAD && q == lab && y == next_lab;
					ll += trans_expf(c, m, xn, lam, k->ExpF, grad, xi, match ? lab : CRFO_LAB_BAD, match ? next_lab : CRFO_LAB_BAD, q, y);
				}
			if (ttot > 1.000001 || ttot < -0.000001) bad = 1;
		}
	}
#undef MAT
#undef LEGAL
	if (!bad) {
		for (uint32_t i = 0; i < m->len; i++) grad[i] -= k->ExpF[i];
		*numer = ll; *logZ = Zx;
		if (A_out) memcpy(A_out, A, sizeof(double) * (size_t)T * P);
		if (B_out) memcpy(B_out, B, sizeof(double) * (size_t)T * P);
	}
	free(X); free(S); free(AD); free(A); free(AT); free(B); free(BV); free(Mt); free(Mn); free(acc);
	if (bad) FAIL("posterior mass check failed (no_dur)");
	return 0;
}

static int ctx_init(ctx_t* k, const crforacle_config* c, const fmap_t* m, const double* lam) {
	k->x2 = NULL;
	k->c = c; k->m = m; k->lam = lam;
	k->ExpF = (double*)malloc(sizeof(double) * (m->len ? m->len : 1));
	k->Mconst = NULL;
	if (!c->use_trans_ftrs && (c->n_states == 1)) {
		uint32_t L = m->L;
		k->Mconst = (double*)malloc(sizeof(double) * (size_t)L * L);
		for (uint32_t p = 0; p < L; p++) for (uint32_t cl = 0; cl < L; cl++)
			k->Mconst[(size_t)p * L + cl] = trans_value(c, m, NULL, lam, p, cl);
	}
	return 0;
}
static void ctx_free(ctx_t* k) { free(k->ExpF); free(k->Mconst); }

static int fb_one(ctx_t* k, uint32_t T, const float* x, const uint32_t* labs,
                  double* grad, double* numer, double* logZ, double* A, double* B) {
	const crforacle_config* c = k->c;
	if (T == 0) FAIL("No features read from this sentence");
	uint32_t* lab4 = (uint32_t*)malloc(sizeof(uint32_t) * 4 * (size_t)T);
	crforacle_window_labs(c, T, labs, lab4);
	int rc;
	if (c->model_type == CRFO_STDFRAME) {
		if (c->max_dur != 1) { free(lab4); FAIL("stdframe requires max_dur==1"); }
		float* Xw = NULL;
		if (has_ext(c)) {   /* context frames / joined stream: the frame-level recursions read the expanded window vectors */
			Xw = (float*)malloc(sizeof(float) * (size_t)T * crforacle_window_width(c));
			if (crforacle_window_ftrs2(c, T, x, k->x2, Xw)) { free(Xw); free(lab4); return 1; }
			x = Xw;
		}
		rc = (c->n_states > 1) ? fb_frame_nstate(k, T, x, lab4, grad, numer, logZ, A, B)
		                       : fb_frame_1state(k, T, x, lab4, grad, numer, logZ, A, B);
		free(Xw);
	} else if (c->model_type == CRFO_STDSEG) rc = fb_stdseg(k, T, x, lab4, grad, numer, logZ, A, B);
	else if (c->model_type == CRFO_STDSEG_NO_DUR || c->model_type == CRFO_STDSEG_NO_DUR_NO_TRANSFTR || c->model_type == CRFO_STDSEG_NO_DUR_NO_SEGTRANSFTR)
		rc = fb_nodur(k, T, x, lab4, grad, numer, logZ, A, B);
	else { free(lab4); FAIL("oracle: model type %u forward-backward not restated yet", c->model_type); }
	free(lab4);
	return rc;
}

typedef struct {
	const crforacle_config* c; const fmap_t* m; const double* lam;
	uint32_t n_utt; const uint32_t* off; const float* ftrs; const float* ftrs2; uint32_t u0; const uint32_t* labs;
	double* grad; double* numer; double* logZ; int rc; char err[512];
} shard_t;

/* first row of utterance u (absolute index) in a stream that carries lc + rc context frames per utterance */
static size_t stream_row(const uint32_t* off_abs, uint32_t u_abs, uint32_t lc, uint32_t rc) { return (size_t)off_abs[0] + (size_t)u_abs * (lc + rc); }

static void* shard_run(void* p) {
	shard_t* s = (shard_t*)p;
	ctx_t k; ctx_init(&k, s->c, s->m, s->lam);
	s->rc = 0;
	for (uint32_t u = 0; u < s->n_utt && !s->rc; u++) {
		uint32_t T = s->off[u + 1] - s->off[u];
		const crforacle_config* c = s->c;
		k.x2 = s->ftrs2 ? s->ftrs2 + stream_row(s->off + u, s->u0 + u, c->left_ctx2, c->right_ctx2) * c->n_base_ftrs2 : NULL;
		s->rc = fb_one(&k, T, s->ftrs + stream_row(s->off + u, s->u0 + u, c->left_ctx, c->right_ctx) * c->n_base_ftrs, s->labs + s->off[u],
		               s->grad, &s->numer[u], &s->logZ[u], NULL, NULL);
	}
	if (s->rc) memcpy(s->err, g_err, sizeof s->err);
	ctx_free(&k);
	return NULL;
}

int crforacle_fwdbwd_mt(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                        uint32_t n_utt, const uint32_t* off, const float* ftrs, const uint32_t* labs,
                        double* grad, double* numer, double* logZ, uint32_t n_threads) {
	return crforacle_fwdbwd_mt2(c, lambda, lambda_len, n_utt, off, ftrs, NULL, labs, grad, numer, logZ, n_threads);
}

int crforacle_fwdbwd_mt2(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                         uint32_t n_utt, const uint32_t* off, const float* ftrs, const float* ftrs2, const uint32_t* labs,
                         double* grad, double* numer, double* logZ, uint32_t n_threads) {
	if (c->n_base_ftrs2 && !ftrs2) FAIL("the configuration joins a second feature stream but none was passed");
	fmap_t m; if (fmap_build(c, &m)) return 1;
	if (m.len != lambda_len) { uint32_t l = m.len; fmap_free(&m); FAIL("lambda length mismatch: map has %u, caller passed %u", l, lambda_len); }
	if (n_threads < 1) n_threads = 1;
	if (n_threads > n_utt) n_threads = n_utt ? n_utt : 1;
	shard_t* sh = (shard_t*)calloc(n_threads, sizeof(shard_t));
	pthread_t* th = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
	uint32_t per = n_utt / n_threads;
	for (uint32_t i = 0; i < n_threads; i++) {
		uint32_t start = i * per, cnt = (i == n_threads - 1) ? n_utt - start : per;
		sh[i].c = c; sh[i].m = &m; sh[i].lam = lambda; sh[i].n_utt = cnt; sh[i].off = off + start;
		sh[i].ftrs = ftrs; sh[i].ftrs2 = c->n_base_ftrs2 ? ftrs2 : NULL; sh[i].u0 = start; sh[i].labs = labs; sh[i].numer = numer + start; sh[i].logZ = logZ + start;
		sh[i].grad = (n_threads == 1) ? grad : (double*)calloc(lambda_len ? lambda_len : 1, sizeof(double));
	}
	if (n_threads == 1) shard_run(&sh[0]);
	else {
		for (uint32_t i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, shard_run, &sh[i]);
		for (uint32_t i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
		for (uint32_t i = 0; i < n_threads; i++) {
			for (uint32_t j = 0; j < lambda_len; j++) grad[j] += sh[i].grad[j];
			free(sh[i].grad);
		}
	}
	int rc = 0;
	for (uint32_t i = 0; i < n_threads; i++) if (sh[i].rc) { rc = sh[i].rc; memcpy(g_err, sh[i].err, sizeof g_err); }
	free(sh); free(th); fmap_free(&m);
	return rc;
}

int crforacle_fwdbwd_dump(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                          uint32_t T, const float* ftrs, const uint32_t* labs,
                          double* grad, double* numer, double* logZ, double* alpha, double* beta) {
	fmap_t m; if (fmap_build(c, &m)) return 1;
	if (m.len != lambda_len) { fmap_free(&m); FAIL("lambda length mismatch"); }
	if (c->n_base_ftrs2) { fmap_free(&m); FAIL("crforacle_fwdbwd_dump takes one feature stream"); }
	ctx_t k; ctx_init(&k, c, &m, lambda);
	int rc = fb_one(&k, T, ftrs, labs, grad, numer, logZ, alpha, beta);
	ctx_free(&k); fmap_free(&m);
	return rc;
}

/* ------------------------------------------------------------------------------------------
 * Viterbi: array restatement of the token-passing decoder
 * CRF_ViterbiDecoder_StdSeg_NoSegTransFtr<CRF_ViterbiNode>::nStateDecode with lm_fst==NULL (free
 * phone loop) and beam 0 (CRF/src/decoders/CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:1369-2398).
 *
 * Facts restated from the reference:
 *  - costs are float; 99999.0 is a finite "infinity" that is added to like any number; a slot is
 *    live iff cost < 99999.0 (.cpp:124-128)
 *  - candidates for segments STARTING at frame s: cross-phone (prev end sub-state -> start sub-state),
 *    scanned over the kept list of frame s-1 in list order, then within-phone (self vs advance,
 *    self only if strictly smaller, .cpp:338); the merge keeps the first arrival on ties (.h:301,346)
 *  - each candidate is deposited for dur=1..D into nodes s..s+D-1 (.cpp:394-404,505-517); node t adds
 *    -S_t[label,dur] to live slots (.cpp:116-166) and keeps, per (phone, sub-state), the best
 *    duration, slots scanned in insertion order => longest duration wins ties (.h:383-465)
 *  - kept-list order = first-appearance order of phones among node t's slots (.cpp:1054-1097)
 *  - final: min over kept list of the end sub-state, first wins (.cpp:2156-2171); traceback :2204-2349
 * ---------------------------------------------------------------------------------------- */
#define VINF 99999.0f

typedef struct { float w; int ptr; } cand_t;  /* ptr = previous (phone*N+sub) or -1 */

/* lm_start / lm_bigram / lm_final (all NULL: the decoder's own free-phone LM): a phone-bigram LM in the free-phone LM's topology (one
 * state per phone; createFreePhoneLmFst .cpp:1270-1348) with a weight on every arc and a final weight per phone state:
 *  - the arc weight is added to the expanding hypothesis BEFORE the transition score, float by float: (prev + lm) + trans
 *    (expandCrossStateFromPrevNode :629, crossStateTransUpdate :467);
 *  - N states per phone: lm_start = unigram costs, lm_bigram = P exit costs on the epsilon arcs back to the LM's start state (the topology
 *    of the N-state free-phone LM, .cpp:1313-1330); the epsilon closure keeps one entry per (start state, previous phone), expanded in
 *    the order the hypotheses were visited (:640-733), i.e. the kept-list order the free-phone path already follows;
 *  - with an input LM the final hypothesis is the minimum over the finalStateSet -- ordered by LM state id, i.e. by phone --
 *    of weight + final weight (expandFinalNode :746-758, addToFinalSet :929-946, selection :2138-2153); that sum is the path cost. */
typedef struct { const float* start; const float* bigram; const float* fin; double beam; } phone_lm_t;
/* beam > 0 (one state per phone): pruning() (.cpp:976-1106) keeps the hypotheses of a node whose weight is < min_weight + beam (float
 * against a double sum); only kept hypotheses are expanded (cross-phone, :573, and within-phone) and may end the path.  In the free-phone
 * / bigram topology every kept hypothesis still reaches every other phone, so every phone has a candidate at every start frame and the
 * kept list is the closed-form order with the pruned phones left out. */

static int viterbi_one(const crforacle_config* c, const fmap_t* m, const double* lam,
                       uint32_t T, const float* x, const float* x2, const phone_lm_t* lm, uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn,
                       uint32_t* n_seg, float* path_cost) {
	uint32_t L = m->L, N = m->nStates, P = m->nAct, D = c->max_dur, W = crforacle_window_width(c);
	if (T == 0) FAIL("empty utterance");
	float* X = (float*)malloc(sizeof(float) * (size_t)T * D * W);
	crforacle_window_ftrs2(c, T, x, x2, X);
	/* candidates for segments starting at frame s: C[s][lab] */
	cand_t* C = (cand_t*)malloc(sizeof(cand_t) * (size_t)T * L);
	float* Wt = (float*)malloc(sizeof(float) * (size_t)T * L);        /* kept weights per frame */
	int* bp = (int*)malloc(sizeof(int) * (size_t)T * L);              /* pointer of the kept slot */
	uint32_t* bd = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)T * L);
	uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)T * P);  /* kept-list phone order per frame */
	uint32_t* arrive = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)T * P); /* arrival order of slots created at frame s */
	uint8_t* kept = (uint8_t*)malloc((size_t)T * P);                        /* beam pruning: phone survives node t */
	const double beam = lm ? lm->beam : 0.0;
	if (lm && !lm->start) lm = NULL;                                        /* beam without an input LM */
	if (beam > 0.0 && N != 1) { free(X); free(C); free(Wt); free(bp); free(bd); free(order); free(arrive); free(kept); FAIL("oracle: beam pruning restated for one state per phone only"); }
	/* (float)(-M[p,c]) for the legal pairs, precomputed when transitions carry no features */
	float* negM = NULL;
	if (!c->use_trans_ftrs) {
		negM = (float*)malloc(sizeof(float) * (size_t)L * L);
		for (uint32_t p = 0; p < L; p++) for (uint32_t cl = 0; cl < L; cl++)
			if (m->tidx[(size_t)p * L + cl] != CRFO_NO_IDX) negM[(size_t)p * L + cl] = -1 * trans_value(c, m, NULL, lam, p, cl);
	}
#define NEGM(p, cl) (negM ? negM[(size_t)(p) * L + (cl)] : (float)(-1 * trans_value(c, m, xs, lam, (p), (cl))))
	int rc = 0;
	for (uint32_t s = 0; s < T && !rc; s++) {
		cand_t* Cs = C + (size_t)s * L;
		uint32_t* arr = arrive + (size_t)s * P;
		if (s == 0) {
			for (uint32_t q = 0; q < P; q++) {
				arr[q] = q;
				for (uint32_t k = 0; k < N; k++) { Cs[q * N + k].w = (k == 0) ? (0.0f + (lm ? lm->start[q] : 0.0f)) + 0.0f : VINF; Cs[q * N + k].ptr = -1; }
			}
		} else {
			/* transition scores of node s from its dur-1 window (…WithoutSegTransFtr.cpp:39-74) */
			const float* xs = X + ((size_t)s * D) * W;
			const float* Wp = Wt + (size_t)(s - 1) * L;
			const uint32_t* ord = order + (size_t)(s - 1) * P;
			const uint8_t* kp = kept + (size_t)(s - 1) * P;
			uint32_t narr = 0;
			uint8_t* seen = (uint8_t*)calloc(P, 1);
			for (uint32_t q = 0; q < P; q++) for (uint32_t k = 0; k < N; k++) { Cs[q * N + k].w = VINF; Cs[q * N + k].ptr = -1; }
			/* cross-phone candidates, in kept-list order x LM-arc order */
			for (uint32_t i = 0; i < P; i++) {
				uint32_t pp = ord[i], pend = pp * N + N - 1;
				if (!kp[pp]) continue;                   /* pruned at node s-1 */
				for (uint32_t q = 0; q < P; q++) {
					if (N == 1 && q == pp) continue;     /* free-phone LM, 1 state: no self arc (.cpp:1332-1346) */
					/* + LM arc weight (0 in the free-phone LM).  N states per phone: the hypothesis first takes the epsilon arc back to the LM's
					 * start state (exit cost of phone pp), then the unigram arc of q: two float adds in that order (:597, :686) */
					float base = !lm ? Wp[pend] + 0.0f : (N == 1 ? Wp[pend] + lm->bigram[(size_t)pp * P + q] : (Wp[pend] + lm->bigram[pp]) + lm->start[q]);
					float tw = NEGM(pend, q * N);
					float cost = base + tw;
					if (!seen[q]) { seen[q] = 1; arr[narr++] = q; Cs[q * N].w = cost; Cs[q * N].ptr = (int)pend; }
					else if (cost < Cs[q * N].w) { Cs[q * N].w = cost; Cs[q * N].ptr = (int)pend; }
				}
			}
			/* within-phone candidates, in kept-list order */
			for (uint32_t i = 0; i < P; i++) {
				uint32_t q = ord[i];
				if (!kp[q]) continue;
				int fresh = !seen[q];
				if (fresh) { seen[q] = 1; arr[narr++] = q; }
				for (uint32_t k = 0; k < N; k++) {
					uint32_t lab = q * N + k;
					float w; int ptr;
					float tw1 = NEGM(lab, lab);
					float n1 = Wp[lab] + tw1;
					if (k == 0) { w = n1; ptr = (int)lab; }
					else {
						float tw2 = NEGM(lab - 1, lab);
						float n2 = Wp[lab - 1] + tw2;
						if (n1 < n2) { w = n1; ptr = (int)lab; } else { w = n2; ptr = (int)lab - 1; }
					}
					if (fresh || w < Cs[lab].w) { Cs[lab].w = w; Cs[lab].ptr = ptr; }
				}
			}
			free(seen);
			if (narr != P) { snprintf(g_err, sizeof g_err, "internal: arrival count"); rc = 1; break; }
		}
		/* node t == s: add state values, choose best duration per (phone, sub-state) */
		uint32_t t = s, dmax = t + 1 < D ? t + 1 : D;
		for (uint32_t lab = 0; lab < L; lab++) {
			float best = 0; int bptr = -1; uint32_t bdur = 0; int have = 0;
			for (uint32_t d = dmax; d >= 1; d--) {   /* longest first == slot insertion order */
				const cand_t* cd = C + (size_t)(t - d + 1) * L + lab;
				float w = cd->w;
				if (w < VINF) {
					float sv = -1 * state_value(c, m, X + ((size_t)t * D + (d - 1)) * W, lam, lab);
					w = w + sv;
				}
				if (!have || w < best) { best = w; bptr = cd->ptr; bdur = d; have = 1; }
			}
			Wt[(size_t)t * L + lab] = best; bp[(size_t)t * L + lab] = bptr; bd[(size_t)t * L + lab] = bdur;
		}
		/* pruning (.cpp:976-1106): min over the node, keep weight < min + beam (the sum in double) */
		{
			float mn = VINF;
			for (uint32_t q = 0; q < P; q++) { float bw = Wt[(size_t)t * L + q * N]; for (uint32_t k = 1; k < N; k++) if (Wt[(size_t)t * L + q * N + k] < bw) bw = Wt[(size_t)t * L + q * N + k]; if (bw < mn) mn = bw; }
			for (uint32_t q = 0; q < P; q++) kept[(size_t)t * P + q] = (beam <= 0.0) ? 1 : ((double)Wt[(size_t)t * L + q] < (double)mn + beam);
		}
		/* kept-list order = arrival order of the slots inserted earliest into node t */
		uint32_t sfirst = t + 1 >= D ? t + 1 - D : 0;
		memcpy(order + (size_t)t * P, arrive + (size_t)sfirst * P, sizeof(uint32_t) * P);
	}
#undef NEGM
	if (!rc) {
		const uint32_t* ord = order + (size_t)(T - 1) * P;
		float minw = VINF; int best = -1;
		for (uint32_t i = 0; i < P; i++) {
			/* no LM: kept-list order; with an LM: LM-state order = phone order, final weights added, states that are not final skipped */
			uint32_t q = lm ? i : ord[i], e = q * N + N - 1;
			if (!kept[(size_t)(T - 1) * P + q]) continue;
			if (lm && isinf(lm->fin[q])) continue;
			float w = Wt[(size_t)(T - 1) * L + e];
			if (lm) w = w + lm->fin[q];
			if (w < minw) { minw = w; best = (int)e; }
		}
		*path_cost = minw;
		if (best < 0) { *n_seg = 0; }
		else {
			/* trace back, then reverse */
			uint32_t ns = 0; int end = (int)T - 1; int lab = best;
			while (end >= 0) {
				uint32_t d = bd[(size_t)end * L + lab];
				out_lab[ns] = (uint32_t)lab; out_dur[ns] = d;
				int start = end + 1 - (int)d;
				int prev = bp[(size_t)end * L + lab];
				/* phone emitted when the arc enters the start sub-state across a phone boundary
				 * (isPhoneStartBoundary, .cpp:2327-2334); the first segment always emits (.cpp:2316) */
				if (start == 0) out_phn[ns] = (uint32_t)lab / N;
				else out_phn[ns] = ((uint32_t)lab % N == 0 && prev != lab) ? (uint32_t)lab / N : CRFO_LAB_BAD;
				ns++;
				if (start == 0) break;
				lab = prev; end = start - 1;
			}
			for (uint32_t i = 0; i < ns / 2; i++) {
				uint32_t a;
				a = out_lab[i]; out_lab[i] = out_lab[ns - 1 - i]; out_lab[ns - 1 - i] = a;
				a = out_dur[i]; out_dur[i] = out_dur[ns - 1 - i]; out_dur[ns - 1 - i] = a;
				a = out_phn[i]; out_phn[i] = out_phn[ns - 1 - i]; out_phn[ns - 1 - i] = a;
			}
			*n_seg = ns;
		}
	}
	free(X); free(C); free(Wt); free(bp); free(bd); free(order); free(arrive); free(negM); free(kept);
	return rc;
}

int crforacle_viterbi(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                      uint32_t n_utt, const uint32_t* off, const float* ftrs,
                      uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                      float* path_cost, double* logZ) {
	return crforacle_viterbi2(c, lambda, lambda_len, n_utt, off, ftrs, NULL, out_lab, out_dur, out_phn, n_seg, path_cost, logZ);
}

int crforacle_viterbi2(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                       uint32_t n_utt, const uint32_t* off, const float* ftrs, const float* ftrs2,
                       uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                       float* path_cost, double* logZ) {
	return crforacle_viterbi_lm(c, lambda, lambda_len, n_utt, off, ftrs, ftrs2, NULL, NULL, NULL, out_lab, out_dur, out_phn, n_seg, path_cost, logZ);
}

int crforacle_viterbi_lm(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                         uint32_t n_utt, const uint32_t* off, const float* ftrs, const float* ftrs2,
                         const float* lm_start, const float* lm_bigram, const float* lm_final,
                         uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                         float* path_cost, double* logZ) {
	return crforacle_viterbi_beam(c, lambda, lambda_len, n_utt, off, ftrs, ftrs2, lm_start, lm_bigram, lm_final, 0.0, out_lab, out_dur, out_phn, n_seg, path_cost, logZ);
}

int crforacle_viterbi_beam(const crforacle_config* c, const double* lambda, uint32_t lambda_len,
                           uint32_t n_utt, const uint32_t* off, const float* ftrs, const float* ftrs2,
                           const float* lm_start, const float* lm_bigram, const float* lm_final, double beam,
                           uint32_t* out_lab, uint32_t* out_dur, uint32_t* out_phn, uint32_t* n_seg,
                           float* path_cost, double* logZ) {
	phone_lm_t lmv = {lm_start, lm_bigram, lm_final, beam};
	const phone_lm_t* lm = (lm_start || beam > 0.0) ? &lmv : NULL;
	if (lm_start && (!lm_bigram || !lm_final)) FAIL("the phone LM needs all three weight arrays");

	if (c->n_base_ftrs2 && !ftrs2) FAIL("the configuration joins a second feature stream but none was passed");
	if (c->model_type != CRFO_STDSEG_NO_DUR_NO_SEGTRANSFTR && c->model_type != CRFO_STDFRAME)
		FAIL("viterbi: only stdframe / stdseg_no_dur_no_segtransftr are accepted (CRFDecode/src/Main.cpp:1065-1076)");
	fmap_t m; if (fmap_build(c, &m)) return 1;
	if (m.len != lambda_len) { fmap_free(&m); FAIL("lambda length mismatch"); }
	int rc = 0;
	for (uint32_t u = 0; u < n_utt && !rc; u++) {
		uint32_t T = off[u + 1] - off[u];
		rc = viterbi_one(c, &m, lambda, T, ftrs + stream_row(off + u, u, c->left_ctx, c->right_ctx) * c->n_base_ftrs,
		                 c->n_base_ftrs2 ? ftrs2 + stream_row(off + u, u, c->left_ctx2, c->right_ctx2) * c->n_base_ftrs2 : NULL, lm,
		                 out_lab + off[u], out_dur + off[u], out_phn + off[u], &n_seg[u], &path_cost[u]);
		if (logZ) logZ[u] = 0.0;
	}
	fmap_free(&m);
	return rc;
}
