// Device kernels of libcrfgpu (sm_100a).  Declarations + parameter blocks; definitions in crf_kernels.cu.
// Citations are relative to the ASR-CRaFT tree.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

namespace crfgpu {

constexpr uint32_t LAB_BAD = 0xffffffffu;
constexpr float VIT_INF = 99999.0f;  // the decoder's finite "infinity" (CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:124-128)

// ---- window expansion (K7) ---------------------------------------------------------------------
struct ExpandParams {
	const float* base;        // [N][F]
	const uint32_t* frame_t;  // [N] frame index inside its utterance
	const uint32_t* steps;    // [D*5] sample offsets (host-computed, see sample_steps())
	float* X;                 // [N][D][Wp]  (Wp >= W: windows start on 128-byte lines, the pad is zero)
	uint32_t N, F, D, W, Wp;
	uint32_t seg_ftrs;        // 1: [5 samples|avg|max|min|one-hot dur], 0: first frame of the window
	uint32_t n0;              // first frame of this launch
	uint32_t dpart;           // durations per CTA (0 = all): small groups keep shared memory low enough to run beside a lattice CTA
};
void launch_expand_windows(const ExpandParams& p, uint32_t n1, cudaStream_t s);   // frames [p.n0, n1)
// virtual windows of the training GEMMs: padded base stream base2[N][Fp] and the aggregate blocks Xa[N][D][Wa] (avg | max | min | zero pad)
// of frames [n0, n1); the five sampled-frame blocks are read from base2 at row shifts, the one-hot block rides in the bias
void launch_virtual_windows(const float* base, const uint32_t* frame_t, float* base2, float* Xa, uint32_t N, uint32_t F, uint32_t Fp, uint32_t D,
                            uint32_t Wa, uint32_t n0, uint32_t n1, cudaStream_t s);

// ---- window streams with context frames and / or a joined second stream (general path; crfgpu_*2 entry points) -----------------------
struct JoinedPart {
	const float* x;           // the stream's frames, lc + T + rc rows per utterance
	uint32_t F, seg, lc, rc, bdelta, width;
};
struct ExpandJoinedParams {
	JoinedPart part[2]; uint32_t n_parts;
	const uint32_t* frame_t; const uint32_t* frame_utt; const uint32_t* steps;
	float* X;                 // [N][D][Wp]
	uint32_t N, D, Wp;
	uint32_t keep_lo, keep_hi; // keep_hi != 0: only the columns [keep_lo, keep_hi) are materialised in windows longer than one frame (zeros elsewhere)
};
void launch_expand_joined(const ExpandJoinedParams& p, cudaStream_t s);

// ---- GEMM-1: state scores  S[n][col0+j] = sum_k A[n][k]*B[j][k] + bias[j] ----------------------
struct ScoreGemmParams {
	const float* A; uint64_t lda;     // [M][K] row stride lda (window features of one duration)
	const float* B; uint32_t ldb;     // [Ncols][K] state weights
	const float* bias;                // [Ncols] (already multiplied by stateBiasVal) or nullptr
	float* C; uint32_t ldc;           // [M][ldc], written at column offset applied by the caller
	uint32_t M, Ncols, K;
	const uint32_t* frame_t; uint32_t min_t;  // rows with frame_t[n] < min_t are skipped (window d needs t >= d-1)
	uint32_t n_fast;                  // set by launch_score_gemm_tc: column tiles run fastest in the grid (see there)
};
void launch_score_gemm(const ScoreGemmParams& p, cudaStream_t s);          // fp32 FFMA tiles (round-1 baseline)
cudaError_t launch_score_gemm_tc(const ScoreGemmParams& p, cudaStream_t s); // tcgen05, split-bf16 operands (crf_tc_gemm.cu)

// ---- dense lattice recursions in the probability domain ----------------------------------------
// One CTA owns U utterances ("slots"); thread c owns label column c.  L <= 1024.
struct DpParams {
	uint32_t L, Lp, P, D;       // labels, padded row stride, phones (= L / D), max duration
	uint32_t n_groups;
	const uint32_t* grp_utt;    // [n_groups*U] utterance id per slot, LAB_BAD = empty slot
	const uint32_t* off;        // [n_utt+1]
	const float* S;             // [N][Lp] state scores
	const float* E;             // [L][Lp]  E[q*Lp+c]  = exp(M[q][c]-Mmax), 0 for illegal pairs
	const float* ET;            // [L][Lp]  ET[c*Lp+q] = E[q][c]
	float* A;                   // [N][Lp] alpha, max-normalised per frame (probability domain)
	float* G;                   // [N][Lp] forward: pushed sums a_t*E ; backward: w~_t = exp(S+beta-kappa)
	double* m;                  // [N] log scale of A
	double* kappa;              // [N] log scale of w~
	double* bbase;              // [N] beta_t[c] = bbase[n] + log(u[c]); u stored in U_out when requested
	float* Uvec;                // [N][Lp] (optional, for alpha/beta dumps) unnormalised backward sums
	double* logZ;               // [n_utt]
	float* Dm;                  // [N][Lp] onehot(reference label) - gamma
	float* R;                   // [N][Lp] exp(S+beta+m_{t-d}+Mmax-logZ), the right factor of xi
	const uint32_t* node_lab;   // [N] label (incl. duration) of the reference segment ending here, or LAB_BAD
	double Mmax;
	double* mass;               // [N] sum_c gamma (posterior-mass check), optional
};
void launch_forward(const DpParams& p, int U, cudaStream_t s);
void launch_backward(const DpParams& p, int U, cudaStream_t s);
// frame-level models (D == 1) with at most 64 labels: one warp per utterance, E in registers (crf_dp_frame.cu); utt_list = the
// utterance ids in the order the warps take them
cudaError_t launch_frame_dp(bool backward, const DpParams& p, const uint32_t* utt_list, uint32_t n_utt, cudaStream_t s);
// both chains in one launch (the beta chain stores u_t in p.Uvec and bbase_t), then the posteriors Dm / R of all frames
cudaError_t launch_frame_dp_pair(const DpParams& p, const uint32_t* utt_list, uint32_t n_utt, cudaStream_t s);
cudaError_t launch_frame_post(const DpParams& p, const uint32_t* frame_t, const uint32_t* frame_utt, uint32_t N, cudaStream_t s);

// ---- cluster-resident variant (crf_dp_cluster.cu): E sliced over the CTAs of a thread-block cluster ----
struct ClusterDpParams : DpParams {
	uint32_t CS, CW, CWp, CWt;    // cluster size, label columns per CTA, smem row stride, column-threads per slot quad
	uint32_t n_clusters;
	const uint32_t* cl_off;       // [n_clusters+1] offsets into cl_list
	const uint32_t* cl_list;      // utterance ids in the order each cluster processes them
	float* xch;                   // [n_clusters][2][UB][Lp] log-domain frame vectors exchanged through L2
	float* xmax;                  // [n_clusters][2][CS][UB] per-CTA maxima of those vectors
};
struct ClusterPlan {
	uint32_t CS, CW, CWp, CWt, threads;
	int UB;                       // utterance slots advanced in lock-step by one cluster (multiple of 4)
	size_t smem;
};
bool plan_cluster_dp(uint32_t L, uint32_t D, int max_smem_optin, ClusterPlan* plan, int ub_cap = 32);
int max_active_clusters(const ClusterPlan& plan);
cudaError_t launch_cluster_dp(bool backward, const ClusterDpParams& p, const ClusterPlan& plan, cudaStream_t s);

constexpr int TC_DP_SLOTS = 16;     // utterances advanced in lock-step by one cluster == MMA N
// ---- tensor-core cluster variant (crf_dp_tc.cu): E hi half in TMEM, lo half in smem, tcgen05.mma per frame,
//      all-gather of the frame vector by bulk DSMEM copies ----
struct TcDpParams : DpParams {
	uint32_t CS, CW, K;           // cluster size, label slice per CTA (multiple of 16), padded label count CS*CW
	uint32_t tmem_cols, ctl_off;  // TMEM columns to allocate, byte offset of the control block in dynamic smem
	uint32_t n_clusters;
	const uint32_t* cl_off;       // [n_clusters*TC_DP_SLOTS+1] offsets of the per-slot utterance lists
	const uint32_t* cl_list;      // utterance ids in the order each slot processes them
	const float* smaxd;           // [N][D] per-duration maxima of S (launch_block_max)
	unsigned long long* dbg;      // optional [32] cycle counters of cluster 0 / CTA 0 (CRFGPU_DP_TIMING=1), else nullptr
};
struct TcDpPlan {
	uint32_t CS, CW, K, tmem_cols, ctl_off;
	size_t smem;
};
bool plan_tc_dp(uint32_t L, uint32_t D, int max_smem_optin, TcDpPlan* plan);
int max_active_tc_clusters(const TcDpPlan& plan);
cudaError_t launch_tc_dp(bool backward, const TcDpParams& p, const TcDpPlan& plan, cudaStream_t s);
void launch_block_max(const float* S, const uint32_t* frame_t, float* smaxd, uint32_t N, uint32_t Lp, uint32_t P, uint32_t D, cudaStream_t s);

// ---- contraction-sliced tensor-core variant (crf_dp_ks.cu): CTA r keeps E[:, slice r] of ALL label rows (hi and lo halves in TMEM) and
//      multiplies it with its own slice of the frame vector; the partial products are reduce-scattered by bulk DSMEM copies ----
struct KsDpParams : DpParams {
	uint32_t CS, CW, K, MT;       // cluster size, contraction slice per CTA (multiple of 16, <= 128), padded label count CS*CW, 128-row tiles
	uint32_t tmem_cols, recv_off, vbuf_off, ps_off, ctl_off;   // TMEM columns to allocate, byte offsets in dynamic shared memory
	uint32_t n_clusters;
	const uint32_t* cl_off;       // [n_clusters*TC_DP_SLOTS+1] offsets of the per-slot utterance lists
	const uint32_t* cl_list;      // utterance ids in the order each slot processes them
	const float* smaxd;           // [N][D] per-duration maxima of S
	unsigned long long* dbg;      // optional [32] cycle counters of cluster 0 / CTA 0 (CRFGPU_DP_TIMING=1), else nullptr
};
struct KsDpPlan {
	uint32_t CS, CW, K, MT, tmem_cols, recv_off, vbuf_off, ps_off, ctl_off;
	size_t smem;
};
bool plan_ks_dp(uint32_t L, uint32_t D, int max_smem_optin, KsDpPlan* plan);
int max_active_ks_clusters(const KsDpPlan& plan);
cudaError_t launch_ks_dp(bool backward, const KsDpParams& p, const KsDpPlan& plan, cudaStream_t s);

// ---- TN GEMM with fp64 scatter epilogue:  out[map(i,j)] += scale * sum_n A[n][i]*B[n][j] --------
struct ReduceGemmParams {
	const float* A; uint64_t lda;   // rows n in [n0,n1), A row used = n - a_row_shift
	const float* B; uint64_t ldb;
	uint32_t n0, n1, a_row_shift;
	uint32_t I, J;                  // output tile extents
	uint32_t ones_col;              // if < J: column j==ones_col of B is the constant 1 (bias feature)
	double scale, ones_scale;       // ones_scale applies to the ones column instead of scale
	// epilogue mode 0: out[row_idx[i] + j]           (state weights; row_idx = sidx of the label)
	// epilogue mode 1: out[pair_idx[i*pair_ld + j]] += scale*Ew[i*e_ld+j]*sum  (transition bias; skip NO_IDX)
	int mode;
	const uint32_t* row_idx;
	const uint32_t* pair_idx; uint32_t pair_ld;
	const float* Ew; uint32_t e_ld;
	double* out;
	uint32_t k_slab;                // rows of n per CTA (split-K)
};
void launch_reduce_gemm(const ReduceGemmParams& p, cudaStream_t s);
// tcgen05 version; m_side_is_b picks which operand's columns ride on the 128-row MMA M side (the wider one should)
cudaError_t launch_reduce_gemm_tc(const ReduceGemmParams& p, bool m_side_is_b, cudaStream_t s);

// the same product from pre-split, pre-tiled operands (launch_tile_mn), fed by bulk copies: out[row_idx[i] + j] += scale * sum_n A[n][i] * B[n][j]
struct TiledReduceParams {
	const unsigned char* At; const unsigned char* Bt;   // 128-column tiles of A (M side) and of B (N side)
	uint32_t N, I, J;                 // frames, columns of A (output rows), columns of B (output columns, the ones column included)
	uint32_t ones_col; double scale, ones_scale;
	const uint32_t* row_idx; double* out;
	uint32_t n_mt, n_nt, n_chunks, slab_chunks;         // set by the launcher
};
size_t tiled_operand_bytes(uint32_t N, uint32_t ncols, uint32_t T);      // T = 128 (M side) or 64 (N side)
cudaError_t launch_tile_mn(const float* src, uint64_t ld, uint32_t ncols, uint32_t ones_col, uint32_t N, bool m_side, unsigned char* dst, cudaStream_t s);
cudaError_t launch_reduce_gemm_tiled(const TiledReduceParams& p, cudaStream_t s);

// K-major twin: C[n][c] = sum_k A[n][k] * B[c][k] + bias[c] from operands tiled by launch_tile_k (rows x 32-feature chunks)
struct TiledScoreParams {
	const unsigned char* At; const unsigned char* Bt;   // 128-row tiles of A (frames) and of B (weight rows)
	const float* bias; float* C; uint32_t ldc;
	uint32_t M, Ncols, K;
	uint32_t n_kc;                    // set by the launcher
};
size_t tiled_k_operand_bytes(uint32_t rows, uint32_t K, uint32_t T);     // T = 128 (A) or 64 (B)
cudaError_t launch_tile_k(const float* src, uint64_t ld, uint32_t rows, uint32_t K, bool m_side, unsigned char* dst, cudaStream_t s);
cudaError_t launch_score_gemm_tiled(const TiledScoreParams& p, cudaStream_t s);

// ---- transition-bias expected counts for ALL durations in one pass (crf_tc_gemm.cu) --------------
// out[pair_idx[q*L + c]] += scale * Ew[q][c] * sum_n A[n-d(c)][q] * R[n][c],  c = (d-1)*P + y
struct XiGemmParams {
	const float* A; uint64_t lda;     // [n_frames][L] forward vectors
	const float* R; uint64_t ldb;     // [n_frames][L] right factors
	uint32_t n_frames, n0, n1, k_slab;
	uint32_t L, P, D;
	const uint32_t* pair_idx;         // [L][L] lambda index of (q,c) or 0xffffffff
	const float* Ew; uint32_t e_ld;   // E[q][c]
	double scale;
	double* out;
};
cudaError_t launch_xi_gemm_tc(const XiGemmParams& p, cudaStream_t s);

// ---- TMA-fed GEMMs over the window stream, all durations in one launch (crf_tma_gemm.cu) ---------
struct ScoreTmaParams {
	const unsigned char* Bt;          // state weights as bf16 hi/lo UMMA tiles, [(d*ntile + jt)*n_chunks + c][8 KB] (lambda_tables_kernel)
	const float* bias;                // [D*P] or nullptr
	float* C; uint32_t ldc;           // S[M][ldc], column (d*P + y)
	uint32_t M, P, K, D, n_chunks, ntile;
	uint32_t a_from_tmem;             // 1: window tile through tensor memory (score_gemm_tmem_kernel)
	uint32_t shared_w;                // 1: every duration block uses the same weight tiles / bias (labels = phones, stdseg_no_dur*)
	float* smaxd;                     // [M][D] per-duration row maxima (-inf where d > t) or nullptr; needs ntile == 1
	const uint32_t* frame_t;
	// virtual windows (virt = 1, needs a_from_tmem): the five sampled-frame blocks are row shifts of the padded base stream base2[M][cpb*32],
	// X holds only avg | max | min; chunks = 5 * cpb + ceil(width of X / 32) in that order (the weight tiles follow it), bias is [D*P]
	uint32_t virt, cpb;
	const float* base2;
	const uint32_t* steps;            // [D*5] sample offsets (sample_steps())
};
// ---- native stdseg_no_dur* lattice recursions, O(P^2 + D*P) per frame (crf_dp_nodur.cu) ---------------------------
constexpr int NODUR_UT = 16;          // utterances that advance in lock-step per group of CTAs
struct NodurParams {
	uint32_t P, Pp, D;            // phones, row stride of the P-wide arrays, durations
	uint64_t Lp;                  // row stride of S / Dm ([N][Lp], column (d-1)*P + y)
	uint32_t n_groups, npt;       // lock-step groups, CTAs (32-phone tiles) per group
	const uint32_t* grp_off;      // [n_groups+1] batches of each group
	const uint32_t* batch_utt;    // [n_batches][NODUR_UT] utterance ids or LAB_BAD
	const uint32_t* off;          // [n_utt+1]
	const float* S; const float* smaxd;   // [N][Lp] state scores, [N][D] their per-duration maxima (-inf where d > t+1)
	const float* E; const float* ET;      // [P][Pp] exp(M - Mmax) and its transpose
	double Mmax;
	float* A; float* LG; double* rho;     // forward: a_t [N][Pp], log(a_t E) [N][Pp], scale rho_t [N]
	double* logZ;                         // [n_utt]
	float* LB; float* R; float* Dm;       // backward: log(E bh_t) [N][Pp], Xi right factors [N][Pp], [ref] - gamma [N][Lp] (Dm: nodur_post_kernel)
	double* kappa;                        // [N] backward scale kappa_t of every frame (read by the posterior pass)
	const uint32_t* node_lab;             // [N] (dur-1)*P + phone where a reference segment ends, else LAB_BAD
	float* xch;                           // [n_groups][2][(Pk + npt) * NODUR_UT] exchange buffers, Pk = P rounded up to 32
	uint32_t* ctr;                        // [n_groups] barrier counters
	unsigned long long* dbg;              // optional [16] cycle counters of CTA 0 (CRFGPU_DP_TIMING=1)
};
size_t nodur_smem_bytes(uint32_t P);
int nodur_max_groups(uint32_t P);         // co-resident groups of ceil(P/32) CTAs (0: the phone count does not fit)
cudaError_t launch_nodur_dp(bool backward, const NodurParams& p, cudaStream_t s);
void launch_nodur_post(const NodurParams& p, const uint32_t* frame_t, const uint32_t* frame_utt, uint32_t N, cudaStream_t s);   // Dm = [ref] - gamma, behind the backward recursion

// Frame-reduction GEMMs (both operands [frames][columns], TMA-fed; one launch covers every duration block d):
//   state gradient  out[row_idx[d*P+y] + j]      += (j == ones_col ? ones_scale : scale) * sum_n X[n][d][j]  * Dm[n][d*P+y]
//   Xi              out[pair_idx[q*L + d*P+y]]    += scale * Ew[q][d*P+y]                 * sum_n A[n-d-1][q] * R[n][d*P+y]
constexpr uint32_t FRAME_GEMM_TILE = 61;   // columns of a duration block per CTA (64-wide box minus up to 3 columns of alignment slack)
struct FrameGemmParams {
	uint32_t N, P, D, ntile, k_slab;  // frames, phones per duration block, durations, FRAME_GEMM_TILE-column tiles per block, frames per CTA
	uint32_t Mext;                    // extent of the 128-row side: state features (+ bias) | lattice labels
	uint32_t ones_col;                // state gradient: index of the constant-1 bias feature or 0xffffffff
	double scale, ones_scale;
	const uint32_t* row_idx;          // state gradient: [D*P] lambda offset of the label's state block
	const uint32_t* pair_idx; uint32_t L; const float* Ew; uint32_t e_ld;   // Xi
	double* out;
	uint32_t a_from_tmem;             // 1: 128-row operand through tensor memory (frame_gemm_tmem_kernel)
	// state gradient over virtual windows (virt = 1, needs a_from_tmem): 128-row tiles = tpb sampled tiles (four 32-feature units each; unit
	// u = features [32 (u % upb), +32) of sampled-frame block u / upb, upb = Fp / 32: rows of base2[N][Fp] at the block's row shift) + the
	// tiles of the aggregate array; lambda row of feature f of block b = b*F + f, of aggregate column a = 5F + a; the
	// constant-1 row sits at aggregate column 3F and counts into the bias (ones_col) and into the one-hot duration weight 8F + d
	uint32_t virt, tpb, F, Fp;
	const float* base2;
	const uint32_t* steps;
};
// X must be 16-byte aligned, Wp % 4 == 0, the first state feature a multiple of 4 and the driver must export cuTensorMapEncodeTiled
bool tma_gemm_eligible(const float* X, uint32_t D, uint32_t Wp, uint32_t sf0);
uint32_t score_tma_chunks(uint32_t K);
bool lattice_tma_eligible(const float* a, uint32_t ld);
// Xi: A, R = lattice arrays [N][ld]
cudaError_t launch_xi_gemm_tma(const float* A, const float* R, uint32_t ld, const FrameGemmParams& p, cudaStream_t s);
// X points at the first state feature of window (frame 0, duration 1)
cudaError_t launch_score_gemm_tma(const float* X, uint32_t Wp, const ScoreTmaParams& p, cudaStream_t s);
cudaError_t launch_state_grad_tma(const float* X, uint32_t Wp, uint32_t K, const float* Dm, uint32_t ldd, const FrameGemmParams& p, cudaStream_t s);

// ---- empirical counts and numerators on the reference path -------------------------------------
struct EmpiricalParams {
	const float* X; uint64_t ldx;     // window features [N][D][W] (ldx = D*W)
	uint32_t W, sf0, nSf;
	const uint32_t* node_lab;         // [N]
	const uint32_t* prev_lab;         // [N] label of the previous reference segment or LAB_BAD
	const uint32_t* frame_utt;        // [N] utterance of the frame
	uint32_t N, L, P;
	uint32_t tL;                      // side of the transition index table: L, or P when labels fold to phones (native stdseg_no_dur*)
	const double* lambda;
	const uint32_t* sidx; const uint32_t* tidx;
	int use_state_bias, use_trans_bias;
	double state_bias_val, trans_bias_val;
	double* grad;                     // transition-bias empirical counts are added here
	double* numer;                    // [n_utt]
	// virtual windows: X = aggregate array [N][D][W] (avg | max | min), base = un-windowed stream [N][F], steps = sample offsets [D*5]
	uint32_t virt, F, D; const float* base; const uint32_t* steps;
};
void launch_empirical(const EmpiricalParams& p, cudaStream_t s);

// ---- Viterbi (token-passing decoder restated as arrays; free-phone LM, beam 0) ------------------
struct VitScoreParams {
	const float* X; uint64_t ldx;     // [N][D][W]
	uint32_t W, sf0, nSf, N, D, L;
	const uint32_t* frame_t;
	const double* Wd;                 // [nSf+1][L] state weights transposed, last row = bias weight
	int use_bias; double bias_val;
	float* negS;                      // [N][D][L]  (float)(-S) computed in fp64 in the reference's order
};
void launch_vit_scores(const VitScoreParams& p, cudaStream_t s);

struct VitParams {
	uint32_t n_utt, L, P, NS, D;
	const uint32_t* order;            // [n_utt] utterance handled by CTA i (longest first), or nullptr = identity
	uint32_t tbW;                     // set by launch_viterbi: frames per traceback window in shared memory
	uint32_t Ppad;                    // set by launch_viterbi: thread = (part of the kept list) * Ppad + target phone in the cross-phone scan
	const uint32_t* off;
	const float* negS;                // [N][D][L]
	const float* crossT;              // [P][P]  crossT[pp*P+q] = (float)(-M[end(pp)][start(q)])
	const float* negDiag;             // [L]     (float)(-M[l][l])
	const float* negOff;              // [L]     (float)(-M[l-1][l]) for sub-state > 0
	const float* negMt; uint32_t E;   // transition FEATURES: per-frame table [N][E], E = P*P + 2L: crossT | negDiag | negOff of the frame (else nullptr)
	// phone-bigram language model (one state per phone; nullptr = the decoder's free-phone LM with weight 0 on every arc):
	// lm_start[P], lm_bigT[to * P + from] (TRANSPOSED like the shared-memory cross table), lm_final[P] (+inf = not a final state)
	const float* lm_start; const float* lm_bigT; const float* lm_final;
	double beam;                      // > 0: beam pruning (pruning(), .cpp:976-1106; one state per phone): a node keeps the hypotheses with weight < min + beam
	const float* lm_exit;             // N states per phone (lm_bigT then nullptr): cost of the epsilon arc phone state -> LM start state; lm_start = unigram costs
	float* candW; int32_t* candP;     // [n_utt][D][L] rings of candidates per start frame
	float* keptW;                     // [n_utt][2][L]
	uint16_t* bp; uint8_t* bd;        // [N][L] back pointer (label or 0xffff) and duration
	uint8_t* gmove;                   // [N] which phone was moved to the back of the kept list (0xff none)
	unsigned long long* dbg;          // CRFGPU_DP_TIMING: cycle counters of CTA 0 (else nullptr)
	uint32_t* out_lab; uint32_t* out_dur; uint32_t* out_phn; uint32_t* n_seg; float* cost;
};
void launch_viterbi(const VitParams& p, cudaStream_t s);
// frame_t / frame_utt / frame_len (any may be null) of a ragged batch from its device-resident utterance offsets [n_utt + 1]
void launch_frame_tables(const uint32_t* off, uint32_t n_utt, uint32_t N, uint32_t* ft, uint32_t* fu, uint32_t* fl, cudaStream_t s);

// ---- Viterbi for large phone sets, one state per phone: cross-phone table sliced over a group of CTAs (crf_viterbi_group.cu) ----
constexpr int VITG_UT = 16;           // utterances that advance in lock-step per group
struct VitGroupParams {
	uint32_t n_utt, P, D;             // one state per phone: labels = phones
	uint32_t n_batches, n_groups, npt;  // batches of VITG_UT utterances dealt to the groups; CTAs per group = ceil(P / 32)
	const uint32_t* off;
	const uint32_t* slot_utt;         // [n_batches][VITG_UT] utterance ids or 0xffffffff
	const float* negS;                // [N][D][P]
	const float* crossT;              // [P][P]
	const float* negDiag;             // [P]
	float2* cand;                     // [n_utt][D][P] ring of candidates per start frame: (cost, back pointer as int bits)
	uint16_t* bp; uint8_t* bd;        // [N][P]
	float* xch;                       // [n_groups][2][Pk][VITG_UT] kept costs of the frame, Pk = P rounded up to 32
	float* finalW;                    // [n_utt][P] kept costs of every utterance's last frame
	uint32_t* ctr;                    // [n_groups] arrival counters
	unsigned long long* dbg;          // CRFGPU_DP_TIMING: cycle counters of CTA 0 (else nullptr)
	uint32_t* out_lab; uint32_t* out_dur; uint32_t* out_phn; uint32_t* n_seg; float* cost;
};
size_t vitg_smem_bytes(uint32_t P);
int vitg_max_groups(uint32_t P);      // groups of ceil(P/32) CTAs that are co-resident (0: the slices do not fit shared memory)
cudaError_t launch_viterbi_group(const VitGroupParams& p, cudaStream_t s);

// ---- frame-level CRF with transition FEATURES (stdtrans, one state per label; crf_dp_transftr.cu) ---------------------
struct TransFtrParams {
	uint32_t L, Lp, Lq;           // labels (<= 128), row stride of S / A / Dm, row stride of M / Xd (>= L*L)
	uint32_t n_utt; const uint32_t* off;
	const float* S; const float* M;       // [N][Lp] state scores, [N][Lq] transition scores M[n][p*L + c] of the frame the arc ENTERS
	const float* E; const float* rowmax;  // [N][Lq] exp(M_n - max M_n) and [N] that maximum (launch_transftr_exp)
	float* A; double* rho;                // forward: alpha normalised to sum 1, its log scale
	double* logZ; double* numer;          // [n_utt]
	float* Dm; float* Xd;                 // backward: [ref] - gamma [N][Lp], [ref pair] - xi [N][Lq] (zero on the first frame of an utterance)
	const uint32_t* labs;                 // [N] reference label per frame
	const uint32_t* tidx;                 // [L][L] lambda index of the pair, 0xffffffff on the illegal pairs of an N-state map (their M is -inf)
};
// segmental model without duration labels + transition features (stdseg_no_dur_no_segtransftr with stdtrans)
struct NodurTfParams {
	uint32_t P, Pp, D; uint64_t Lp; uint32_t Lq;   // phones (<= 128), stride of the P-wide arrays, durations (<= 31), stride of S / Dm, stride of M / Xd
	uint32_t n_utt; const uint32_t* off;
	const float* S; const float* M;               // [N][Lp] scores of (d,y); [N][Lq] M_n[y'][y] from frame n's duration-1 window
	const float* E; const float* rowmax;          // [N][Lq] exp(M_n - max M_n) and [N] that maximum (launch_transftr_exp)
	float* A; float* LG; double* rho;             // forward: alpha_t (sum 1), log A_t[y] - rho_t, log scale
	double* logZ; double* numer;
	float* Dm; float* Xd;                         // [ref] - gamma [N][Lp] (written by launch_nodur_post); backward: [ref pair] - xi stored at the frame the new segment starts in
	float* LB; double* kappa;                     // backward: beta_t[y] - kappa_t [N][Pp] and kappa_t [N] for the posterior pass
	const uint32_t* node_lab; const uint32_t* next_lab;   // (dur-1)*P + phone where a reference segment ends; phone of the NEXT reference segment there
	const uint32_t* tidx;                         // [P][P] lambda index of the pair, 0xffffffff on the illegal pairs of an N-state map (their M is -inf)
};
size_t nodur_tf_smem_bytes(uint32_t P);
cudaError_t launch_nodur_tf_dp(bool backward, const NodurTfParams& p, cudaStream_t s);
size_t transftr_smem_bytes(uint32_t L);
cudaError_t launch_transftr_dp(bool backward, const TransFtrParams& p, cudaStream_t s);
// E_n = exp(M_n - max M_n), rowmax[n] = max M_n for every frame (rows of LL entries, stride Lq): the exp of the recursions, taken off their chains
void launch_transftr_exp(const float* M, float* E, float* rowmax, uint32_t N, uint32_t LL, uint32_t Lq, cudaStream_t s);

// ---- lambda-derived tables and the trainer's update on the device (crf_lambda.cu) -----------------------------------
struct LambdaTablesParams {
	const double* lam;
	// training tables over the lattice labels (null Ws: skip)
	uint32_t Le, Lpe, nSf;                 // rows of the tables, row stride of E / ET, state features per label
	const uint32_t* sidx; const uint32_t* tidx;   // [>= Le] state block offsets, [Le*Le] transition indices of the lattice labels
	int use_state_bias, use_trans_bias; double state_bias_val, trans_bias_val;
	float* Ws; float* bias; float* E; float* ET;  // E / ET must be zero before the launch
	double* tmax;                          // device scalar: max transition score (read back by the host as Mmax)
	unsigned char* Wt; uint32_t wt_P, wt_D, wt_chunks;   // weight tiles of the TMA-fed score GEMM (null: skip)
	// virtual windows (wt_virt): chunk c < 5*wt_cpb holds features [32 (c % cpb), +32) of sampled-frame block c / cpb (zero beyond F), the
	// chunks behind them the aggregate columns 5F ...; bias_dy[d*wt_P + y] = state bias + the one-hot duration weight lambda[sidx + 8F + d]
	uint32_t wt_virt, wt_cpb, wt_F, wt_Dd; float* bias_dy;
	// transition-FEATURE weights (stdtrans, null Wtr: skip): Wtr[(p*L0 + c)][f] = lambda[tidx0(p,c) + f], tbias = bias weight * transBiasVal
	float* Wtr; float* tbias; uint32_t nTf;
	// decoder transition-FEATURE weights (null WdT: skip): WdT[f*vtE + e] = lambda[vt_base[e] + f] for the vtE = P*P + 2L decoder pairs
	double* WdT; const uint32_t* vt_base; uint32_t vtE;
	// decoder tables over the model's own labels (null Wd: skip)
	double* Wd; float* crossT; float* negDiag; float* negOff;
	const uint32_t* sidx0; const uint32_t* tidx0; uint32_t L0, NS, P0;
};
cudaError_t launch_lambda_tables(const LambdaTablesParams& p, cudaStream_t s);
struct SgdParams {
	double* lambda; const double* grad; uint64_t len;
	double n_active, lr; int use_gvar; double inv_square_var; int use_adagrad; double eta, eps;
	double* grad_sqr_acc; double* lambda_acc; double* lambda_sqr_acc;   // optional (AdaGrad state, averaged-model accumulators)
};
cudaError_t launch_sgd_update(const SgdParams& p, cudaStream_t s);

// ---- posterior-mass assertion of computeExpF (CRF_StdStateNode.cpp:252-275, CRF_StdSegStateNode.cpp:417-436) ----
void launch_posterior_mass(const float* Dm, uint64_t ld, uint32_t L, const uint32_t* node_lab, uint32_t N, float lo, float hi, float* mass,
                           double* n_bad, cudaStream_t s);

// ---- small helpers ------------------------------------------------------------------------------
void launch_fill_f32(float* p, uint64_t n, float v, cudaStream_t s);
void launch_fill_f64(double* p, uint64_t n, double v, cudaStream_t s);
// alpha/beta in the log domain for tests: alpha = m + log A, beta = bbase + log U (LOG0 where undefined)
void launch_dump_alpha_beta(const DpParams& p, uint32_t N, const uint32_t* frame_t, const uint32_t* frame_len,
                            double* alpha, double* beta, cudaStream_t s);

}  // namespace crfgpu
