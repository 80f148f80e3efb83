"""GPU parity tests (run on the B200 with -m gpu): the CUDA path, called through the C ABI
(libcrfgpu.so via crf_b200), against (a) golden vectors produced by the unmodified reference and
(b) the C oracle on fresh seeded inputs.

Tolerances (north_star: "fp32 relative tolerance, e.g. 1e-4"):
  logZ, numerator : |gpu - ref| <= 1e-5 * max(1, |ref|)
  gradient        : |gpu - ref| <= 1e-4 * max(|ref|) + 1e-4 * |ref|  (element-wise; the device computes
                    scores and posteriors in fp32, the reference in fp64)
  alpha / beta    : |gpu - ref| <= 1e-4 * max(1, |ref|) on every lattice entry whose posterior is
                    above 1e-30 (entries below the fp32 range flush to -inf on the device)
  Viterbi         : labels, durations, emitted phones and the float path cost are bit-exact.
"""
import numpy as np
import pytest

import crf_b200
from oracle.binding import make_config
from helpers import load_cases, split_segs, synth_batch

pytestmark = pytest.mark.gpu

TRAIN = load_cases("train_golden.npz")
VIT = load_cases("viterbi_golden.npz")
WIN = load_cases("window_golden.npz")
NODUR = load_cases("train_nodur_golden.npz")
NSTATE = load_cases("train_nodur_nstate_golden.npz")
TRANSFTR = load_cases("train_transftr_golden.npz")
TRANSFTR_NS = load_cases("train_transftr_nstate_golden.npz")
VIT_TF = load_cases("viterbi_transftr_golden.npz")
JOINED = load_cases("joined_golden.npz")
VIT_LM = load_cases("viterbi_lm_golden.npz")
VIT_BEAM = {k: v for k, v in VIT_LM.items() if k.startswith("beam")}
VIT_LM = {k: v for k, v in VIT_LM.items() if not k.startswith("beam")}


def gpu(cfg):
    return crf_b200.CrfGpu(crf_b200.copy_config(cfg))


def assert_train_close(got, want, what=""):
    g, n, z = got
    gw, nw, zw = want
    np.testing.assert_allclose(z, zw, rtol=1e-5, atol=1e-5, err_msg=f"logZ {what}")
    np.testing.assert_allclose(n, nw, rtol=1e-5, atol=1e-5, err_msg=f"numerator {what}")
    tol = 1e-4 * np.abs(gw).max() + 1e-4 * np.abs(gw)
    bad = np.abs(g - gw) > tol
    assert not bad.any(), f"gradient {what}: {bad.sum()} of {bad.size} entries off, worst {np.abs(g - gw).max():.3e} (|g|max {np.abs(gw).max():.3e})"


# lattice kernels: 0 = one CTA per utterance group (E from L2), 1 = cluster-resident E with FFMA, 2 = cluster-resident E with
# tcgen05, output labels sliced over the cluster, 3 = tcgen05, contraction index sliced, E whole in tensor memory (default); GEMMs: 0 = fp32 FFMA tiles, 1 = tcgen05 split-bf16 with register-staged operands, 2 = 1 + TMA-fed window GEMMs (default)
IMPLS = {"tc": {}, "tc_msplit": {"dp_impl": 2}, "tc_frame_lattice": {"frame_impl": 1}, "tc_msplit_frame_lattice": {"dp_impl": 2, "frame_impl": 1}, "frame_sequential": {"frame_impl": 2}, "tc_ffma_gemm": {"gemm_impl": 0}, "tc_reg_gemm": {"gemm_impl": 1}, "tc_tmem_all": {"tma_mask": 63}, "tc_smem_all": {"tma_mask": 7}, "cluster": {"dp_impl": 1}, "cluster_u4": {"dp_impl": 1, "cluster_slots": 4},
         "legacy_u1": {"dp_impl": 0, "slots": 1}, "legacy_u4": {"dp_impl": 0, "slots": 4, "gemm_impl": 0}}


@pytest.mark.parametrize("name", sorted(TRAIN))
@pytest.mark.parametrize("impl", sorted(IMPLS))
def test_fwdbwd_matches_reference_golden(name, impl):
    c = TRAIN[name]
    m = gpu(c["cfg"])
    assert m.lambda_len == len(c["lam"])
    for k, v in IMPLS[impl].items():
        m.set_option(k, v)
    m.set_lambda(c["lam"])
    got = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name)
    m.close()


@pytest.mark.parametrize("name", sorted(NODUR))
@pytest.mark.parametrize("impl", ["tc", "cluster", "legacy_u4"])
def test_fwdbwd_nodur_matches_reference_golden(name, impl):
    """stdseg_no_dur* training (the reference's CRF_StdSegStateNode_WithoutDurLab* nodes): the device runs the stdseg
    recursions on the (duration, phone) label set with tied weights."""
    c = NODUR[name]
    m = gpu(c["cfg"])
    assert m.lambda_len == len(c["lam"])
    for k, v in IMPLS[impl].items():
        m.set_option(k, v)
    m.set_lambda(c["lam"])
    got = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name)
    m.close()


@pytest.mark.parametrize("name", sorted(NODUR))
def test_fwdbwd_nodur_native_matches_reference_golden(name):
    """The same goldens through the native O(P^2 + D*P) recursion (crf_dp_nodur.cu), which is what large phone sets run on."""
    c = NODUR[name]
    m = gpu(c["cfg"])
    m.set_option("nodur_impl", 1)
    m.set_lambda(c["lam"])
    got = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name)
    m.close()


@pytest.mark.parametrize("P,D,F,n_utt,t_lo,t_hi,scale", [(48, 10, 12, 40, 5, 90, 0.05), (200, 12, 16, 21, 30, 70, 0.02), (1024, 30, 8, 3, 20, 45, 0.01)])
def test_fwdbwd_nodur_native_matches_oracle_fresh(oracle, P, D, F, n_utt, t_lo, t_hi, scale):
    """stdseg_no_dur_no_segtransftr up to the cfg5 geometry (1024 phones, maxDur 30) against the oracle's native restatement;
    several lock-step batches per group, ragged lengths, phone counts that are not multiples of the 32-phone tiles."""
    rng = np.random.default_rng(P + D)
    off, ftrs, labs = synth_batch(rng, n_utt, t_lo, t_hi, F, P, 1, D + 3)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=F, max_dur=D, n_actual_labs=P, extract_seg_ftrs=1)
    lam = rng.uniform(-scale, scale, oracle.lambda_len(cfg))
    want = oracle.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=8)
    m = gpu(cfg)
    m.set_option("nodur_impl", 1)
    m.set_lambda(lam)
    got = m.fwdbwd(off, ftrs, labs)
    assert_train_close(got, want, f"P={P} D={D}")
    m.close()


def test_toy_known_answers_on_gpu():
    c = TRAIN["toy_stdframe"]
    m = gpu(c["cfg"])
    m.set_lambda(c["lam"])
    g, n, z = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    np.testing.assert_allclose(z, [5.488352272567, 4.152084629056, 5.488352272567], rtol=1e-6)
    np.testing.assert_allclose(n, [-0.09, -0.01, -0.09], atol=1e-7)
    assert abs(np.sum(g ** 2) - 41.991090456415) < 1e-3


@pytest.mark.parametrize("kind", ["frame61", "frame3state", "stdseg_small", "stdseg_cfg4_shape"])
def test_fwdbwd_matches_oracle_fresh(oracle, kind):
    rng = np.random.default_rng({"frame61": 1, "frame3state": 2, "stdseg_small": 3, "stdseg_cfg4_shape": 4}[kind])
    if kind == "frame61":      # cfg2 geometry, short utterances
        off, ftrs, labs = synth_batch(rng, 9, 20, 120, 105, 61, 2, 15)
        cfg = make_config("stdframe", n_labs=61, n_base_ftrs=105)
        scale = 0.25
    elif kind == "frame3state":
        off, ftrs, labs = synth_batch(rng, 5, 10, 60, 20, 12, 3, 9, states=3)
        cfg = make_config("stdframe", n_labs=36, n_base_ftrs=20, n_states=3)
        scale = 0.25
    elif kind == "stdseg_small":
        off, ftrs, labs = synth_batch(rng, 7, 1, 50, 8, 6, 1, 9)
        cfg = make_config("stdseg", n_labs=6 * 5, n_base_ftrs=8, max_dur=5, n_actual_labs=6, extract_seg_ftrs=1)
        scale = 0.05
    else:                       # cfg4 geometry (61 phones x maxDur 10, 105 base features), tiny batch
        off, ftrs, labs = synth_batch(rng, 3, 12, 40, 105, 61, 2, 14)
        cfg = make_config("stdseg", n_labs=610, n_base_ftrs=105, max_dur=10, n_actual_labs=61, extract_seg_ftrs=1)
        scale = 0.01
    lam = rng.uniform(-scale, scale, oracle.lambda_len(cfg))
    want = oracle.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=4)
    m = gpu(cfg)
    m.set_lambda(lam)
    got = m.fwdbwd(off, ftrs, labs)
    assert_train_close(got, want, kind)


def test_alpha_beta_match_oracle(oracle):
    import ctypes as C
    rng = np.random.default_rng(7)
    off, ftrs, labs = synth_batch(rng, 1, 37, 37, 8, 6, 1, 9)
    cfg = make_config("stdseg", n_labs=6 * 4, n_base_ftrs=8, max_dur=4, n_actual_labs=6, extract_seg_ftrs=1)
    lam = rng.uniform(-0.05, 0.05, oracle.lambda_len(cfg))
    T, L = int(off[-1]), 24
    a = np.zeros((T, L)); b = np.zeros((T, L)); g = np.zeros(len(lam)); nu = np.zeros(1); z = np.zeros(1)
    P = lambda x, t: x.ctypes.data_as(C.POINTER(t))
    rc = oracle.lib.crforacle_fwdbwd_dump(C.byref(cfg), P(lam, C.c_double), C.c_uint32(len(lam)), C.c_uint32(T),
                                          P(ftrs, C.c_float), P(labs, C.c_uint32), P(g, C.c_double), P(nu, C.c_double),
                                          P(z, C.c_double), P(a, C.c_double), P(b, C.c_double))
    assert rc == 0
    m = gpu(cfg)
    m.set_option("keep_lattice", 1)
    m.set_lambda(lam)
    m.stage(off, ftrs, labs)
    m.fwdbwd_staged()
    ga, gb = m.fetch_alpha_beta()
    defined = a > -1e300
    assert np.array_equal(defined, ga > -1e300)
    post = np.where(defined & (b > -1e300), a + b - z[0], -np.inf)
    live = post > np.log(1e-30)
    np.testing.assert_allclose(ga[live], a[live], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(gb[live], b[live], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", sorted(VIT))
def test_viterbi_bit_exact_vs_reference_golden(name):
    c = VIT[name]
    m = gpu(c["cfg"])
    m.set_lambda(c["lam"])
    segs, cost = m.viterbi(c["off"], c["ftrs"])
    want = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])
    assert [len(s[0]) for s in segs] == [int(k) for k in c["nseg"]]
    for got, exp in zip(segs, want):
        assert np.array_equal(got[0], exp[0]), name
        assert np.array_equal(got[1], exp[1]), name
        assert np.array_equal(got[2], exp[2]), name
    assert np.array_equal(cost.view(np.uint32), c["cost"].view(np.uint32))
    m.close()


@pytest.mark.parametrize("P,N,D,segf,F", [(61, 3, 1, 0, 105), (61, 1, 1, 0, 105), (48, 1, 10, 1, 12), (20, 3, 4, 1, 9), (1024, 1, 30, 1, 8)])
def test_viterbi_bit_exact_vs_oracle_fresh(oracle, P, N, D, segf, F):
    """the last case is the cfg5 geometry (1024 phones, maxDur 30)"""
    rng = np.random.default_rng(P * 100 + N * 10 + D)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P * N, n_base_ftrs=F, n_states=N, max_dur=D,
                      extract_seg_ftrs=segf)
    lens = rng.integers(1, 120, 12)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    m = gpu(cfg)
    for lam_kind in ("random", "ties", "quantised"):
        n = oracle.lambda_len(cfg)
        ftrs = rng.random((int(off[-1]), F), dtype=np.float32)
        if lam_kind == "random":
            lam = rng.uniform(-0.25, 0.25, n)
        elif lam_kind == "ties":
            lam = np.zeros(n)
        else:
            lam = np.round(rng.uniform(-1, 1, n) * 2) / 2
            ftrs = (np.round(ftrs * 2) / 2).astype(np.float32)
        want, wcost, _ = oracle.viterbi(cfg, lam, off, ftrs)
        m.set_lambda(lam)
        segs, cost = m.viterbi(off, ftrs)
        for got, exp in zip(segs, want):
            assert all(np.array_equal(x, y) for x, y in zip(got, exp)), lam_kind
        assert np.array_equal(cost.view(np.uint32), wcost.view(np.uint32)), lam_kind


@pytest.mark.parametrize("name", sorted(k for k in WIN if k.startswith("win_")))
def test_window_expansion_bit_exact(name):
    c = WIN[name]
    m = gpu(c["cfg"])
    got = m.expand_windows(c["x"])
    exp = c["out"]
    valid = ~np.isnan(exp)
    assert np.array_equal(got[valid].view(np.uint32), exp[valid].view(np.uint32))
    assert np.all(got[~valid] == 0)


def test_unsupported_geometries_fail_loudly():
    with pytest.raises(crf_b200.CrfGpuError) as ei:
        m = gpu(make_config("stdseg", n_labs=10, n_base_ftrs=4, max_dur=2, n_actual_labs=5, use_trans_ftrs=1))
        m.set_lambda(np.zeros(m.lambda_len))
        m.fwdbwd([0, 3], np.zeros((3, 4), np.float32), np.zeros(3, np.uint32))
    assert ei.value.code == 2
    m = gpu(make_config("stdframe", n_labs=5, n_base_ftrs=4))
    m.set_lambda(np.zeros(m.lambda_len))
    with pytest.raises(crf_b200.CrfGpuError):   # empty utterance: reference throws "No features read from this sentence"
        m.fwdbwd([0, 0, 3], np.zeros((3, 4), np.float32), np.zeros(3, np.uint32))


def test_size_independent_properties_full_cfg2_shape():
    """At BASELINE cfg2 size (61 labels, 105 features, ~300-frame utterances): with lambda = 0 every path is
    equally likely, so logZ = T*log(61) - and the gradient's state-bias block sums to zero per frame
    (sum_c (onehot - gamma) = 0), a checksum that does not need the oracle.  The bound on that checksum is
    1e-5 of the summed magnitudes (2 per frame): the split-bf16 tensor-core GEMM carries ~16 mantissa bits and
    its fp32 accumulation chain is 2048 frames long."""
    rng = np.random.default_rng(5)
    off, ftrs, labs = synth_batch(rng, 64, 200, 400, 105, 61, 3, 20)
    cfg = make_config("stdframe", n_labs=61, n_base_ftrs=105)
    m = gpu(cfg)
    m.set_lambda(np.zeros(m.lambda_len))
    g, n, z = m.fwdbwd(off, ftrs, labs)
    T = np.diff(off.astype(np.int64))
    np.testing.assert_allclose(z, T * np.log(61.0), rtol=1e-6)
    assert np.all(n == 0)
    sidx, _ = m.index_maps()
    bias = g[sidx + 105]
    bound = 1e-5 * 2 * float(off[-1])
    assert abs(bias.sum()) < bound
    lam = rng.uniform(-0.25, 0.25, m.lambda_len)
    m.set_lambda(lam)
    g, n, z = m.fwdbwd(off, ftrs, labs)
    assert abs(g[sidx + 105].sum()) < bound
    assert np.all(n - z < 0)    # log-likelihood of the reference path is negative


def _uniform_segment_stats(T, P, D):
    """lambda = 0: log(number of labelled segmentations of T frames with segments <= D, P labels each) and the expected
    number of segments under the uniform distribution over them (log-domain DP, O(T*D))."""
    logz = np.full(T + 1, -np.inf)
    logz[0] = 0.0
    lp = np.log(float(P))
    for t in range(1, T + 1):
        prev = logz[max(0, t - D):t]
        mx = prev.max()
        logz[t] = lp + mx + np.log(np.exp(prev - mx).sum())
    exp_segs = 0.0
    for t in range(T):            # segment ending at frame t (inclusive) with duration d
        d = np.arange(1, min(t + 1, D) + 1)
        exp_segs += np.exp(lp + logz[t + 1 - d] + logz[T - 1 - t] - logz[T]).sum()
    return logz[T], exp_segs


def test_size_independent_properties_full_cfg5():
    """BASELINE cfg5 at full size (64 utterances x 2000 frames, 1024 phones, maxDur 30, 542 segment features) on the native
    no_dur recursion.  lambda = 0: logZ is the log-count of labelled segmentations, the state-bias gradient of phone y is
    (#reference segments of y) - E[#segments]/P and the transition-bias gradients sum to (#ref transitions) - (E[#segments] -
    #utterances), all from a host DP that needs no oracle.  Random lambda: sum of state-bias gradients - sum of transition-bias
    gradients = 0 (#segments - #transitions = #utterances for the reference path and in expectation) and numerator < logZ."""
    import workloads
    P, D = 1024, 30
    off, ftrs, labs = workloads.cfg5_batch()
    cfg = crf_b200.make_config(**workloads.cfg5_kwargs())
    m = crf_b200.CrfGpu(cfg)
    n_utt, T = len(off) - 1, 2000
    m.set_lambda(np.zeros(m.lambda_len))
    g, n, z = m.fwdbwd(off, ftrs, labs)
    logz, exp_segs = _uniform_segment_stats(T, P, D)
    np.testing.assert_allclose(z, logz, rtol=1e-6)
    assert np.all(n == 0)
    sidx, tidx = m.index_maps()
    nS = 8 * 64 + D
    ref_cnt = np.zeros(P)
    n_ref = 0
    for u in range(n_utt):
        r = m.group_labels(labs[off[u]:off[u + 1]])
        ends = r[:, 0] != 0xffffffff
        np.add.at(ref_cnt, r[ends, 0].astype(np.int64), 1.0)
        n_ref += int(ends.sum())
    bias = g[sidx + nS]
    np.testing.assert_allclose(bias, ref_cnt - n_utt * exp_segs / P, atol=2e-2)
    assert abs(bias.sum() - (n_ref - n_utt * exp_segs)) < 1e-4 * n_utt * exp_segs
    tsum = g[tidx.reshape(-1)].sum()
    assert abs(tsum - ((n_ref - n_utt) - (n_utt * exp_segs - n_utt))) < 1e-4 * n_utt * exp_segs
    lam = workloads.lam_for("cfg5", m.lambda_len)
    m.set_lambda(lam)
    g, n, z = m.fwdbwd(off, ftrs, labs)
    assert np.all(np.isfinite(z)) and np.all(n - z < 0)
    assert abs(g[sidx + nS].sum() - g[tidx.reshape(-1)].sum() - 0.0) < 1e-4 * n_utt * exp_segs
    m.close()


@pytest.mark.parametrize("adagrad", [0, 1])
def test_sgd_update_on_device_matches_reference_rule(adagrad):
    """crfgpu_sgd_update restates CRF_SGTrainer.cpp:299-325 (grad / n_active, the gvar quirk, SGD or AdaGrad, lambdaAcc /
    lambdaSqrAcc) on the device and rebuilds every lambda-derived table by a kernel: after two updates lambda and the
    accumulators equal the host rule applied to the same gradients to 1 ulp, and the next gradient equals the one a fresh handle
    computes from that lambda through crfgpu_set_lambda."""
    c = TRAIN["stdseg_d10_segftr"]
    m = gpu(c["cfg"])
    lam = c["lam"].copy()
    m.set_lambda(lam)
    n_act, lr, eta, eps, isv = 4.0, float(np.float32(0.1)), 0.05, 1e-6, 0.01
    acc = np.zeros_like(lam); sqr = np.zeros_like(lam); gsq = np.zeros_like(lam)
    for _ in range(2):
        g, _, _ = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
        m.sgd_update(n_act, lr=lr, use_gvar=1, inv_square_var=isv, use_adagrad=adagrad, eta=eta, eps=eps)
        gg = g / n_act
        gg = gg - gg * isv
        if adagrad:
            gsq = gsq + gg * gg
            lam = lam + eta / (np.sqrt(gsq) + eps) * gg
        else:
            lam = lam + lr * gg
        acc = acc + lam; sqr = sqr + lam * lam
    got, gacc, gsqr, ggsq = m.get_lambda(with_state=True)
    np.testing.assert_allclose(got, lam, rtol=1e-14, atol=1e-16)
    np.testing.assert_allclose(gacc, acc, rtol=1e-14, atol=1e-16)
    np.testing.assert_allclose(gsqr, sqr, rtol=1e-13, atol=1e-16)
    if adagrad:
        np.testing.assert_allclose(ggsq, gsq, rtol=1e-13, atol=1e-300)
    g_dev = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    m2 = gpu(c["cfg"])
    m2.set_lambda(got)
    g_host = m2.fwdbwd(c["off"], c["ftrs"], c["labs"])
    for a, b in zip(g_dev, g_host):      # same tables -> same results up to the order of the fp64 atomics of the GEMM epilogues
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-9 * np.abs(b).max())
    m.close(); m2.close()


def test_decode_tables_from_device_lambda_bit_exact(oracle):
    """Viterbi after an on-device update decodes exactly what the oracle decodes from the fetched lambda."""
    rng = np.random.default_rng(11)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=12, n_base_ftrs=9, max_dur=4, extract_seg_ftrs=1)
    off, ftrs, labs = synth_batch(rng, 6, 10, 60, 9, 12, 1, 6)
    m = gpu(cfg)
    m.set_lambda(rng.uniform(-0.3, 0.3, m.lambda_len))
    m.fwdbwd(off, ftrs, labs)
    m.sgd_update(6.0, lr=0.5)
    lam = m.get_lambda()
    segs, cost = m.viterbi(off, ftrs)
    want, wcost, _ = oracle.viterbi(cfg, lam, off, ftrs)
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), wcost.view(np.uint32))
    m.close()


@pytest.mark.parametrize("kind", ["stdseg", "nodur_native", "nodur_tied", "frame", "frame_transftr", "frame_transftr_even", "nodur_transftr"])
def test_fwdbwd_edge_lengths_match_oracle(oracle, kind):
    """Ragged edge cases the reference handles: utterances of 1, 2, 3 frames (shorter than maxDur, so most windows never exist),
    exactly maxDur frames, and a long one, in one batch; reference segments longer than maxDur are split by the label grouping."""
    rng = np.random.default_rng(17)
    lens = np.array([1, 2, 3, 1, 6, 7, 40, 2], np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    F, P, D = 7, (6 if kind in ("frame_transftr_even", "nodur_transftr") else 5), 6
    ftrs = rng.random((int(off[-1]), F), dtype=np.float32)
    labs = np.zeros(int(off[-1]), np.uint32)
    for u in range(len(lens)):
        t = int(off[u])
        while t < off[u + 1]:
            d = int(rng.integers(1, 12))
            labs[t:min(t + d, int(off[u + 1]))] = rng.integers(0, P)
            t += d
    if kind == "stdseg":
        cfg = make_config("stdseg", n_labs=P * D, n_base_ftrs=F, max_dur=D, n_actual_labs=P, extract_seg_ftrs=1)
    elif kind == "frame":
        cfg = make_config("stdframe", n_labs=P, n_base_ftrs=F)
    elif kind.startswith("frame_transftr"):      # (the bulk-copied matrix of the transition-feature recursions: no copy at all for one frame)
        cfg = make_config("stdframe", n_labs=P, n_base_ftrs=F, use_trans_ftrs=1)
    elif kind == "nodur_transftr":
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=F, max_dur=D, extract_seg_ftrs=1, use_trans_ftrs=1, trans_fidx=(0, F - 1))
    else:
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=F, max_dur=D, n_actual_labs=P, extract_seg_ftrs=1)
    lam = rng.uniform(-0.3, 0.3, oracle.lambda_len(cfg))
    want = oracle.fwdbwd(cfg, lam, off, ftrs, labs)
    m = gpu(cfg)
    if kind.startswith("nodur"):
        m.set_option("nodur_impl", 1 if kind == "nodur_native" else 2)
    m.set_lambda(lam)
    got = m.fwdbwd(off, ftrs, labs)
    assert_train_close(got, want, kind)
    m.close()


def test_empty_batch_is_a_no_op():
    cfg = make_config("stdseg", n_labs=12, n_base_ftrs=4, max_dur=3, n_actual_labs=4, extract_seg_ftrs=1)
    m = gpu(cfg)
    m.set_lambda(np.zeros(m.lambda_len))
    g, n, z = m.fwdbwd(np.zeros(1, np.uint32), np.zeros((0, 4), np.float32), np.zeros(0, np.uint32))
    assert not g.any() and len(n) == 0 and len(z) == 0
    m.close()


def test_prefetch_takes_over_buffers_and_falls_back():
    """crfgpu_prefetch_batch: a batch that was prefetched is staged by taking its buffers over (same results as a plain call);
    staging a DIFFERENT batch after a prefetch ignores the prefetch."""
    a, b = TRAIN["stdseg_d10_segftr"], TRAIN["stdseg_d4_segftr"]
    m = gpu(a["cfg"])
    m.set_lambda(a["lam"])
    want = m.fwdbwd(a["off"], a["ftrs"], a["labs"])
    fa = np.ascontiguousarray(a["ftrs"], np.float32)
    for _ in range(3):                                # steady state: prefetch the next copy of the batch while this one computes
        m.stage(a["off"], fa, a["labs"])
        m.fwdbwd_staged()
        m.prefetch(a["off"], fa)
        got = m.fetch_fwdbwd()
        for x, y in zip(got, want):
            np.testing.assert_allclose(x, y, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(y).max()))
    m.prefetch(a["off"], fa)
    shifted = np.ascontiguousarray(fa[::-1])          # other contents at another address: must be staged normally
    got = m.fwdbwd(a["off"], shifted, a["labs"])
    assert not np.allclose(got[2], want[2])
    got = m.fwdbwd(a["off"], fa, a["labs"])
    for x, y in zip(got, want):
        np.testing.assert_allclose(x, y, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(y).max()))
    m.close()
    assert b is not None


@pytest.mark.parametrize("case", ["stdseg_d10_segftr", "frame_1state"])
def test_prefetch_train_batch_hands_over_label_tables(case):
    """crfgpu_prefetch_train_batch: the read-ahead also builds and copies the label tables; the staged batch gives the plain call's
    results, and a batch staged with OTHER labels (another array) rebuilds its tables instead of taking the prefetched ones."""
    a = TRAIN[case]
    m = gpu(a["cfg"])
    m.set_lambda(a["lam"])
    want = m.fwdbwd(a["off"], a["ftrs"], a["labs"])
    fa, la = np.ascontiguousarray(a["ftrs"], np.float32), np.ascontiguousarray(a["labs"], np.uint32)
    for _ in range(3):
        m.stage(a["off"], fa, la)
        m.fwdbwd_staged()
        m.prefetch(a["off"], fa, la)
        got = m.fetch_fwdbwd()
        for x, y in zip(got, want):
            np.testing.assert_allclose(x, y, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(y).max()))
    # same features, different labels at another address: the prefetched label tables must not be used
    m.prefetch(a["off"], fa, la)
    lb = la.copy(); lb[:] = la[::-1]
    m.stage(a["off"], fa, lb); m.fwdbwd_staged()
    got_b = m.fetch_fwdbwd()
    want_b = m.fwdbwd(a["off"], a["ftrs"], lb)
    for x, y in zip(got_b, want_b):
        np.testing.assert_allclose(x, y, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(y).max()))
    assert not np.allclose(got_b[1], want[1])
    m.close()


@pytest.mark.parametrize("name", ["train_nodur_joined_recipe", "train_frame_joined_transftr", "train_stdseg_ctx"])
def test_prefetch_train_batch2_joined_streams(name):
    """crfgpu_prefetch_train_batch2: the read-ahead of a model with context frames / a joined second stream -- both streams, the joined
    windows and the label tables on the side stream; the staged batch gives the reference's golden, a batch at other addresses (or
    with other labels) is staged normally."""
    c = JOINED[name]
    f2 = f2_of(c)
    m = gpu(c["cfg"])
    m.set_lambda(c["lam"])
    fa, la = np.ascontiguousarray(c["ftrs"], np.float32), np.ascontiguousarray(c["labs"], np.uint32)
    fb = None if f2 is None else np.ascontiguousarray(f2, np.float32)
    if fb is None:      # (context frames only: the read-ahead entry point takes the first stream alone)
        fb = np.zeros((1, 1), np.float32)
    two = c["cfg"].n_base_ftrs2 > 0
    for _ in range(3):
        m.stage(c["off"], fa, la, ftrs2=fb if two else None)
        m.fwdbwd_staged()
        m.prefetch(c["off"], fa, la, ftrs2=fb)
        got = m.fetch_fwdbwd()
        assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name + " read ahead")
    # other labels at another address: the windows are taken over, the label tables rebuilt
    m.prefetch(c["off"], fa, la, ftrs2=fb)
    lb = la.copy()
    got = m.fwdbwd(c["off"], fa, lb, ftrs2=fb if two else None)
    assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name + " other label array")
    # other features at another address: staged normally
    m.prefetch(c["off"], fa, la, ftrs2=fb)
    got = m.fwdbwd(c["off"], fa.copy(), la, ftrs2=fb if two else None)
    assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name + " other feature array")
    m.close()


@pytest.mark.parametrize("name", sorted(TRANSFTR))
def test_fwdbwd_transition_features_match_reference_golden(name):
    """crf_featuremap=stdtrans on frame-level models: transition scores as a tensor-core GEMM, streamed recursions, both gradients
    as reduce-GEMMs (crf_dp_transftr.cu), against goldens produced by the reference."""
    c = TRANSFTR[name]
    m = gpu(c["cfg"])
    assert m.lambda_len == len(c["lam"])
    m.set_lambda(c["lam"])
    got = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name)
    m.close()


@pytest.mark.parametrize("name", sorted(TRANSFTR_NS))
def test_fwdbwd_transition_features_nstate_match_reference_golden(name):
    """crf_featuremap=stdtrans with N states per label (frame-level CRF_StdNStateNode and the segmental
    CRF_StdSegNStateNode_WithoutDurLab_WithoutSegTransFtr): the illegal pairs of the N-state map score -inf in the per-frame transition
    matrices and carry no gradient rows; against goldens produced by the reference."""
    c = TRANSFTR_NS[name]
    m = gpu(c["cfg"])
    assert m.lambda_len == len(c["lam"])
    m.set_lambda(c["lam"])
    got = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name)
    m.close()


@pytest.mark.parametrize("name", ["frame_transftr_20labs", "nodur_transftr_d10"])
def test_fwdbwd_transition_features_register_staged_gemms(name):
    """option tf_tiled 0: the labels^2 GEMMs on the register-staged kernels (fp32 operands) instead of the pre-tiled bf16 operands fed
    by bulk copies -- both routes against the reference's golden"""
    c = TRANSFTR[name]
    m = gpu(c["cfg"])
    m.set_option("tf_tiled", 0)
    m.set_lambda(c["lam"])
    got = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name + " tf_tiled=0")
    m.set_option("tf_tiled", 1)
    m.set_lambda(c["lam"])
    got = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name + " tf_tiled=1")
    m.close()


def test_fwdbwd_nodur_transition_features_timit_recipe_shape(oracle):
    """stdseg_no_dur_no_segtransftr + stdtrans at the shape of the production TIMIT recipe (48 phones, maxDur 10, transition features
    from the duration-1 window), against the oracle's native restatement (pinned to goldens from the reference's no_dur nodes)."""
    rng = np.random.default_rng(29)
    F, P, D = 13, 48, 10
    off, ftrs, labs = synth_batch(rng, 6, 5, 60, F, P, 1, 14)
    w = 8 * F + D
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=F, max_dur=D, n_actual_labs=P, extract_seg_ftrs=1,
                      use_trans_ftrs=1, trans_fidx=(0, 5 * F - 1), state_fidx=(0, w - 1))
    lam = rng.uniform(-0.02, 0.02, oracle.lambda_len(cfg))
    want = oracle.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=6)
    m = gpu(cfg)
    m.set_lambda(lam)
    got = m.fwdbwd(off, ftrs, labs)
    assert_train_close(got, want, "no_dur stdtrans")
    m.close()


def test_fwdbwd_transition_features_cfg2_shape_matches_oracle(oracle):
    """61 labels, 105 features for states and transitions (the cfg2 geometry with stdtrans: dim(lambda) = 61*(106 + 61*106))."""
    rng = np.random.default_rng(23)
    off, ftrs, labs = synth_batch(rng, 5, 8, 40, 105, 61, 2, 9)
    cfg = make_config("stdframe", n_labs=61, n_base_ftrs=105, use_trans_ftrs=1)
    lam = rng.uniform(-0.02, 0.02, oracle.lambda_len(cfg))
    want = oracle.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=4)
    m = gpu(cfg)
    m.set_lambda(lam)
    got = m.fwdbwd(off, ftrs, labs)
    assert_train_close(got, want, "stdtrans cfg2 shape")
    m.close()


@pytest.mark.parametrize("kind", ["frame_144", "frame_169", "nodur_144_d10", "nodur_161_d3", "frame_170_refused", "nodur_162_refused"])
def test_fwdbwd_transition_features_beyond_128_labels(oracle, kind):
    """stdtrans past 128 labels (48 phones x 3 states = 144 is the TIMIT recipe with crf_states 3): 192-thread CTAs, the label count
    bounded by the double-buffered L x L tile in shared memory (169 frame-level, 161 segmental); one more label is refused loudly."""
    rng = np.random.default_rng(144)
    F = 6
    if kind.endswith("refused"):
        cfg = (make_config("stdframe", n_labs=170, n_base_ftrs=F, use_trans_ftrs=1) if kind.startswith("frame") else
               make_config("stdseg_no_dur_no_segtransftr", n_labs=162, n_base_ftrs=F, max_dur=3, extract_seg_ftrs=1, use_trans_ftrs=1, trans_fidx=(0, 5)))
        m = gpu(cfg)
        m.set_lambda(np.zeros(m.lambda_len))
        off, ftrs, labs = synth_batch(rng, 2, 5, 9, F, 100, 1, 3)
        with pytest.raises(crf_b200.CrfGpuError) as e:
            m.fwdbwd(off, ftrs, labs)
        assert e.value.code == 2      # CRFGPU_ERR_UNSUPPORTED
        m.close()
        return
    if kind == "frame_144":
        off, ftrs, labs = synth_batch(rng, 4, 20, 60, F, 48, 3, 9, states=3)
        cfg = make_config("stdframe", n_labs=144, n_base_ftrs=F, n_states=3, use_trans_ftrs=1)
    elif kind == "frame_169":
        off, ftrs, labs = synth_batch(rng, 3, 10, 40, F, 169, 1, 5)
        cfg = make_config("stdframe", n_labs=169, n_base_ftrs=F, use_trans_ftrs=1, trans_fidx=(1, 4))
    elif kind == "nodur_144_d10":
        off, ftrs, labs = synth_batch(rng, 3, 30, 70, F, 48, 6, 30, states=3)
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=144, n_base_ftrs=F, n_states=3, max_dur=10, extract_seg_ftrs=1,
                          use_trans_ftrs=1, trans_fidx=(0, 5 * F - 1))
    else:
        off, ftrs, labs = synth_batch(rng, 3, 10, 40, F, 161, 1, 6)
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=161, n_base_ftrs=F, max_dur=3, extract_seg_ftrs=1, use_trans_ftrs=1, trans_fidx=(0, 11))
    lam = rng.uniform(-0.05, 0.05, oracle.lambda_len(cfg))
    want = oracle.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=4)
    m = gpu(cfg)
    m.set_lambda(lam)
    got = m.fwdbwd(off, ftrs, labs)
    assert_train_close(got, want, kind)
    m.close()


@pytest.mark.parametrize("name", sorted(VIT_TF))
def test_viterbi_transition_features_bit_exact_vs_reference_golden(name):
    """decoding with crf_featuremap=stdtrans: per-frame decoder tables from fp64 scores in the reference's order"""
    c = VIT_TF[name]
    m = gpu(c["cfg"])
    m.set_lambda(c["lam"])
    segs, cost = m.viterbi(c["off"], c["ftrs"])
    want = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])
    assert [len(s[0]) for s in segs] == [int(k) for k in c["nseg"]]
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp)), name
    assert np.array_equal(cost.view(np.uint32), c["cost"].view(np.uint32)), name
    m.close()


def test_viterbi_transition_features_recipe_shape_vs_oracle(oracle):
    """48 phones, maxDur 10, transition features from the duration-1 window (the TIMIT recipe's decode), bit-exact against the oracle"""
    rng = np.random.default_rng(31)
    F, P, D = 13, 48, 10
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=F, max_dur=D, extract_seg_ftrs=1,
                      use_trans_ftrs=1, trans_fidx=(0, 5 * F - 1))
    lens = rng.integers(1, 90, 9)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    ftrs = rng.random((int(off[-1]), F), dtype=np.float32)
    lam = rng.uniform(-0.25, 0.25, oracle.lambda_len(cfg))
    want, wcost, _ = oracle.viterbi(cfg, lam, off, ftrs)
    m = gpu(cfg)
    m.set_lambda(lam)
    segs, cost = m.viterbi(off, ftrs)
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), wcost.view(np.uint32))
    m.close()


@pytest.mark.parametrize("name", sorted(NSTATE))
@pytest.mark.parametrize("impl", ["native", "tied_tc", "tied_cluster"])
def test_fwdbwd_nodur_nstate_matches_reference_golden(name, impl):
    """N states per phone in stdseg_no_dur_no_[seg]transftr (CRF_StdSegNStateNode_WithoutDurLab_WithoutSegTransFtr): the native recursion
    and the tied (duration, label) expansion, both with E = 0 on the pairs the N-state map does not have."""
    c = NSTATE[name]
    m = gpu(c["cfg"])
    assert m.lambda_len == len(c["lam"])
    if impl == "native":
        m.set_option("nodur_impl", 1)
    else:
        m.set_option("nodur_impl", 2)
        for k, v in IMPLS[impl[5:]].items():
            m.set_option(k, v)
    m.set_lambda(c["lam"])
    got = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name)
    m.close()


@pytest.mark.parametrize("P,NS,D,F,n_utt", [(48, 3, 10, 12, 24), (61, 3, 6, 10, 9), (100, 2, 4, 8, 6)])
def test_fwdbwd_nodur_nstate_matches_oracle_fresh(oracle, P, NS, D, F, n_utt):
    """the TIMIT phone sets with 3 states per phone (144 / 183 labels) against the oracle, native recursion and (where it fits) tied"""
    rng = np.random.default_rng(P * NS + D)
    off, ftrs, labs = synth_batch(rng, n_utt, 8, 80, F, P, 3, 3 * D, states=NS)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P * NS, n_base_ftrs=F, n_states=NS, max_dur=D, extract_seg_ftrs=1)
    lam = rng.uniform(-0.05, 0.05, oracle.lambda_len(cfg))
    want = oracle.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=8)
    for impl in (1, 2):
        if impl == 2 and P * NS * D > 1024:
            continue
        m = gpu(cfg)
        m.set_option("nodur_impl", impl)
        m.set_lambda(lam)
        got = m.fwdbwd(off, ftrs, labs)
        assert_train_close(got, want, f"P={P} NS={NS} D={D} impl={impl}")
        m.close()


def _one_state_viterbi_goldens():
    return sorted(n for n, c in VIT.items() if c["cfg"].n_states == 1 and c["cfg"].n_labs >= 2 and not c["cfg"].use_trans_ftrs)


@pytest.mark.parametrize("name", _one_state_viterbi_goldens())
def test_viterbi_group_sliced_bit_exact_vs_reference_golden(name):
    """every one-state golden (ties, quantised weights, D = 1..3) through the group-sliced recursion (crf_viterbi_group.cu),
    which large phone sets run on by default"""
    c = VIT[name]
    m = gpu(c["cfg"])
    m.set_option("vit_impl", 2)
    m.set_lambda(c["lam"])
    segs, cost = m.viterbi(c["off"], c["ftrs"])
    want = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])
    assert [len(s[0]) for s in segs] == [int(k) for k in c["nseg"]]
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp)), name
    assert np.array_equal(cost.view(np.uint32), c["cost"].view(np.uint32)), name
    m.close()


@pytest.mark.parametrize("P,D,F,n_utt,t_hi", [(2, 1, 4, 5, 40), (33, 2, 6, 19, 70), (61, 10, 12, 40, 130), (200, 12, 8, 37, 90), (1024, 30, 8, 21, 140)])
def test_viterbi_group_sliced_matches_single_cta_and_oracle(oracle, P, D, F, n_utt, t_hi):
    """group-sliced recursion == one CTA per utterance == oracle, bit for bit: several batches of 16 per group, ragged lengths
    (including single frames), phone counts off the 32-phone tiles, random / all-ties / quantised weights"""
    rng = np.random.default_rng(P * 7 + D)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=F, max_dur=D, extract_seg_ftrs=1 if D > 1 else 0)
    lens = rng.integers(1, t_hi, n_utt)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    for lam_kind in ("random", "ties", "quantised"):
        n = oracle.lambda_len(cfg)
        ftrs = rng.random((int(off[-1]), F), dtype=np.float32)
        if lam_kind == "random":
            lam = rng.uniform(-0.25, 0.25, n)
        elif lam_kind == "ties":
            lam = np.zeros(n)
        else:
            lam = np.round(rng.uniform(-1, 1, n) * 2) / 2
            ftrs = (np.round(ftrs * 2) / 2).astype(np.float32)
        res = {}
        for impl in (1, 2):
            m = gpu(cfg)
            m.set_option("vit_impl", impl)
            m.set_lambda(lam)
            res[impl] = m.viterbi(off, ftrs)
            m.close()
        for a, b in zip(res[1][0], res[2][0]):
            assert all(np.array_equal(x, y) for x, y in zip(a, b)), lam_kind
        assert np.array_equal(res[1][1].view(np.uint32), res[2][1].view(np.uint32)), lam_kind
        if P <= 200 or lam_kind == "random":
            want, wcost, _ = oracle.viterbi(cfg, lam, off, ftrs)
            for got, exp in zip(res[2][0], want):
                assert all(np.array_equal(x, y) for x, y in zip(got, exp)), lam_kind
            assert np.array_equal(res[2][1].view(np.uint32), wcost.view(np.uint32)), lam_kind


def test_viterbi_multi_chunk_staging_bit_exact_vs_oracle(oracle):
    """a decode batch big enough for crfgpu_stage_batch to copy it in four chunks (>= 16384 frames): the decoder's fp64 scores are
    launched chunk by chunk behind the copies, and the paths must still be the oracle's bit for bit"""
    rng = np.random.default_rng(77)
    P, NS, D, F = 5, 3, 2, 6
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P * NS, n_base_ftrs=F, n_states=NS, max_dur=D, extract_seg_ftrs=1)
    lens = rng.integers(150, 400, 80)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    assert off[-1] >= 16384
    ftrs = rng.random((int(off[-1]), F), dtype=np.float32)
    lam = rng.uniform(-0.25, 0.25, oracle.lambda_len(cfg))
    want, wcost, _ = oracle.viterbi(cfg, lam, off, ftrs)
    m = gpu(cfg)
    m.set_lambda(lam)
    for _ in range(2):      # second call: the read-after-restage path
        segs, cost = m.viterbi(off, ftrs)
        for got, exp in zip(segs, want):
            assert all(np.array_equal(x, y) for x, y in zip(got, exp))
        assert np.array_equal(cost.view(np.uint32), wcost.view(np.uint32))
    # a new lambda after staging must not reuse the scores launched by the staging call
    lam2 = rng.uniform(-0.25, 0.25, oracle.lambda_len(cfg))
    want2, wcost2, _ = oracle.viterbi(cfg, lam2, off, ftrs)
    m.stage(off, ftrs)
    m.set_lambda(lam2)
    m.viterbi_staged()
    segs, cost = m.fetch_viterbi(off)
    for got, exp in zip(segs, want2):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), wcost2.view(np.uint32))
    # the recursion launched by viterbi_staged for the whole batch (option vit_eager 0) instead of per chunk on side streams
    m.set_option("vit_eager", 0)
    segs, cost = m.viterbi(off, ftrs)
    for got, exp in zip(segs, want2):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), wcost2.view(np.uint32))
    m.close()


def test_viterbi_eager_paths_dropped_when_beam_or_lm_change(oracle):
    """crfgpu_stage_batch launches the recursion of a decode batch chunk by chunk; a beam or a phone LM set between staging and
    crfgpu_viterbi_staged must not be answered with those paths"""
    rng = np.random.default_rng(78)
    P, D, F = 9, 2, 5
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=F, max_dur=D, extract_seg_ftrs=1)
    lens = rng.integers(150, 400, 70)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    assert off[-1] >= 16384
    ftrs = rng.random((int(off[-1]), F), dtype=np.float32)
    lam = rng.uniform(-0.5, 0.5, oracle.lambda_len(cfg))
    lm = (rng.uniform(0, 2, P).astype(np.float32), rng.uniform(0, 2, (P, P)).astype(np.float32), rng.uniform(0, 2, P).astype(np.float32))
    m = gpu(cfg)
    m.set_lambda(lam)
    m.stage(off, ftrs)
    m.set_phone_lm(*lm)
    m.viterbi_staged()
    segs, cost = m.fetch_viterbi(off)
    want, wcost, _ = oracle.viterbi(cfg, lam, off, ftrs, lm=lm)
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), wcost.view(np.uint32))
    m.stage(off, ftrs)
    m.set_beam(3.0)
    m.viterbi_staged()
    segs, cost = m.fetch_viterbi(off)
    want, wcost, _ = oracle.viterbi(cfg, lam, off, ftrs, lm=lm, beam=3.0)
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), wcost.view(np.uint32))
    # ... and a batch staged with both set takes the per-chunk route with them
    segs, cost = m.viterbi(off, ftrs)
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), wcost.view(np.uint32))
    m.close()


# ------------------------------------------------------------------------------------------------------------------
# Slot refill (continuous batching) of the lattice kernels.  A cluster's 16 slots each work through their own utterance
# list; lists hold more than one utterance only when the batch has more utterances than slots (240 on a B200 for cfg4), so
# "max_clusters" = 1 makes EVERY slot refill many times on a small batch.  Round 1 shipped a kernel whose bookkeeping warp
# read the old utterance's last scale from a ring entry the refilled utterance had already overwritten (wrong logZ whenever
# (len - 1) % 32 == 0) and no test reached the refill path: these do, with those lengths over-represented.
def _refill_lengths(rng, n_utt, t_hi):
    special = np.array([1, 1, 2, 33, 33, 65, 97, 129, 32, 34, 64, 66], np.int64)
    special = special[special <= t_hi]
    lens = rng.integers(1, t_hi + 1, n_utt)
    k = min(len(lens) // 2, 4 * len(special))
    lens[rng.permutation(n_utt)[:k]] = rng.choice(special, k)
    return lens


def _labelled_batch(rng, lens, F, P, seg_hi, states=1):
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    ftrs = rng.random((int(off[-1]), F), dtype=np.float32)
    labs = np.zeros(int(off[-1]), np.uint32)
    for u in range(len(lens)):
        t, prev = int(off[u]), -1
        while t < off[u + 1]:
            d = int(rng.integers(1, seg_hi + 1))
            lab = int(rng.integers(0, P))
            while lab == prev and P > 1:
                lab = int(rng.integers(0, P))
            e = min(t + d, int(off[u + 1]))
            if states == 1:
                labs[t:e] = lab
            else:
                n = e - t
                labs[t:e] = lab * states + np.minimum(np.arange(n) * states // max(n, 1), states - 1)
            prev, t = lab, e
    return off, ftrs, labs


REFILL_KINDS = {
    # kind: (config kwargs, options, n_utt, t_hi, lambda scale)
    "stdseg_tc": (dict(model_type="stdseg", n_labs=12, n_base_ftrs=5, max_dur=3, n_actual_labs=4, extract_seg_ftrs=1), {}, 640, 200, 0.3),
    "stdseg_tc_long": (dict(model_type="stdseg", n_labs=12, n_base_ftrs=5, max_dur=3, n_actual_labs=4, extract_seg_ftrs=1), {}, 96, 800, 0.3),
    "stdseg_cluster_ffma": (dict(model_type="stdseg", n_labs=12, n_base_ftrs=5, max_dur=3, n_actual_labs=4, extract_seg_ftrs=1), {"dp_impl": 1}, 640, 200, 0.3),
    "stdseg_tc_msplit": (dict(model_type="stdseg", n_labs=12, n_base_ftrs=5, max_dur=3, n_actual_labs=4, extract_seg_ftrs=1), {"dp_impl": 2}, 640, 200, 0.3),
    "stdseg_d10_msplit": (dict(model_type="stdseg", n_labs=50, n_base_ftrs=6, max_dur=10, n_actual_labs=5, extract_seg_ftrs=1), {"dp_impl": 2}, 320, 150, 0.1),
    "stdseg_200labels": (dict(model_type="stdseg", n_labs=200, n_base_ftrs=4, max_dur=4, n_actual_labs=50, extract_seg_ftrs=1), {}, 96, 60, 0.1),
    "stdseg_200labels_msplit": (dict(model_type="stdseg", n_labs=200, n_base_ftrs=4, max_dur=4, n_actual_labs=50, extract_seg_ftrs=1), {"dp_impl": 2}, 96, 60, 0.1),
    "stdseg_d10": (dict(model_type="stdseg", n_labs=50, n_base_ftrs=6, max_dur=10, n_actual_labs=5, extract_seg_ftrs=1), {}, 320, 150, 0.1),
    "nodur_tied": (dict(model_type="stdseg_no_dur_no_segtransftr", n_labs=6, n_base_ftrs=5, max_dur=4, n_actual_labs=6, extract_seg_ftrs=1), {"nodur_impl": 2}, 400, 150, 0.3),
    "nodur_nstate_tied": (dict(model_type="stdseg_no_dur_no_segtransftr", n_labs=9, n_base_ftrs=5, n_states=3, max_dur=3, extract_seg_ftrs=1), {"nodur_impl": 2}, 400, 150, 0.3),
    "frame_on_lattice": (dict(model_type="stdframe", n_labs=7, n_base_ftrs=5), {"frame_impl": 1}, 640, 200, 0.5),
    "frame3state_on_lattice": (dict(model_type="stdframe", n_labs=12, n_base_ftrs=5, n_states=3), {"frame_impl": 1}, 400, 150, 0.5),
}


@pytest.mark.parametrize("kind", sorted(REFILL_KINDS))
def test_slot_refill_matches_oracle(oracle, kind):
    kw, opts, n_utt, t_hi, scale = REFILL_KINDS[kind]
    rng = np.random.default_rng(sum(map(ord, kind)))
    lens = _refill_lengths(rng, n_utt, t_hi)
    states = kw.get("n_states", 1)
    P = kw.get("n_actual_labs", kw["n_labs"] // states) if kw["model_type"] != "stdframe" else kw["n_labs"] // states
    off, ftrs, labs = _labelled_batch(rng, lens, kw["n_base_ftrs"], P, 2 * kw.get("max_dur", 1) + 2, states)
    cfg = make_config(**kw)
    lam = rng.uniform(-scale, scale, oracle.lambda_len(cfg))
    want = oracle.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=8)
    m = gpu(cfg)
    for k, v in opts.items():
        m.set_option(k, v)
    m.set_option("max_clusters", 1)          # 16 slots (tcgen05) / one cluster (FFMA) for the whole batch: every slot refills ~n_utt/16 times
    m.set_lambda(lam)
    got = m.fwdbwd(off, ftrs, labs)
    assert_train_close(got, want, kind)
    # the same batch with all clusters resident (one or two utterances per slot): same per-utterance results
    m.set_option("max_clusters", 0)
    got2 = m.fwdbwd(off, ftrs, labs)
    assert_train_close(got2, want, kind + " (all clusters)")
    m.close()


@pytest.mark.parametrize("dp_impl", [3, 2])
def test_slot_refill_cfg4_geometry_matches_oracle(oracle, dp_impl):
    """cfg4's own geometry (610 labels: clusters of 8 CTAs, E in tensor memory) with every slot refilled: 64 short utterances
    on ONE cluster, lengths 1, 33 and 65 among them."""
    rng = np.random.default_rng(610)
    lens = rng.integers(1, 40, 64)
    lens[[3, 17, 40]] = 1
    lens[[5, 21, 33, 50]] = 33
    lens[[9]] = 65
    off, ftrs, labs = _labelled_batch(rng, lens, 105, 61, 14)
    cfg = make_config("stdseg", n_labs=610, n_base_ftrs=105, max_dur=10, n_actual_labs=61, extract_seg_ftrs=1)
    lam = rng.uniform(-0.01, 0.01, oracle.lambda_len(cfg))
    want = oracle.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=8)
    m = gpu(cfg)
    m.set_option("dp_impl", dp_impl)
    m.set_option("max_clusters", 1)
    m.set_lambda(lam)
    got = m.fwdbwd(off, ftrs, labs)
    assert ("dp_ks_kernel" if dp_impl == 3 else "dp_tc_kernel") in m.plan_info()
    assert_train_close(got, want, "cfg4 geometry, one cluster")
    m.close()


def block_relative_errors(g, gw, sidx):
    """worst |g - gw| per label block of lambda (a label's state weights + its incoming transition weights), relative to the block's
    own largest entry -- tighter than one tolerance scaled by the global maximum"""
    starts = np.sort(np.asarray(sidx, np.int64))
    bounds = np.concatenate([starts, [len(gw)]])
    out = np.zeros(len(starts))
    for i in range(len(starts)):
        a, b = bounds[i], bounds[i + 1]
        out[i] = np.abs(g[a:b] - gw[a:b]).max() / max(np.abs(gw[a:b]).max(), 1e-300)
    return out


@pytest.mark.parametrize("dp_impl", [3, 2])
def test_cfg4_bench_shard_matches_reference_golden(dp_impl):
    """THE bench workload (cfg4 on workloads.timit_train_batch(0, 462), 138 137 frames, the minibatch rank 0 times) against the
    golden the unmodified reference produced for it (tests/golden/make_golden_cfg4_shard0.py): per-utterance numerator and logZ
    to 1e-5 relative, the 891 210-entry gradient to 1e-4 (element-wise against the global maximum, per label block against the
    block's own maximum, and in relative L2 norm), log-likelihood sum to 1e-6 relative."""
    import json
    import os
    import workloads
    from helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "cfg4_shard0_golden.npz"))
    pin = json.load(open(os.path.join(GOLDEN, "cfg4_shard0_pin.json")))
    off, ftrs, labs = workloads.timit_train_batch(0, 462)
    m = crf_b200.CrfGpu(crf_b200.make_config(**workloads.cfg4_kwargs()))
    m.set_option("dp_impl", dp_impl)
    m.set_lambda(workloads.lam_for("cfg4", m.lambda_len))
    g, n, lz = m.fwdbwd(off, ftrs, labs)
    gw = z["grad32"].astype(np.float64)
    assert_train_close((g, n, lz), (gw, z["numer"], z["logZ"]), "cfg4 bench shard 0")
    ll = float((n - lz).sum())
    assert abs(ll - pin["loglik"]) <= 1e-6 * abs(pin["loglik"]), (ll, pin["loglik"])
    rel_l2 = np.linalg.norm(g - gw) / np.linalg.norm(gw)
    sidx, _ = m.index_maps()
    blk = block_relative_errors(g, gw, sidx)
    print(f"cfg4 shard 0: loglik {ll:.6f} (golden {pin['loglik']:.6f}), gradient rel-L2 error {rel_l2:.3e}, worst label block {blk.max():.3e}")
    assert rel_l2 <= 1e-4
    assert blk.max() <= 2e-4
    assert abs(np.sum(g * g) - pin["grad_sq"]) <= 2e-4 * pin["grad_sq"]
    m.close()


# ------------------------------------------------------------------------------------------------------------------
# CRF_LogMath / computeExpF error semantics: the reference throws overflow_error for NaN / Inf / empty log-sums and
# runtime_error when the posterior masses of a frame leave its band; the C ABI reports CRFGPU_ERR_NUMERIC.
def test_posterior_mass_matches_oracle(oracle):
    """sum over labels of the frame posteriors, from the device's own posterior array, against the oracle's alpha / beta: the probability
    that a segment ends on the frame (segmental) and 1 (frame-level)"""
    import ctypes as C
    rng = np.random.default_rng(19)
    off, ftrs, labs = synth_batch(rng, 1, 43, 43, 8, 6, 1, 9)
    cfg = make_config("stdseg", n_labs=6 * 4, n_base_ftrs=8, max_dur=4, n_actual_labs=6, extract_seg_ftrs=1)
    lam = rng.uniform(-0.3, 0.3, oracle.lambda_len(cfg))
    T, L = int(off[-1]), 24
    a = np.zeros((T, L)); b = np.zeros((T, L)); g = np.zeros(len(lam)); nu = np.zeros(1); z = np.zeros(1)
    P = lambda x, t: x.ctypes.data_as(C.POINTER(t))
    rc = oracle.lib.crforacle_fwdbwd_dump(C.byref(cfg), P(lam, C.c_double), C.c_uint32(len(lam)), C.c_uint32(T), P(ftrs, C.c_float),
                                          P(labs, C.c_uint32), P(g, C.c_double), P(nu, C.c_double), P(z, C.c_double), P(a, C.c_double), P(b, C.c_double))
    assert rc == 0
    with np.errstate(over="ignore"):
        want = np.where((a > -1e300) & (b > -1e300), np.exp(np.minimum(a + b - z[0], 0.0)), 0.0).sum(axis=1)
    m = gpu(cfg)
    m.set_lambda(lam)
    m.fwdbwd(off, ftrs, labs)
    mass = m.fetch_posterior_mass()
    np.testing.assert_allclose(mass, want, atol=2e-5)
    assert abs(mass[-1] - 1.0) < 1e-5 and mass.min() >= -1e-6 and mass.max() <= 1 + 1e-5      # every path ends on the last frame
    m.close()
    off, ftrs, labs = synth_batch(rng, 6, 1, 80, 105, 61, 2, 9)
    m = gpu(make_config("stdframe", n_labs=61, n_base_ftrs=105))
    m.set_lambda(rng.uniform(-0.25, 0.25, m.lambda_len))
    m.fwdbwd(off, ftrs, labs)
    np.testing.assert_allclose(m.fetch_posterior_mass(), 1.0, atol=2e-5)
    m.close()


@pytest.mark.parametrize("kind", ["stdseg", "frame", "nodur_native", "frame_lattice"])
def test_nan_features_return_numeric_error_and_handle_survives(oracle, kind):
    """a NaN in one frame poisons that utterance's lattice: the reference throws from CRF_LogMath, the device reports CRFGPU_ERR_NUMERIC
    (non-finite logZ and posterior masses outside the band) -- and the same handle computes the next, clean batch correctly"""
    rng = np.random.default_rng(23)
    F, P, D = 7, 5, 3
    off, ftrs, labs = synth_batch(rng, 5, 4, 30, F, P, 1, 5)
    if kind == "stdseg":
        cfg = make_config("stdseg", n_labs=P * D, n_base_ftrs=F, max_dur=D, n_actual_labs=P, extract_seg_ftrs=1)
    elif kind == "nodur_native":
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=F, max_dur=D, n_actual_labs=P, extract_seg_ftrs=1)
    else:
        cfg = make_config("stdframe", n_labs=P, n_base_ftrs=F)
    lam = rng.uniform(-0.3, 0.3, oracle.lambda_len(cfg))
    m = gpu(cfg)
    if kind == "nodur_native":
        m.set_option("nodur_impl", 1)
    if kind == "frame_lattice":
        m.set_option("frame_impl", 1)
    m.set_lambda(lam)
    bad = ftrs.copy()
    bad[int(off[2]) + 1, 3] = np.nan
    with pytest.raises(crf_b200.CrfGpuError) as ei:
        m.fwdbwd(off, bad, labs)
    assert ei.value.code == 4, str(ei.value)
    got = m.fwdbwd(off, ftrs, labs)
    assert_train_close(got, oracle.fwdbwd(cfg, lam, off, ftrs, labs), kind)
    m.close()


def test_overflowing_lambda_returns_numeric_error():
    """weights that leave the double range in the reference (expE throws overflow_error): CRFGPU_ERR_NUMERIC"""
    rng = np.random.default_rng(29)
    off, ftrs, labs = synth_batch(rng, 3, 5, 20, 6, 4, 1, 4)
    m = gpu(make_config("stdseg", n_labs=8, n_base_ftrs=6, max_dur=2, n_actual_labs=4, extract_seg_ftrs=1))
    lam = np.zeros(m.lambda_len)
    lam[:] = np.inf
    m.set_lambda(lam)
    with pytest.raises(crf_b200.CrfGpuError) as ei:
        m.fwdbwd(off, ftrs, labs)
    assert ei.value.code == 4
    m.close()


def test_plan_info_names_the_kernels():
    c = TRAIN["stdseg_d10_segftr"]
    m = gpu(c["cfg"])
    m.set_lambda(c["lam"])
    m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    info = m.plan_info()
    assert "lattice=dp_ks_kernel" in info and "locksteps=" in info and "decode:" in info
    m.set_option("dp_impl", 2)
    m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    assert "lattice=dp_tc_kernel" in m.plan_info()
    m.set_option("dp_impl", 1)
    m.fwdbwd(c["off"], c["ftrs"], c["labs"])
    assert "FFMA fallback" in m.plan_info()
    m.close()


# ---- window streams with context frames / boundary deltas / a joined second stream (crfgpu_*2) ----------------------------------
def f2_of(c):
    return c["ftrs2"] if c["cfg"].n_base_ftrs2 else None


@pytest.mark.parametrize("name", sorted(k for k in JOINED if k.startswith("win_")))
def test_joined_windows_bit_exact(name):
    """expand_joined_kernel against the reference's own window streams (CRF_InFtrStream_SeqMultiWindow with context frames, joined by
    QN_InFtrStream_JoinFtrs; goldens of make_golden_joined.py)"""
    c = JOINED[name]
    m = gpu(c["cfg"])
    got = m.expand_windows(c["ftrs"], f2_of(c))
    assert got.shape == c["win"].shape
    assert np.array_equal(got.view(np.uint32), c["win"].view(np.uint32))
    m.close()


@pytest.mark.parametrize("name", sorted(k for k in JOINED if k.startswith("train_")))
def test_joined_fwdbwd_matches_reference_golden(name):
    c = JOINED[name]
    m = gpu(c["cfg"])
    assert m.lambda_len == len(c["lam"])
    m.set_lambda(c["lam"])
    got = m.fwdbwd(c["off"], c["ftrs"], c["labs"], ftrs2=f2_of(c))
    assert_train_close(got, (c["grad"], c["numer"], c["logZ"]), name)
    # staged path gives the same
    m.stage(c["off"], c["ftrs"], c["labs"], ftrs2=f2_of(c)); m.fwdbwd_staged()
    assert_train_close(m.fetch_fwdbwd(), (c["grad"], c["numer"], c["logZ"]), name + " staged")
    m.close()


@pytest.mark.parametrize("name", sorted(k for k in JOINED if k.startswith("vit_")))
def test_joined_viterbi_bit_exact(name):
    c = JOINED[name]
    m = gpu(c["cfg"])
    m.set_lambda(c["lam"])
    segs, cost = m.viterbi(c["off"], c["ftrs"], ftrs2=f2_of(c))
    want = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])
    assert [len(s_[0]) for s_ in segs] == [int(k) for k in c["nseg"]]
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), c["cost"].view(np.uint32))
    m.close()


def test_timit_recipe_shape_matches_oracle(oracle):
    """The production TIMIT recipe at its real shape (demo/segmental-timit-demo.cfg.in:11-48): stdseg_no_dur_no_segtransftr + stdtrans,
    48 phones, maxDur 10, stream 1 = 144 inputs -> 1162 segment features (state), stream 2 = the same inputs padded by 6 frames on
    each side -> 13 x 144 = 1872 context features (transition); training and decoding against the oracle."""
    rng = np.random.default_rng(48)
    P, D, F = 48, 10, 144
    off, _, labs = synth_batch(rng, 5, 20, 90, 1, P, 2, 14)
    n = int(off[-1])
    f1 = rng.random((n, F), dtype=np.float32)
    f2 = rng.random((n + 12 * (len(off) - 1), F), dtype=np.float32)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=F, max_dur=D, n_actual_labs=P, extract_seg_ftrs=1,
                      n_base_ftrs2=F, left_ctx2=6, right_ctx2=6, use_trans_ftrs=1, state_fidx=(0, 1161), trans_fidx=(1162, 3033))
    lam = rng.uniform(-0.01, 0.01, oracle.lambda_len(cfg))
    want = oracle.fwdbwd(cfg, lam, off, f1, labs, n_threads=8, ftrs2=f2)
    m = gpu(cfg)
    assert m.lambda_len == len(lam) == 48 * 1163 + 48 * 48 * 1873
    m.set_lambda(lam)
    got = m.fwdbwd(off, f1, labs, ftrs2=f2)
    assert_train_close(got, want, "TIMIT recipe shape")
    segs, cost = m.viterbi(off, f1, ftrs2=f2)
    wsegs, wcost, _ = oracle.viterbi(cfg, lam, off, f1, f2)
    for a, b in zip(segs, wsegs):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert np.array_equal(cost.view(np.uint32), wcost.view(np.uint32))
    m.close()


# ---- decoding against a phone-bigram language model (crfgpu_set_phone_lm) --------------------------------------------------------
@pytest.mark.parametrize("name", sorted(VIT_LM))
def test_viterbi_phone_lm_bit_exact(name):
    """nStateDecode with an input lm_fst for complete phone-bigram LMs (one state per phone): labels, durations, emitted phones and the
    float path cost (final weight included) against the reference's decoder; dropping the LM restores the free-phone result"""
    c = VIT_LM[name]
    m = gpu(c["cfg"])
    m.set_lambda(c["lam"])
    free, free_cost = m.viterbi(c["off"], c["ftrs"])
    m.set_phone_lm(c["lm_start"], c["lm_bigram"], c["lm_final"])
    segs, cost = m.viterbi(c["off"], c["ftrs"])
    want = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])
    assert [len(s_[0]) for s_ in segs] == [int(k) for k in c["nseg"]]
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), c["cost"].view(np.uint32))
    m.set_phone_lm()
    again, again_cost = m.viterbi(c["off"], c["ftrs"])
    assert np.array_equal(again_cost.view(np.uint32), free_cost.view(np.uint32))
    for a, b in zip(again, free):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    m.close()


def test_phone_lm_matches_oracle_at_timit_size(oracle):
    """61 phones, maxDur 10, 105-dim segment features, a smoothed random bigram: the device against the oracle on ragged utterances"""
    rng = np.random.default_rng(61)
    P, D, F = 61, 10, 105
    off, ftrs, _ = synth_batch(rng, 24, 30, 400, F, P, 2, 14)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=F, max_dur=D, extract_seg_ftrs=1)
    lam = rng.uniform(-0.05, 0.05, oracle.lambda_len(cfg))
    prob = rng.dirichlet(np.ones(P) * 0.5, P) * 0.9 + 0.1 / P
    bg = (-np.log(prob)).astype(np.float32); st = (-np.log(rng.dirichlet(np.ones(P)))).astype(np.float32); fin = rng.uniform(0, 1, P).astype(np.float32)
    wsegs, wcost, _ = oracle.viterbi(cfg, lam, off, ftrs, lm=(st, bg, fin))
    m = gpu(cfg)
    m.set_lambda(lam)
    m.set_phone_lm(st, bg, fin)
    segs, cost = m.viterbi(off, ftrs)
    for a, b in zip(segs, wsegs):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert np.array_equal(cost.view(np.uint32), wcost.view(np.uint32))
    m.close()


def test_phone_lm_rejects_what_is_not_implemented():
    import ctypes as C
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=12, n_base_ftrs=4, n_states=3, max_dur=1)
    m = gpu(cfg)
    z = np.zeros(16, np.float32); zp = z.ctypes.data_as(C.POINTER(C.c_float))
    # N states per phone: the free-phone LM returns to the start state through epsilon arcs -- a bigram table has no place in it
    assert m.lib.crfgpu_set_phone_lm(m.h, zp, zp, zp) == 2
    m.close()
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=4, n_base_ftrs=4, max_dur=1)
    m = gpu(cfg)
    assert m.lib.crfgpu_set_phone_unigram_lm(m.h, zp, zp, zp) == 2      # ... and one state per phone has no epsilon arcs
    m.close()
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=5, n_base_ftrs=4, max_dur=2, extract_seg_ftrs=1)
    m = gpu(cfg)
    bg = np.zeros((5, 5), np.float32); bg[1, 3] = np.inf
    with pytest.raises(crf_b200.CrfGpuError) as e:
        m.set_phone_lm(np.zeros(5, np.float32), bg, np.zeros(5, np.float32))
    assert e.value.code == 2          # a missing arc
    m.close()


@pytest.mark.parametrize("name", sorted(VIT_BEAM))
def test_viterbi_beam_bit_exact(name):
    """crfgpu_set_beam: the decoder's beam pruning (pruning(), expansion threshold) with the free-phone LM and with a bigram LM against the
    reference's nStateDecode with the same beam -- labels, durations, emitted phones, float cost; beam 0 afterwards is the unpruned result"""
    c = VIT_BEAM[name]
    m = gpu(c["cfg"])
    m.set_lambda(c["lam"])
    if len(c["lm_start"]):
        m.set_phone_lm(c["lm_start"], c["lm_bigram"], c["lm_final"])
    full, full_cost = m.viterbi(c["off"], c["ftrs"])
    m.set_beam(float(c["beam"][0]))
    segs, cost = m.viterbi(c["off"], c["ftrs"])
    want = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])
    assert [len(s_[0]) for s_ in segs] == [int(k) for k in c["nseg"]]
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), c["cost"].view(np.uint32))
    assert np.all(cost >= full_cost)                     # pruning can only lose paths
    m.set_beam(0.0)
    again, again_cost = m.viterbi(c["off"], c["ftrs"])
    assert np.array_equal(again_cost.view(np.uint32), full_cost.view(np.uint32))
    m.close()


def test_beam_matches_oracle_at_timit_size(oracle):
    rng = np.random.default_rng(62)
    P, D, F = 61, 10, 105
    off, ftrs, _ = synth_batch(rng, 16, 30, 300, F, P, 2, 14)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=F, max_dur=D, extract_seg_ftrs=1)
    lam = rng.uniform(-0.05, 0.05, oracle.lambda_len(cfg))
    m = gpu(cfg)
    m.set_lambda(lam)
    for beam in (0.3, 1.0, 4.0):
        wsegs, wcost, _ = oracle.viterbi(cfg, lam, off, ftrs, beam=beam)
        m.set_beam(beam)
        segs, cost = m.viterbi(off, ftrs)
        for a, b in zip(segs, wsegs):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        assert np.array_equal(cost.view(np.uint32), wcost.view(np.uint32))
    m.close()
