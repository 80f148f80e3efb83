"""CPU tests: the C restatement (oracle/crf_oracle.c) against (a) the committed golden vectors that
were produced by the unmodified reference (tests/golden/make_golden.py), (b) the SURVEY.md 8(c)
known answers for the bundled toy set and (c), when oracle/_ref is built, the reference itself
on fresh seeded inputs."""
import numpy as np
import pytest

from oracle.binding import make_config
from helpers import load_cases, split_segs, synth_batch

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


@pytest.mark.parametrize("name", sorted(VIT_LM))
def test_viterbi_lm_golden_bit_exact(oracle, name):
    """decoding against a phone-bigram LM: the oracle against the reference's nStateDecode driven with an input lm_fst
    (goldens of make_golden_lm.py: random / quantised costs, non-final phone states, transition features)"""
    c = VIT_LM[name]
    segs, cost, _ = oracle.viterbi(c["cfg"], c["lam"], c["off"], c["ftrs"], lm=(c["lm_start"], c["lm_bigram"], c["lm_final"]))
    want = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])
    assert [len(s_[0]) for s_ in segs] == [int(k) for k in c["nseg"]]
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), c["cost"].view(np.uint32))


def f2_of(c):
    return c["ftrs2"] if c["cfg"].n_base_ftrs2 else None


@pytest.mark.parametrize("name", sorted(k for k in JOINED if k.startswith("win_")))
def test_joined_window_golden_bit_exact(oracle, name):
    """context frames / boundary deltas / joined second stream: window vectors against the reference's own streams (make_golden_joined.py)"""
    c = JOINED[name]
    got = np.nan_to_num(oracle.window_ftrs(c["cfg"], c["ftrs"], f2_of(c)), nan=0.0)
    assert np.array_equal(got.view(np.uint32), c["win"].view(np.uint32))


@pytest.mark.parametrize("name", sorted(k for k in JOINED if k.startswith("train_")))
def test_joined_train_golden(oracle, name):
    """training over joined / context window streams (incl. the TIMIT recipe's layout: state features from stream 1's segment
    features, transition features from stream 2's context frames) against goldens produced by the reference"""
    c = JOINED[name]
    assert oracle.lambda_len(c["cfg"]) == len(c["lam"])
    grad, numer, logz = oracle.fwdbwd(c["cfg"], c["lam"], c["off"], c["ftrs"], c["labs"], ftrs2=f2_of(c))
    np.testing.assert_allclose(logz, c["logZ"], rtol=1e-13)
    np.testing.assert_allclose(numer, c["numer"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(grad, c["grad"], rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize("name", sorted(k for k in JOINED if k.startswith("vit_")))
def test_joined_viterbi_golden_bit_exact(oracle, name):
    c = JOINED[name]
    segs, cost, _ = oracle.viterbi(c["cfg"], c["lam"], c["off"], c["ftrs"], f2_of(c))
    want = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), c["cost"].view(np.uint32))


def test_toy_known_answers(oracle):
    """logZ, numerators, grad[0..3] and |grad|^2 on CRFTrain/test.ascii."""
    c = TRAIN["toy_stdframe"]
    grad, numer, logz = oracle.fwdbwd(c["cfg"], c["lam"], c["off"], c["ftrs"], c["labs"])
    np.testing.assert_allclose(logz, [5.488352272567, 4.152084629056, 5.488352272567], rtol=0, atol=1e-11)
    np.testing.assert_allclose(numer, [-0.09, -0.01, -0.09], rtol=0, atol=1e-14)
    np.testing.assert_allclose(grad[:4], [1.677756141, -1.589524031, -1.304487168, 0.088232110], atol=1e-9)
    assert abs(np.sum(grad ** 2) - 41.991090456415) < 1e-10


@pytest.mark.parametrize("name", sorted(TRAIN))
def test_train_golden(oracle, name):
    c = TRAIN[name]
    assert oracle.lambda_len(c["cfg"]) == len(c["lam"])
    grad, numer, logz = oracle.fwdbwd(c["cfg"], c["lam"], c["off"], c["ftrs"], c["labs"])
    np.testing.assert_allclose(logz, c["logZ"], rtol=1e-13)
    np.testing.assert_allclose(numer, c["numer"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(grad, c["grad"], rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize("tied", [False, True])
@pytest.mark.parametrize("name", sorted(NODUR))
def test_train_nodur_golden(oracle, name, tied):
    """stdseg_no_dur / _no_transftr / _no_segtransftr training: the native O(P^2 + D*P) restatement (fb_nodur) and the tied
    (duration, phone) restatement, both against goldens produced by the reference's own no_dur node classes and grad builders."""
    c = NODUR[name]
    assert oracle.lambda_len(c["cfg"]) == len(c["lam"])
    grad, numer, logz = oracle.fwdbwd(c["cfg"], c["lam"], c["off"], c["ftrs"], c["labs"], tied=tied)
    np.testing.assert_allclose(logz, c["logZ"], rtol=1e-12)
    np.testing.assert_allclose(numer, c["numer"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(grad, c["grad"], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("name", sorted(NSTATE))
def test_train_nodur_nstate_golden(oracle, name):
    """N states per phone in the segmental no_dur models (CRF_StdSegNStateNode_WithoutDurLab_WithoutSegTransFtr): every sub-state is a
    segment of its own; goldens from the reference (make_golden.py nstate), one of them with reference paths that skip sub-states."""
    c = NSTATE[name]
    assert oracle.lambda_len(c["cfg"]) == len(c["lam"])
    grad, numer, logz = oracle.fwdbwd(c["cfg"], c["lam"], c["off"], c["ftrs"], c["labs"])
    np.testing.assert_allclose(logz, c["logZ"], rtol=1e-13)
    np.testing.assert_allclose(numer, c["numer"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(grad, c["grad"], rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize("name", sorted(TRANSFTR))
def test_train_transftr_golden(oracle, name):
    """frame-level CRFs with transition FEATURES (stdtrans) against goldens produced by the reference (make_golden_transftr.py)."""
    c = TRANSFTR[name]
    assert oracle.lambda_len(c["cfg"]) == len(c["lam"])
    grad, numer, logz = oracle.fwdbwd(c["cfg"], c["lam"], c["off"], c["ftrs"], c["labs"])
    np.testing.assert_allclose(logz, c["logZ"], rtol=1e-13)
    np.testing.assert_allclose(numer, c["numer"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(grad, c["grad"], rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize("name", sorted(TRANSFTR_NS))
def test_train_transftr_nstate_golden(oracle, name):
    """transition FEATURES with N states per label, frame-level (CRF_StdNStateNode) and segmental without duration labels
    (CRF_StdSegNStateNode_WithoutDurLab_WithoutSegTransFtr), against goldens produced by the reference (make_golden_transftr_nstate.py)."""
    c = TRANSFTR_NS[name]
    assert oracle.lambda_len(c["cfg"]) == len(c["lam"])
    grad, numer, logz = oracle.fwdbwd(c["cfg"], c["lam"], c["off"], c["ftrs"], c["labs"])
    np.testing.assert_allclose(logz, c["logZ"], rtol=1e-13)
    np.testing.assert_allclose(numer, c["numer"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(grad, c["grad"], rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize("name", sorted(VIT_TF))
def test_viterbi_transftr_golden_bit_exact(oracle, name):
    """decoding with transition FEATURES against the reference decoder (goldens of make_golden_transftr.py)"""
    c = VIT_TF[name]
    segs, cost, _ = oracle.viterbi(c["cfg"], c["lam"], c["off"], c["ftrs"])
    want = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), c["cost"].view(np.uint32))


def test_train_threads_match_single(oracle):
    c = TRAIN["frame_1state"]
    g1, n1, z1 = oracle.fwdbwd(c["cfg"], c["lam"], c["off"], c["ftrs"], c["labs"], n_threads=1)
    g3, n3, z3 = oracle.fwdbwd(c["cfg"], c["lam"], c["off"], c["ftrs"], c["labs"], n_threads=3)
    np.testing.assert_allclose(g3, g1, rtol=1e-12, atol=1e-12)
    assert np.array_equal(n1, n3) and np.array_equal(z1, z3)


@pytest.mark.parametrize("name", sorted(VIT))
def test_viterbi_golden_bit_exact(oracle, name):
    c = VIT[name]
    segs, cost, _ = oracle.viterbi(c["cfg"], c["lam"], c["off"], c["ftrs"])
    want = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])
    assert [len(s[0]) for s in segs] == [int(k) for k in c["nseg"]]
    for got, exp in zip(segs, want):
        assert np.array_equal(got[0], exp[0]) and np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])
    assert np.array_equal(cost.view(np.uint32), c["cost"].view(np.uint32))


def test_viterbi_survey_tie_goldens(oracle):
    """SURVEY.md 9.5 all-ties table (lambda = 0, two constant features)."""
    def run(P, N, D, T):
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P * N, n_base_ftrs=2, n_states=N, max_dur=D)
        lam = np.zeros(oracle.lambda_len(cfg))
        segs, _, _ = oracle.viterbi(cfg, lam, [0, T], np.ones((T, 2), np.float32))
        return [int(v) for v in segs[0][0]], [int(np.int32(v)) for v in segs[0][2]]
    assert run(5, 1, 1, 6) == ([0, 1, 0, 1, 0, 1], [0, 1, 0, 1, 0, 1])
    assert run(5, 1, 1, 7) == ([0, 1, 0, 1, 0, 1, 0], [0, 1, 0, 1, 0, 1, 0])
    assert run(5, 3, 1, 7) == ([0, 0, 1, 2, 0, 1, 2], [0, -1, -1, -1, 0, -1, -1])
    assert run(4, 1, 3, 7)[0] == [0, 1, 0]
    assert run(4, 1, 3, 8)[0] == [0, 1, 0]


@pytest.mark.parametrize("name", sorted(k for k in WIN if k.startswith("win_")))
def test_window_features_bit_exact(oracle, name):
    c = WIN[name]
    got = oracle.window_ftrs(c["cfg"], c["x"])
    exp = c["out"]
    assert np.array_equal(np.isnan(got), np.isnan(exp))
    assert np.array_equal(got[~np.isnan(got)].view(np.uint32), exp[~np.isnan(exp)].view(np.uint32))


@pytest.mark.parametrize("name", sorted(k for k in WIN if k.startswith("lab_")))
def test_label_grouping(oracle, name):
    c = WIN[name]
    assert np.array_equal(oracle.window_labs(c["cfg"], c["labs"]), c["out"])


def test_lambda_len_formulas(oracle):
    """cfg2 10 187, cfg3 23 424, cfg4 891 210."""
    assert oracle.lambda_len(make_config("stdframe", n_labs=61, n_base_ftrs=105)) == 10187
    assert oracle.lambda_len(make_config("stdframe", n_labs=183, n_base_ftrs=105, n_states=3)) == 23424
    assert oracle.lambda_len(make_config("stdseg", n_labs=610, n_base_ftrs=105, max_dur=10, n_actual_labs=61,
                                         extract_seg_ftrs=1)) == 891210


# ---- live comparison with the reference (only where oracle/_ref exists) ---------------------
@pytest.mark.parametrize("kind", ["frame1", "frame3", "stdseg", "frame1_transftr"])
def test_fwdbwd_vs_reference_live(oracle, reflib, kind):
    rng = np.random.default_rng({"frame1": 11, "frame3": 12, "stdseg": 13, "frame1_transftr": 14}[kind])
    if kind == "frame3":
        off, ftrs, labs = synth_batch(rng, 4, 3, 30, 6, 4, states=3)
        cfg = make_config("stdframe", n_labs=12, n_base_ftrs=6, n_states=3)
    else:
        off, ftrs, labs = synth_batch(rng, 4, 1, 30, 6, 5)
        if kind == "frame1":
            cfg = make_config("stdframe", n_labs=5, n_base_ftrs=6)
        elif kind == "frame1_transftr":
            cfg = make_config("stdframe", n_labs=5, n_base_ftrs=6, use_trans_ftrs=1, trans_fidx=(1, 3))
        else:
            cfg = make_config("stdseg", n_labs=15, n_base_ftrs=6, max_dur=3, n_actual_labs=5, extract_seg_ftrs=1)
    n = reflib.lambda_len(cfg)
    assert n == oracle.lambda_len(cfg)
    lam = rng.uniform(-0.1, 0.1, n)
    g1, n1, z1 = reflib.fwdbwd(cfg, lam, off, ftrs, labs)
    g2, n2, z2 = oracle.fwdbwd(cfg, lam, off, ftrs, labs)
    np.testing.assert_allclose(z2, z1, rtol=1e-13)
    np.testing.assert_allclose(n2, n1, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(g2, g1, rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize("P,N,D,segf", [(6, 1, 1, 0), (5, 3, 1, 0), (5, 1, 4, 1), (3, 2, 3, 1)])
def test_viterbi_vs_reference_live(oracle, reflib, P, N, D, segf):
    rng = np.random.default_rng(100 + P * 7 + N * 3 + D)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P * N, n_base_ftrs=3, n_states=N, max_dur=D,
                      extract_seg_ftrs=segf)
    n = reflib.lambda_len(cfg)
    off = np.concatenate([[0], np.cumsum([1, 2, 4, 11, 35])]).astype(np.uint32)
    for lam in (rng.uniform(-0.5, 0.5, n), np.zeros(n), np.round(rng.uniform(-1, 1, n))):
        ftrs = (np.round(rng.random((int(off[-1]), 3)) * 4) / 4).astype(np.float32)
        s1, c1, _ = reflib.viterbi(cfg, lam, off, ftrs)
        s2, c2, _ = oracle.viterbi(cfg, lam, off, ftrs)
        for a, b in zip(s1, s2):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        assert np.array_equal(c1.view(np.uint32), c2.view(np.uint32))


@pytest.mark.parametrize("name", sorted(VIT_BEAM))
def test_viterbi_beam_golden_bit_exact(oracle, name):
    """beam pruning (one state per phone, free-phone and bigram LM): the oracle against the reference's nStateDecode with input_beam > 0"""
    c = VIT_BEAM[name]
    lm = (c["lm_start"], c["lm_bigram"], c["lm_final"]) if len(c["lm_start"]) else None
    segs, cost, _ = oracle.viterbi(c["cfg"], c["lam"], c["off"], c["ftrs"], lm=lm, beam=float(c["beam"][0]))
    want = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])
    assert [len(s_[0]) for s_ in segs] == [int(k) for k in c["nseg"]]
    for got, exp in zip(segs, want):
        assert all(np.array_equal(x, y) for x, y in zip(got, exp))
    assert np.array_equal(cost.view(np.uint32), c["cost"].view(np.uint32))
