"""The C++ host layer (asr-craft_b200/host/crf_host.h: CRF_Model / CRF_FeatureMap_config / CRF_GradBuilder /
CRF_Minibatch_GradAccumulator / CRF_ViterbiDecoder_StdSeg_NoSegTransFtr over the C ABI), driven by host_selftest the way
CRFTrain / CRFDecode drive the reference classes, against goldens produced by the unmodified reference."""
import os
import subprocess

import numpy as np
import pytest

from helpers import load_cases, split_segs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "asr-craft_b200", "host", "host_selftest")
TRAIN = load_cases("train_golden.npz")
NODUR = load_cases("train_nodur_golden.npz")
VIT = load_cases("viterbi_golden.npz")
JOINED = load_cases("joined_golden.npz")


def write_case(path, c, decode_arcs=None):
    cfg = c["cfg"]
    n_utt, N = len(c["off"]) - 1, int(c["off"][-1])
    grad = c.get("grad", np.zeros(len(c["lam"])))
    with open(path, "w") as f:
        arcs = decode_arcs if decode_arcs is not None else np.zeros((0, 3), np.int64)
        f.write(f"{0 if decode_arcs is None else 1} {cfg.model_type} {cfg.n_labs} {cfg.n_base_ftrs} {cfg.n_states} {cfg.max_dur} {cfg.n_actual_labs} "
                f"{cfg.extract_seg_ftrs} {n_utt} {N} {len(c['lam'])} {len(arcs)}\n")
        for a in (c["off"], c["lam"], c["ftrs"].ravel(), c.get("labs", np.zeros(N)), c.get("logZ", np.zeros(n_utt))[:n_utt],
                  c.get("numer", np.zeros(n_utt)), grad, np.asarray(arcs).ravel()):
            f.write(" ".join(repr(float(x)) for x in np.asarray(a, np.float64)) + "\n")
        if cfg.n_base_ftrs2:      # trailer: the joined second stream and the feature ranges of the map
            f2 = c["ftrs2"]
            f.write(f"{cfg.n_base_ftrs2} {cfg.extract_seg_ftrs2} {cfg.left_ctx2} {cfg.right_ctx2} {cfg.boundary_delta2} {cfg.use_trans_ftrs} "
                    f"{cfg.state_fidx_start} {cfg.state_fidx_end} {cfg.trans_fidx_start} {cfg.trans_fidx_end} {f2.shape[0]}\n")
            f.write(" ".join(repr(float(x)) for x in f2.ravel()) + "\n")


def test_host_layer_fails_loudly_without_device():
    assert os.path.exists(EXE), "run `make -C asr-craft_b200` (or __graft_entry__.build()) first"
    r = subprocess.run([EXE, "--no-device"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no usable CUDA device" in r.stdout or "CUDA device is present" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["toy_stdframe", "frame_3state", "stdseg_d4_segftr", "nodur:stdseg_no_dur_no_segtransftr_d4_s1"])
def test_host_layer_training_matches_reference_golden(tmp_path, name):
    c = NODUR[name[6:]] if name.startswith("nodur:") else TRAIN[name]
    path = str(tmp_path / "case.txt")
    write_case(path, c)
    r = subprocess.run([EXE, path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host_selftest ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["toy", "rand_P7N1D3s1", "rand_P4N3D2s1", "ties_P5N3D1s0"])
def test_host_layer_decoding_matches_reference_golden(tmp_path, name):
    c = VIT[name]
    lab, dur, phn = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])[0]
    arcs = np.stack([lab.astype(np.int64) + 1, np.where(phn == 0xffffffff, 0, phn.astype(np.int64) + 1), dur.astype(np.int64)], axis=1)
    path = str(tmp_path / "case.txt")
    write_case(path, c, arcs)
    r = subprocess.run([EXE, path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host_selftest ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["train_nodur_joined_recipe", "train_nodur_joined_bdelta", "train_frame_joined_transftr", "vit_rand_P7N1D3c2", "vit_quant_P4N3D2c1"])
def test_host_layer_joined_streams_match_reference_golden(tmp_path, name):
    """the C++ host layer over a JOINED stream (CRF_FeatureStream::join; CRF_Model::setSecondStream = CRFTrain's ftr2_* options): the
    minibatch seam over two views, the per-utterance seam and the decode seam against goldens produced by the reference with two
    window streams (the TIMIT recipe's layout: transition features from the context frames of the padded second stream)"""
    c = JOINED[name]
    arcs = None
    if name.startswith("vit_"):
        lab, dur, phn = split_segs(c["lab"], c["dur"], c["phn"], c["nseg"])[0]
        arcs = np.stack([lab.astype(np.int64) + 1, np.where(phn == 0xffffffff, 0, phn.astype(np.int64) + 1), dur.astype(np.int64)], axis=1)
    path = str(tmp_path / "case.txt")
    write_case(path, c, arcs)
    r = subprocess.run([EXE, path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host_selftest ok (joined streams)" in r.stdout, r.stdout + r.stderr


# ---------------------------------------------------------------------------------------------------------------------
# CRF_Minibatch_GradAccumulator with several streams (CRF_Minibatch_GradAccumulator.cpp:201-322 over the contiguous views of
# CRF_FeatureStreamManager.cpp:425-464), restated here as the checker
def reference_minibatches(n_utt, n_streams, minibatch):
    """[(list of (stream, first utterance, count)), nStreams_active] per accumulateGradient call of one epoch"""
    per = n_utt // n_streams
    first = [s * per for s in range(n_streams)]
    count = [per] * (n_streams - 1) + [n_utt - (n_streams - 1) * per]
    pos = [0] * n_streams
    calls = []
    while True:
        active = [s for s in range(n_streams) if pos[s] < count[s]]       # strmsSegids[stream] != QN_SEGID_BAD at the start of the call
        if not active:
            return calls
        take = []
        for s in active:
            share = minibatch // n_streams + (1 if s < minibatch % n_streams else 0)
            k = min(max(share, 1), count[s] - pos[s])                      # do { ... } while: at least one, stops at the end of the view
            take.append((s, first[s] + pos[s], k))
            pos[s] += k
        calls.append((take, len(active)))


@pytest.mark.parametrize("n_utt,n_streams,minibatch", [(11, 3, 4), (3696, 8, 512), (10, 4, 4), (7, 1, 3), (9, 2, 9), (64, 8, 64), (13, 5, 7)])
def test_minibatch_composition_matches_reference_rule(n_utt, n_streams, minibatch):
    r = subprocess.run([EXE, "--plan", str(n_utt), str(n_streams), str(minibatch)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    got = []
    for line in r.stdout.strip().splitlines():
        body, act = line.rsplit("/", 1)
        take = []
        for tok in body.split():
            s, rest = tok.split(":")
            f, k = rest.split("+")
            take.append((int(s), int(f), int(k)))
        got.append((take, int(act)))
    assert got == reference_minibatches(n_utt, n_streams, minibatch)
    assert sum(k for take, _ in got for _, _, k in take) == n_utt


def test_minibatch_smaller_than_streams_is_an_error():
    r = subprocess.run([EXE, "--plan", "10", "4", "3"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "less than the number of threads" in r.stderr


def _accumulator_case(tmp_path, oracle, n_utt=11):
    from helpers import synth_batch
    from oracle.binding import make_config
    rng = np.random.default_rng(41)
    off, ftrs, labs = synth_batch(rng, n_utt, 3, 40, 6, 5, 1, 7)
    cfg = make_config("stdseg", n_labs=15, n_base_ftrs=6, max_dur=3, n_actual_labs=5, extract_seg_ftrs=1)
    lam = rng.uniform(-0.2, 0.2, oracle.lambda_len(cfg))
    per_utt = []
    for u in range(n_utt):
        so = np.array([0, off[u + 1] - off[u]], np.uint32)
        per_utt.append(oracle.fwdbwd(cfg, lam, so, ftrs[off[u]:off[u + 1]], labs[off[u]:off[u + 1]]))
    path = str(tmp_path / "case.txt")
    write_case(path, dict(cfg=cfg, off=off, lam=lam, ftrs=ftrs, labs=labs))
    return path, per_utt, len(lam)


def _check_accumulator_run(path, per_utt, n_lam, n_streams, minibatch, n_devices):
    r = subprocess.run([EXE, path, "acc", str(n_streams), str(minibatch), str(n_devices)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host_selftest ok" in r.stdout, r.stdout + r.stderr
    rec = np.fromfile(path + ".acc.bin", np.float64).reshape(-1, 4 + n_lam)
    want = reference_minibatches(len(per_utt), n_streams, minibatch)
    assert len(rec) == len(want)
    for k, (take, n_active) in enumerate(want):
        utts = [u for _, f, c in take for u in range(f, f + c)]
        g = sum(per_utt[u][0] for u in utts) / n_active               # gradient averaged over the ACTIVE streams (.cpp:306-308)
        numer = sum(float(per_utt[u][1][0]) for u in utts)
        zx = sum(float(per_utt[u][2][0]) for u in utts)
        assert rec[k, 0] == len(utts) and rec[k, 1] == (1.0 if k == len(want) - 1 else 0.0)
        np.testing.assert_allclose(rec[k, 2], numer, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(rec[k, 3], zx, rtol=1e-5)
        tol = 1e-4 * np.abs(g).max() + 1e-4 * np.abs(g)
        assert np.all(np.abs(rec[k, 4:] - g) <= tol), f"minibatch {k}"


@pytest.mark.gpu
@pytest.mark.parametrize("n_streams,minibatch", [(3, 4), (1, 4), (4, 4), (2, 11)])
def test_accumulator_streams_match_reference_rule(tmp_path, oracle, n_streams, minibatch):
    """an uneven corpus (11 utterances: views of 3, 3, 5 for three streams): per-stream shares, streams running dry one after the other,
    gradient / nStreams_active, end-of-iteration flag -- every minibatch of the epoch against per-utterance oracle results"""
    path, per_utt, n_lam = _accumulator_case(tmp_path, oracle)
    _check_accumulator_run(path, per_utt, n_lam, n_streams, minibatch, 1)


@pytest.mark.gpu
def test_accumulator_two_devices_allreduce_matches_single_device(tmp_path, oracle):
    """the same epoch with the streams spread over TWO GPUs of one process: per-device batches + ONE NCCL all-reduce of
    [gradient | sum numer, sum logZ, n_utt] (crfgpu_comm_init_all / crfgpu_allreduce_grad) == the single-process sums"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    path, per_utt, n_lam = _accumulator_case(tmp_path, oracle)
    _check_accumulator_run(path, per_utt, n_lam, 4, 6, 2)
    _check_accumulator_run(path, per_utt, n_lam, 3, 4, 2)      # device 1 runs dry before device 0: it joins the all-reduce with an empty batch


@pytest.mark.gpu
def test_trainer_resume_matches_uninterrupted_run(tmp_path):
    """CRF_SGTrainer with AdaGrad: two iterations in one run == one iteration + a resumed second one (init_iter, presentations, lambdaAcc and
    gradSqrAcc carried over), and the .iK.gradSqrAcc.out / .done.train files of the reference are written"""
    c = TRAIN["stdseg_d4_segftr"]
    path = str(tmp_path / "case.txt")
    write_case(path, c)
    r = subprocess.run([EXE, path, "resume"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host_selftest ok" in r.stdout, r.stdout + r.stderr
