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
