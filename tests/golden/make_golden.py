"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libcrfref.so, built by
`make -C oracle ref` from /root/reference).  Run here (the container that has /root/reference);
the fixtures are committed so that the GPU box, which has no /root/reference, can still pin
both the C oracle and the CUDA path against the reference's own outputs.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.binding import RefLib, make_config  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CFG_KEYS = ["model_type", "n_labs", "n_base_ftrs", "n_states", "max_dur", "n_actual_labs", "extract_seg_ftrs",
            "use_state_ftrs", "state_fidx_start", "state_fidx_end", "use_trans_ftrs", "trans_fidx_start",
            "trans_fidx_end", "use_state_bias", "use_trans_bias", "state_bias_val", "trans_bias_val"]


def cfg_to_array(cfg):
    return np.array([float(getattr(cfg, k)) for k in CFG_KEYS], np.float64)


def synth(rng, n_utt, t_lo, t_hi, F, P, seg_lo=1, seg_hi=8, states=1):
    lens = rng.integers(t_lo, t_hi + 1, n_utt)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    ftrs = rng.random((int(off[-1]), F), dtype=np.float32)
    labs = np.zeros(int(off[-1]), np.uint32)
    for u in range(n_utt):
        t, prev = int(off[u]), -1
        while t < off[u + 1]:
            d = int(rng.integers(seg_lo, seg_hi + 1))
            lab = int(rng.integers(0, P))
            while lab == prev:
                lab = int(rng.integers(0, P))
            e = min(t + d, int(off[u + 1]))
            if states == 1:
                labs[t:e] = lab
            else:  # walk through the sub-states left to right
                n = e - t
                sub = np.minimum(np.arange(n) * states // max(n, 1), states - 1)
                labs[t:e] = lab * states + sub
            prev, t = lab, e
    return off, ftrs, labs


def toy():
    """The bundled CRFTrain/test.ascii + test.lab.ascii (SURVEY.md 8c known answers)."""
    ref_dir = "/root/reference/CRFTrain"
    f = [line.split() for line in open(os.path.join(ref_dir, "test.ascii"))]
    lab = [line.split() for line in open(os.path.join(ref_dir, "test.lab.ascii"))]
    utts = sorted(set(int(r[0]) for r in f))
    off = [0]
    for u in utts:
        off.append(off[-1] + sum(1 for r in f if int(r[0]) == u))
    ftrs = np.array([[float(x) for x in r[2:]] for r in f], np.float32)
    labs = np.array([int(r[2]) for r in lab], np.uint32)
    return np.array(off, np.uint32), ftrs, labs


def main():
    ref = RefLib()
    rng = np.random.default_rng(20260101)

    # ---- training goldens -------------------------------------------------------------------
    train = {}
    off, ftrs, labs = toy()
    cfg = make_config("stdframe", n_labs=4, n_base_ftrs=3)
    lam = np.array([0.01 * ((7 * i) % 11) - 0.05 for i in range(ref.lambda_len(cfg))])
    train["toy_stdframe"] = (cfg, lam, off, ftrs, labs)
    cfg = make_config("stdseg", n_labs=8, n_base_ftrs=3, max_dur=2, n_actual_labs=4, extract_seg_ftrs=1)
    lam = np.array([0.01 * ((7 * i) % 11) - 0.05 for i in range(ref.lambda_len(cfg))])
    train["toy_stdseg_d2"] = (cfg, lam, off, ftrs, labs)

    off, ftrs, labs = synth(rng, 6, 3, 40, 9, 7)
    cfg = make_config("stdframe", n_labs=7, n_base_ftrs=9)
    train["frame_1state"] = (cfg, rng.uniform(-0.25, 0.25, ref.lambda_len(cfg)), off, ftrs, labs)
    off3, ftrs3, labs3 = synth(rng, 5, 4, 40, 9, 5, states=3)
    cfg = make_config("stdframe", n_labs=15, n_base_ftrs=9, n_states=3)
    train["frame_3state"] = (cfg, rng.uniform(-0.25, 0.25, ref.lambda_len(cfg)), off3, ftrs3, labs3)
    cfg = make_config("stdseg", n_labs=7 * 4, n_base_ftrs=9, max_dur=4, n_actual_labs=7, extract_seg_ftrs=1)
    train["stdseg_d4_segftr"] = (cfg, rng.uniform(-0.05, 0.05, ref.lambda_len(cfg)), off, ftrs, labs)
    cfg = make_config("stdseg", n_labs=7 * 3, n_base_ftrs=9, max_dur=3, n_actual_labs=7, extract_seg_ftrs=0)
    train["stdseg_d3_frameftr"] = (cfg, rng.uniform(-0.05, 0.05, ref.lambda_len(cfg)), off, ftrs, labs)
    offl, ftrsl, labsl = synth(rng, 3, 30, 60, 12, 11, seg_lo=2, seg_hi=14)
    cfg = make_config("stdseg", n_labs=11 * 10, n_base_ftrs=12, max_dur=10, n_actual_labs=11, extract_seg_ftrs=1)
    train["stdseg_d10_segftr"] = (cfg, rng.uniform(-0.02, 0.02, ref.lambda_len(cfg)), offl, ftrsl, labsl)

    out = {}
    for name, (cfg, lam, off, ftrs, labs) in train.items():
        grad, numer, logz = ref.fwdbwd(cfg, lam, off, ftrs, labs)
        out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/lam": lam, f"{name}/off": off, f"{name}/ftrs": ftrs,
                    f"{name}/labs": labs, f"{name}/grad": grad, f"{name}/numer": numer, f"{name}/logZ": logz})
        print(f"train {name}: lambda {len(lam)}, logZ {logz[:3]}, |grad|^2 {np.sum(grad ** 2):.12f}")
    np.savez_compressed(os.path.join(OUT, "train_golden.npz"), **out)

    # ---- Viterbi goldens --------------------------------------------------------------------
    out = {}
    cases = []
    toff, tftrs, _ = toy()
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=4, n_base_ftrs=3)
    lam = np.array([0.01 * ((7 * i) % 11) - 0.05 for i in range(ref.lambda_len(cfg))])
    cases.append(("toy", cfg, lam, toff, tftrs))
    lens = [1, 2, 3, 5, 9, 30, 47]
    voff = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    for (P, N, D, segf) in [(5, 1, 1, 0), (5, 3, 1, 0), (7, 1, 3, 1), (4, 3, 2, 1), (4, 2, 4, 1), (6, 3, 5, 0), (3, 1, 6, 1)]:
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P * N, n_base_ftrs=4, n_states=N, max_dur=D,
                          extract_seg_ftrs=segf)
        n = ref.lambda_len(cfg)
        vf = rng.random((int(voff[-1]), 4), dtype=np.float32)
        cases.append((f"rand_P{P}N{N}D{D}s{segf}", cfg, rng.uniform(-0.5, 0.5, n), voff, vf))
        cases.append((f"ties_P{P}N{N}D{D}s{segf}", cfg, np.zeros(n), voff, vf))
        qf = (np.round(vf * 2) / 2).astype(np.float32)
        cases.append((f"quant_P{P}N{N}D{D}s{segf}", cfg, np.round(rng.uniform(-1, 1, n) * 2) / 2, voff, qf))
    # 61 phones x 3 states, the cfg3 shape at reduced length
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=183, n_base_ftrs=20, n_states=3)
    o183 = np.array([0, 150, 242], np.uint32)
    cases.append(("cfg3_small", cfg, rng.uniform(-0.25, 0.25, ref.lambda_len(cfg)), o183,
                  rng.random((242, 20), dtype=np.float32)))
    names = []
    for name, cfg, lam, off, ftrs in cases:
        segs, cost, logz = ref.viterbi(cfg, lam, off, ftrs)
        nseg = np.array([len(s[0]) for s in segs], np.uint32)
        out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/lam": lam, f"{name}/off": off, f"{name}/ftrs": ftrs,
                    f"{name}/nseg": nseg, f"{name}/cost": cost, f"{name}/logZ": logz,
                    f"{name}/lab": np.concatenate([s[0] for s in segs]) if len(segs) else np.zeros(0, np.uint32),
                    f"{name}/dur": np.concatenate([s[1] for s in segs]),
                    f"{name}/phn": np.concatenate([s[2] for s in segs])})
        names.append(name)
    print("viterbi cases:", len(names))
    np.savez_compressed(os.path.join(OUT, "viterbi_golden.npz"), **out)

    # ---- window-stream goldens --------------------------------------------------------------
    out = {}
    x = rng.random((23, 5), dtype=np.float32)
    for D in (2, 3, 10, 30):
        cfg = make_config("stdseg", n_labs=3 * D, n_base_ftrs=5, max_dur=D, n_actual_labs=3, extract_seg_ftrs=1)
        out[f"win_D{D}/cfg"] = cfg_to_array(cfg)
        out[f"win_D{D}/x"] = x
        out[f"win_D{D}/out"] = ref.window_ftrs(cfg, x)
    lab = np.array([0] * 3 + [2] * 11 + [1] * 1 + [0] * 25 + [3] * 7, np.uint32)
    for D in (1, 2, 3, 10):
        cfg = make_config("stdseg", n_labs=4 * D, n_base_ftrs=5, max_dur=D, n_actual_labs=4)
        out[f"lab_D{D}/cfg"] = cfg_to_array(cfg)
        out[f"lab_D{D}/labs"] = lab
        out[f"lab_D{D}/out"] = ref.window_labs(cfg, lab)
    np.savez_compressed(os.path.join(OUT, "window_golden.npz"), **out)
    print("done")


def nodur_training():
    """stdseg_no_dur* training goldens (own file and own RNG so that the older fixtures do not change)."""
    ref = RefLib()
    rng = np.random.default_rng(20260202)
    train = {}
    off, ftrs, labs = synth(rng, 6, 3, 40, 9, 7, seg_lo=1, seg_hi=6)
    for mt, D, segf in [("stdseg_no_dur_no_segtransftr", 4, 1), ("stdseg_no_dur_no_segtransftr", 3, 0), ("stdseg_no_dur", 4, 1),
                        ("stdseg_no_dur_no_transftr", 2, 1), ("stdseg_no_dur_no_segtransftr", 1, 0)]:
        cfg = make_config(mt, n_labs=7, n_base_ftrs=9, max_dur=D, extract_seg_ftrs=segf)
        train[f"{mt}_d{D}_s{segf}"] = (cfg, rng.uniform(-0.1, 0.1, ref.lambda_len(cfg)), off, ftrs, labs)
    offl, ftrsl, labsl = synth(rng, 3, 30, 60, 12, 11, seg_lo=2, seg_hi=14)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=11, n_base_ftrs=12, max_dur=10, extract_seg_ftrs=1)
    train["nodur_d10_segftr"] = (cfg, rng.uniform(-0.05, 0.05, ref.lambda_len(cfg)), offl, ftrsl, labsl)
    off3, ftrs3, labs3 = synth(rng, 4, 4, 40, 9, 5, states=3)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=15, n_base_ftrs=9, n_states=3)
    train["nodur_3state_d1"] = (cfg, rng.uniform(-0.25, 0.25, ref.lambda_len(cfg)), off3, ftrs3, labs3)
    out = {}
    for name, (cfg, lam, off, ftrs, labs) in train.items():
        grad, numer, logz = ref.fwdbwd(cfg, lam, off, ftrs, labs)
        out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/lam": lam, f"{name}/off": off, f"{name}/ftrs": ftrs,
                    f"{name}/labs": labs, f"{name}/grad": grad, f"{name}/numer": numer, f"{name}/logZ": logz})
        print(f"train {name}: lambda {len(lam)}, logZ {logz[:3]}, |grad|^2 {np.sum(grad ** 2):.12f}")
    np.savez_compressed(os.path.join(OUT, "train_nodur_golden.npz"), **out)


def nodur_nstate_training():
    """N-state segmental training (CRF_StdSegNStateNode_WithoutDurLab_WithoutSegTransFtr): every sub-state is a segment of its own."""
    ref = RefLib()
    rng = np.random.default_rng(20260303)
    train = {}
    off, ftrs, labs = synth(rng, 5, 6, 45, 8, 5, seg_lo=3, seg_hi=12, states=3)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=15, n_base_ftrs=8, n_states=3, max_dur=4, extract_seg_ftrs=1)
    train["nstate3_p5_d4_segftr"] = (cfg, rng.uniform(-0.1, 0.1, ref.lambda_len(cfg)), off, ftrs, labs)
    cfg = make_config("stdseg_no_dur_no_transftr", n_labs=15, n_base_ftrs=8, n_states=3, max_dur=3, extract_seg_ftrs=0)
    train["nstate3_p5_d3_notransftr"] = (cfg, rng.uniform(-0.1, 0.1, ref.lambda_len(cfg)), off, ftrs, labs)
    off2, ftrs2, labs2 = synth(rng, 4, 3, 30, 6, 4, seg_lo=1, seg_hi=9, states=2)   # short phones skip sub-states: illegal reference pairs
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=8, n_base_ftrs=6, n_states=2, max_dur=3, extract_seg_ftrs=1)
    train["nstate2_p4_d3_skips"] = (cfg, rng.uniform(-0.2, 0.2, ref.lambda_len(cfg)), off2, ftrs2, labs2)
    off3, ftrs3, labs3 = synth(rng, 3, 30, 70, 10, 6, seg_lo=6, seg_hi=30, states=3)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=18, n_base_ftrs=10, n_states=3, max_dur=10, extract_seg_ftrs=1)
    train["nstate3_p6_d10_segftr"] = (cfg, rng.uniform(-0.05, 0.05, ref.lambda_len(cfg)), off3, ftrs3, labs3)
    out = {}
    for name, (cfg, lam, off, ftrs, labs) in train.items():
        grad, numer, logz = ref.fwdbwd(cfg, lam, off, ftrs, labs)
        out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/lam": lam, f"{name}/off": off, f"{name}/ftrs": ftrs,
                    f"{name}/labs": labs, f"{name}/grad": grad, f"{name}/numer": numer, f"{name}/logZ": logz})
        print(f"train {name}: lambda {len(lam)}, logZ {logz[:3]}, numer {numer[:3]}, |grad|^2 {np.sum(grad ** 2):.12f}")
    np.savez_compressed(os.path.join(OUT, "train_nodur_nstate_golden.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "nstate":
        nodur_nstate_training()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "nodur":
        nodur_training()
        sys.exit(0)
    main()
