"""Generates tests/golden/joined_golden.npz from the UNMODIFIED reference (oracle/_ref/libcrfref.so): window streams with context
frames (first_frame_left_ctx_ftrs / first_frame_right_ctx_ftrs / last_frame_right_ctx_ftrs / boundary_delta_ftrs,
CRF/src/io/CRF_InFtrStream_SeqMultiWindow.cpp:897-1110) and a second feature stream joined behind the first
(CRF_FeatureStreamManager::join, CRF/src/io/CRF_FeatureStreamManager.cpp:482-500) -- the layout of the TIMIT recipe
(demo/segmental-timit-demo.cfg.in:16-33: state features = segment features of stream 1, transition features = 13 context frames
of the padded stream 2).  Three groups of cases: window vectors, training (gradient / numerator / logZ), Viterbi.

    python tests/golden/make_golden_joined.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.binding import Config, RefLib, make_config, window_width  # noqa: E402
from make_golden import synth  # noqa: E402


def cfg_to_array(cfg):
    """every field of the configuration, the context / joined-stream options included (make_golden.py keeps the 17 older fields so
    that its fixtures stay byte-identical)"""
    return np.array([float(getattr(cfg, f[0])) for f in Config._fields_], np.float64)

OUT = os.path.dirname(os.path.abspath(__file__))


def stream(rng, off, F, lc, rc):
    """rows of a stream that carries lc + rc context frames per utterance"""
    return rng.random((int(off[-1]) + (len(off) - 1) * (lc + rc), F), dtype=np.float32)


def main():
    ref = RefLib()
    rng = np.random.default_rng(20261018)
    out = {}
    # ---- window vectors ----
    wins = {
        "win_seg_ctx": dict(max_dur=4, extract_seg_ftrs=1, left_ctx=2, right_ctx=3),
        "win_first_ctx": dict(max_dur=4, extract_seg_ftrs=0, left_ctx=2, right_ctx=1),
        "win_bdelta": dict(max_dur=3, extract_seg_ftrs=0, left_ctx=3, right_ctx=1, boundary_delta=1),
        "win_frame_ctx": dict(max_dur=1, left_ctx=2, right_ctx=2),
        "win_join_ctx2": dict(max_dur=5, extract_seg_ftrs=1, n_base_ftrs2=3, left_ctx2=2, right_ctx2=2),
        "win_join_bdelta2": dict(max_dur=5, extract_seg_ftrs=1, n_base_ftrs2=3, left_ctx2=3, right_ctx2=2, boundary_delta2=1),
        "win_join_seg2": dict(max_dur=10, extract_seg_ftrs=1, left_ctx=1, n_base_ftrs2=2, extract_seg_ftrs2=1, right_ctx2=1),
    }
    for name, kw in wins.items():
        cfg = make_config("stdseg_no_dur_no_segtransftr" if kw["max_dur"] > 1 else "stdframe", n_labs=5, n_base_ftrs=4, **kw)
        T = 23
        off = np.array([0, T], np.uint32)
        f1 = stream(rng, off, 4, cfg.left_ctx, cfg.right_ctx)
        f2 = stream(rng, off, cfg.n_base_ftrs2, cfg.left_ctx2, cfg.right_ctx2) if cfg.n_base_ftrs2 else np.zeros((0, 1), np.float32)
        w = ref.window_ftrs(cfg, f1, f2 if cfg.n_base_ftrs2 else None)
        out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/ftrs": f1, f"{name}/ftrs2": f2, f"{name}/win": np.nan_to_num(w, nan=0.0)})
        print(f"{name}: width {w.shape[2]}")
    # ---- training ----
    train = {}
    off, _, labs = synth(rng, 6, 1, 40, 5, 6)
    cfg = make_config("stdframe", n_labs=6, n_base_ftrs=5, left_ctx=2, right_ctx=1)
    train["frame_ctx"] = (cfg, 0.2, off, labs)
    off, _, labs = synth(rng, 5, 1, 45, 4, 4, seg_lo=1, seg_hi=6)
    cfg = make_config("stdseg", n_labs=12, n_base_ftrs=4, max_dur=3, n_actual_labs=4, extract_seg_ftrs=1, left_ctx=1, right_ctx=2)
    train["stdseg_ctx"] = (cfg, 0.05, off, labs)
    # the recipe's layout in small: segment features of stream 1 -> state features, context frames of stream 2 -> transition features
    off, _, labs = synth(rng, 5, 1, 45, 6, 7, seg_lo=1, seg_hi=9)
    w1, w2 = window_width(6, 4, 1), window_width(5, 4, 0, 2, 2)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=7, n_base_ftrs=6, max_dur=4, n_actual_labs=7, extract_seg_ftrs=1,
                      n_base_ftrs2=5, left_ctx2=2, right_ctx2=2, use_trans_ftrs=1, state_fidx=(0, w1 - 1), trans_fidx=(w1, w1 + w2 - 1))
    train["nodur_joined_recipe"] = (cfg, 0.05, off, labs)
    w2 = window_width(5, 4, 0, 3, 2, 1)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=7, n_base_ftrs=6, max_dur=4, n_actual_labs=7, extract_seg_ftrs=1,
                      n_base_ftrs2=5, left_ctx2=3, right_ctx2=2, boundary_delta2=1, use_trans_ftrs=1, state_fidx=(0, w1 - 1),
                      trans_fidx=(w1, w1 + w2 - 1))
    train["nodur_joined_bdelta"] = (cfg, 0.05, off, labs)
    off, _, labs = synth(rng, 4, 2, 50, 4, 5)
    cfg = make_config("stdframe", n_labs=5, n_base_ftrs=4, n_base_ftrs2=3, left_ctx2=1, right_ctx2=1, use_trans_ftrs=1,
                      state_fidx=(0, 3), trans_fidx=(4, 12))
    train["frame_joined_transftr"] = (cfg, 0.2, off, labs)
    # no transition features: the joined vector only widens the state features (tied / native no_dur lattice paths)
    off, _, labs = synth(rng, 4, 3, 40, 4, 5, seg_lo=1, seg_hi=7)
    cfg = make_config("stdseg_no_dur", n_labs=5, n_base_ftrs=4, max_dur=5, n_actual_labs=5, extract_seg_ftrs=1, right_ctx=1,
                      n_base_ftrs2=3, left_ctx2=1, right_ctx2=1)
    train["nodur_joined_state_only"] = (cfg, 0.05, off, labs)
    for name, (cfg, scale, off, labs) in train.items():
        f1 = stream(rng, off, cfg.n_base_ftrs, cfg.left_ctx, cfg.right_ctx)
        f2 = stream(rng, off, cfg.n_base_ftrs2, cfg.left_ctx2, cfg.right_ctx2) if cfg.n_base_ftrs2 else np.zeros((0, 1), np.float32)
        lam = rng.uniform(-scale, scale, ref.lambda_len(cfg))
        grad, numer, logz = ref.fwdbwd(cfg, lam, off, f1, labs, ftrs2=f2 if cfg.n_base_ftrs2 else None)
        out.update({f"train_{name}/cfg": cfg_to_array(cfg), f"train_{name}/lam": lam, f"train_{name}/off": off, f"train_{name}/ftrs": f1,
                    f"train_{name}/ftrs2": f2, f"train_{name}/labs": labs, f"train_{name}/grad": grad, f"train_{name}/numer": numer,
                    f"train_{name}/logZ": logz})
        print(f"train_{name}: lambda {len(lam)}, logZ {logz[:3]}, |grad|^2 {np.sum(grad ** 2):.12f}")
    # ---- Viterbi ----
    lens = [1, 2, 3, 5, 9, 30, 47]
    voff = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    n_vit = 0
    for (P, N, D, lc2) in [(5, 1, 1, 1), (5, 3, 1, 2), (7, 1, 3, 2), (4, 3, 2, 1), (6, 1, 5, 3)]:
        w1 = window_width(4, D, 1 if D > 1 else 0)
        w2 = window_width(3, D, 0, lc2, lc2)
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P * N, n_base_ftrs=4, n_states=N, max_dur=D, extract_seg_ftrs=1 if D > 1 else 0,
                          n_base_ftrs2=3, left_ctx2=lc2, right_ctx2=lc2, use_trans_ftrs=1, state_fidx=(0, w1 - 1), trans_fidx=(w1, w1 + w2 - 1))
        n = ref.lambda_len(cfg)
        f1 = stream(rng, voff, 4, 0, 0)
        f2 = stream(rng, voff, 3, lc2, lc2)
        q1, q2 = (np.round(f1 * 2) / 2).astype(np.float32), (np.round(f2 * 2) / 2).astype(np.float32)
        for kind, lam, a, b in [("rand", rng.uniform(-0.5, 0.5, n), f1, f2), ("quant", np.round(rng.uniform(-1, 1, n) * 2) / 2, q1, q2)]:
            name = f"vit_{kind}_P{P}N{N}D{D}c{lc2}"
            segs, cost, logz = ref.viterbi(cfg, lam, voff, a, b)
            nseg = np.array([len(sg[0]) for sg in segs], np.uint32)
            out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/lam": lam, f"{name}/off": voff, f"{name}/ftrs": a, f"{name}/ftrs2": b,
                        f"{name}/nseg": nseg, f"{name}/cost": cost,
                        f"{name}/lab": np.concatenate([sg[0] for sg in segs]), f"{name}/dur": np.concatenate([sg[1] for sg in segs]),
                        f"{name}/phn": np.concatenate([sg[2] for sg in segs])})
            n_vit += 1
    print("viterbi cases with a joined stream:", n_vit)
    np.savez_compressed(os.path.join(OUT, "joined_golden.npz"), **out)


if __name__ == "__main__":
    main()
