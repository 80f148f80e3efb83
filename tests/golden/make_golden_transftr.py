"""Generates tests/golden/train_transftr_golden.npz from the UNMODIFIED reference (oracle/_ref/libcrfref.so): frame-level CRFs with
transition FEATURES (crf_featuremap=stdtrans: CRF_StdFeatureMap::computeTransMatrixValue / computeTransExpF with a feature slice,
CRF/src/ftrmaps/CRF_StdFeatureMap.cpp:94-110,197-223).  Kept apart from make_golden.py so that the other fixtures stay byte-identical.

    python tests/golden/make_golden_transftr.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.binding import RefLib, make_config  # noqa: E402
from make_golden import cfg_to_array, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = RefLib()
    rng = np.random.default_rng(20260202)
    cases = {}
    off, ftrs, labs = synth(rng, 6, 1, 40, 9, 7)
    # all 9 features for states and transitions, both biases
    cfg = make_config("stdframe", n_labs=7, n_base_ftrs=9, use_trans_ftrs=1, trans_fidx=(0, 8))
    cases["frame_transftr_all"] = (cfg, rng.uniform(-0.2, 0.2, ref.lambda_len(cfg)), off, ftrs, labs)
    # transition features on a slice only, no transition bias
    cfg = make_config("stdframe", n_labs=7, n_base_ftrs=9, use_trans_ftrs=1, trans_fidx=(2, 5), use_trans_bias=0)
    cases["frame_transftr_slice_nobias"] = (cfg, rng.uniform(-0.2, 0.2, ref.lambda_len(cfg)), off, ftrs, labs)
    off2, ftrs2, labs2 = synth(rng, 4, 20, 70, 13, 20, seg_lo=1, seg_hi=5)
    cfg = make_config("stdframe", n_labs=20, n_base_ftrs=13, use_trans_ftrs=1, trans_fidx=(0, 12), state_fidx=(3, 12))
    cases["frame_transftr_20labs"] = (cfg, rng.uniform(-0.1, 0.1, ref.lambda_len(cfg)), off2, ftrs2, labs2)
    # segmental models without duration labels: transition features from the duration-1 window of the frame the new segment starts in
    # (CRF_StdSegStateNode_WithoutDurLab_WithoutSegTransFtr; the production TIMIT recipe, demo/segmental-timit-demo.cfg.in:11-48)
    off3, ftrs3, labs3 = synth(rng, 5, 1, 45, 6, 7, seg_lo=1, seg_hi=9)
    w = 8 * 6 + 4
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=7, n_base_ftrs=6, max_dur=4, n_actual_labs=7, extract_seg_ftrs=1,
                      use_trans_ftrs=1, trans_fidx=(0, w - 1))
    cases["nodur_transftr_d4_all"] = (cfg, rng.uniform(-0.05, 0.05, ref.lambda_len(cfg)), off3, ftrs3, labs3)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=7, n_base_ftrs=6, max_dur=4, n_actual_labs=7, extract_seg_ftrs=1,
                      use_trans_ftrs=1, trans_fidx=(6, 17), state_fidx=(0, 8 * 6 - 1), use_trans_bias=0)
    cases["nodur_transftr_d4_slice_nobias"] = (cfg, rng.uniform(-0.05, 0.05, ref.lambda_len(cfg)), off3, ftrs3, labs3)
    off4, ftrs4, labs4 = synth(rng, 3, 25, 60, 5, 12, seg_lo=2, seg_hi=14)
    w = 8 * 5 + 10
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=12, n_base_ftrs=5, max_dur=10, n_actual_labs=12, extract_seg_ftrs=1,
                      use_trans_ftrs=1, trans_fidx=(0, 5 * 5 - 1))
    cases["nodur_transftr_d10"] = (cfg, rng.uniform(-0.03, 0.03, ref.lambda_len(cfg)), off4, ftrs4, labs4)
    out = {}
    for name, (cfg, lam, off, ftrs, labs) in cases.items():
        grad, numer, logz = ref.fwdbwd(cfg, lam, off, ftrs, labs)
        out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/lam": lam, f"{name}/off": off, f"{name}/ftrs": ftrs,
                    f"{name}/labs": labs, f"{name}/grad": grad, f"{name}/numer": numer, f"{name}/logZ": logz})
        print(f"{name}: lambda {len(lam)}, logZ {logz[:3]}, |grad|^2 {np.sum(grad ** 2):.12f}")
    np.savez_compressed(os.path.join(OUT, "train_transftr_golden.npz"), **out)

    # ---- decoding with transition features (CRF_ViterbiDecoder_StdSeg_NoSegTransFtr over nodes whose transMatrix comes from the
    #      duration-1 window of the frame the segment starts in) ----
    out = {}
    lens = [1, 2, 3, 5, 9, 30, 47]
    voff = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    for (P, N, D, segf) in [(5, 1, 1, 0), (5, 3, 1, 0), (7, 1, 3, 1), (4, 3, 2, 1), (6, 1, 5, 1)]:
        w = 4 if (D == 1 or not segf) else 8 * 4 + D
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P * N, n_base_ftrs=4, n_states=N, max_dur=D, extract_seg_ftrs=segf,
                          use_trans_ftrs=1, trans_fidx=(0, min(w, 12) - 1))
        n = ref.lambda_len(cfg)
        vf = rng.random((int(voff[-1]), 4), dtype=np.float32)
        qf = (np.round(vf * 2) / 2).astype(np.float32)
        for kind, lam, f in [("rand", rng.uniform(-0.5, 0.5, n), vf), ("ties", np.zeros(n), vf), ("quant", np.round(rng.uniform(-1, 1, n) * 2) / 2, qf)]:
            name = f"{kind}_P{P}N{N}D{D}s{segf}"
            segs, cost, logz = ref.viterbi(cfg, lam, voff, f)
            nseg = np.array([len(sg[0]) for sg in segs], np.uint32)
            out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/lam": lam, f"{name}/off": voff, f"{name}/ftrs": f,
                        f"{name}/nseg": nseg, f"{name}/cost": cost, f"{name}/logZ": logz,
                        f"{name}/lab": np.concatenate([sg[0] for sg in segs]), f"{name}/dur": np.concatenate([sg[1] for sg in segs]),
                        f"{name}/phn": np.concatenate([sg[2] for sg in segs])})
    print("viterbi cases with transition features:", len(out) // 10)
    np.savez_compressed(os.path.join(OUT, "viterbi_transftr_golden.npz"), **out)


if __name__ == "__main__":
    main()
