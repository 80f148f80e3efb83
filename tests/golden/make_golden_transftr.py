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


if __name__ == "__main__":
    main()
