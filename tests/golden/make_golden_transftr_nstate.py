"""Generates tests/golden/train_transftr_nstate_golden.npz from the UNMODIFIED reference (oracle/_ref/libcrfref.so): transition FEATURES
(crf_featuremap=stdtrans) with N states per label -- frame-level (CRF_StdNStateNode, CRF/src/nodes/CRF_StdNStateNode.cpp:65-108: diag /
offDiag / dense blocks of computeTransMatrixValue) and segmental without duration labels
(CRF_StdSegNStateNode_WithoutDurLab_WithoutSegTransFtr.cpp:38-80: the same three blocks from the duration-1 window of the node).
Only the legal pairs of the N-state map carry weights (CRF_StdFeatureMap.cpp:369-405).

    python tests/golden/make_golden_transftr_nstate.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.binding import RefLib, make_config, window_width  # noqa: E402
from make_golden import cfg_to_array, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = RefLib()
    rng = np.random.default_rng(20261019)
    cases = {}
    # ---- frame-level ----
    off, ftrs, labs = synth(rng, 5, 3, 40, 7, 4, seg_lo=3, seg_hi=9, states=3)
    cfg = make_config("stdframe", n_labs=12, n_base_ftrs=7, n_states=3, use_trans_ftrs=1, trans_fidx=(0, 6))
    cases["frame_n3_all"] = (cfg, 0.2, off, ftrs, labs)
    cfg = make_config("stdframe", n_labs=12, n_base_ftrs=7, n_states=3, use_trans_ftrs=1, trans_fidx=(2, 4), use_trans_bias=0)
    cases["frame_n3_slice_nobias"] = (cfg, 0.2, off, ftrs, labs)
    off2, ftrs2, labs2 = synth(rng, 4, 1, 30, 5, 5, seg_lo=1, seg_hi=6, states=2)      # one-frame phones skip the second sub-state
    cfg = make_config("stdframe", n_labs=10, n_base_ftrs=5, n_states=2, use_trans_ftrs=1, trans_fidx=(0, 4))
    cases["frame_n2_skips"] = (cfg, 0.2, off2, ftrs2, labs2)
    off3, ftrs3, labs3 = synth(rng, 3, 20, 50, 6, 42, seg_lo=3, seg_hi=8, states=3)     # 126 labels: the largest 3-state label set of the kernels
    cfg = make_config("stdframe", n_labs=126, n_base_ftrs=6, n_states=3, use_trans_ftrs=1, trans_fidx=(0, 5), state_fidx=(1, 5))
    cases["frame_n3_126labs"] = (cfg, 0.1, off3, ftrs3, labs3)
    # the segmental node classes with window length 1
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=12, n_base_ftrs=7, n_states=3, use_trans_ftrs=1, trans_fidx=(0, 6))
    cases["nodur_n3_d1"] = (cfg, 0.2, off, ftrs, labs)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=7, n_base_ftrs=7, use_trans_ftrs=1, trans_fidx=(0, 6))
    o1, f1, l1 = synth(rng, 4, 1, 30, 7, 7)
    cases["nodur_n1_d1"] = (cfg, 0.2, o1, f1, l1)
    # ---- segmental, N states per phone ----
    w = window_width(7, 4, 1)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=12, n_base_ftrs=7, n_states=3, max_dur=4, extract_seg_ftrs=1,
                      use_trans_ftrs=1, trans_fidx=(0, w - 1))
    cases["nodur_n3_d4_all"] = (cfg, 0.05, off, ftrs, labs)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=10, n_base_ftrs=5, n_states=2, max_dur=3, extract_seg_ftrs=1,
                      use_trans_ftrs=1, trans_fidx=(5, 14), state_fidx=(0, 8 * 5 - 1), use_trans_bias=0)
    cases["nodur_n2_d3_skips_slice_nobias"] = (cfg, 0.1, off2, ftrs2, labs2)
    off4, ftrs4, labs4 = synth(rng, 3, 30, 70, 6, 6, seg_lo=6, seg_hi=30, states=3)
    cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=18, n_base_ftrs=6, n_states=3, max_dur=10, extract_seg_ftrs=1,
                      use_trans_ftrs=1, trans_fidx=(0, 5 * 6 - 1))
    cases["nodur_n3_d10"] = (cfg, 0.03, off4, ftrs4, labs4)
    out = {}
    for name, (cfg, scale, off, ftrs, labs) in cases.items():
        lam = rng.uniform(-scale, scale, ref.lambda_len(cfg))
        grad, numer, logz = ref.fwdbwd(cfg, lam, off, ftrs, labs)
        out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/lam": lam, f"{name}/off": off, f"{name}/ftrs": ftrs,
                    f"{name}/labs": labs, f"{name}/grad": grad, f"{name}/numer": numer, f"{name}/logZ": logz})
        print(f"{name}: lambda {len(lam)}, logZ {logz[:3]}, numer {numer[:3]}, |grad|^2 {np.sum(grad ** 2):.12f}")
    np.savez_compressed(os.path.join(OUT, "train_transftr_nstate_golden.npz"), **out)


if __name__ == "__main__":
    main()
