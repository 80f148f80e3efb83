"""Generates tests/golden/viterbi_lm_golden.npz from the UNMODIFIED reference (oracle/_ref/libcrfref.so): nStateDecode with an input
language model (CRF_ViterbiDecoder_StdSeg_NoSegTransFtr.cpp:1369-2398, lm_fst != NULL) for the class of LMs the device implements --
complete phone-bigram LMs in the topology of the decoder's own free-phone LM (one state per phone; createFreePhoneLmFst :1270-1348):
random costs, quantised costs (ties in every frame), phone states that are not final, with and without transition features; and, for N states
per phone, unigram + exit-cost LMs in the topology of the N-state free-phone LM (epsilon arcs back to the start state; a plain phone
insertion penalty among them).

    python tests/golden/make_golden_lm.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.binding import RefLib, make_config  # noqa: E402
from make_golden_joined import cfg_to_array  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = RefLib()
    rng = np.random.default_rng(20261019)
    out = {}
    lens = [1, 2, 3, 5, 9, 30, 47, 120]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    n = 0
    for (P, D, seg, tf) in [(5, 1, 0, 0), (7, 3, 1, 0), (6, 5, 1, 0), (6, 4, 1, 1), (48, 10, 1, 0)]:
        w = 4 if (D == 1 or not seg) else 8 * 4 + D
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=4, max_dur=D, extract_seg_ftrs=seg,
                          use_trans_ftrs=tf, trans_fidx=(0, min(w, 12) - 1))
        nl = ref.lambda_len(cfg)
        f = rng.random((int(off[-1]), 4), dtype=np.float32)
        qf = (np.round(f * 2) / 2).astype(np.float32)
        for kind in ("rand", "quant", "nofinal"):
            lam = rng.uniform(-0.5, 0.5, nl)
            st = rng.uniform(0, 3, P).astype(np.float32); bg = rng.uniform(0, 3, (P, P)).astype(np.float32); fin = rng.uniform(0, 2, P).astype(np.float32)
            x = f
            if kind == "quant":
                lam = np.round(rng.uniform(-1, 1, nl) * 2) / 2
                st, bg, fin, x = np.round(st), np.round(bg), np.round(fin), qf
            if kind == "nofinal":
                fin[::2] = np.inf
            segs, cost, _ = ref.viterbi(cfg, lam, off, x, lm=(st, bg, fin))
            name = f"{kind}_P{P}D{D}s{seg}t{tf}"
            nseg = np.array([len(sg[0]) for sg in segs], np.uint32)
            out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/lam": lam, f"{name}/off": off, f"{name}/ftrs": x,
                        f"{name}/lm_start": st, f"{name}/lm_bigram": bg, f"{name}/lm_final": fin, f"{name}/nseg": nseg, f"{name}/cost": cost,
                        f"{name}/lab": np.concatenate([sg[0] for sg in segs]), f"{name}/dur": np.concatenate([sg[1] for sg in segs]),
                        f"{name}/phn": np.concatenate([sg[2] for sg in segs])})
            n += 1
    # N states per phone: unigram + exit costs in the topology of the N-state free-phone LM (epsilon arcs back to the start state)
    for (P, N, D, seg) in [(5, 3, 1, 0), (4, 3, 2, 1), (6, 2, 3, 1), (61, 3, 1, 0)]:
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P * N, n_base_ftrs=4, n_states=N, max_dur=D, extract_seg_ftrs=seg)
        nl = ref.lambda_len(cfg)
        f = rng.random((int(off[-1]), 4), dtype=np.float32)
        qf = (np.round(f * 2) / 2).astype(np.float32)
        for kind in ("rand", "quant", "nofinal", "penalty"):
            lam = rng.uniform(-0.5, 0.5, nl)
            uni = rng.uniform(0, 3, P).astype(np.float32); ex = rng.uniform(0, 3, P).astype(np.float32); fin = rng.uniform(0, 2, P).astype(np.float32)
            x = f
            if kind == "quant":
                lam = np.round(rng.uniform(-1, 1, nl) * 2) / 2
                uni, ex, fin, x = np.round(uni), np.round(ex), np.round(fin), qf
            if kind == "nofinal":
                fin[::2] = np.inf
            if kind == "penalty":      # a plain phone insertion penalty
                uni[:] = 0.0; ex[:] = 1.5; fin[:] = 0.0
            segs, cost, _ = ref.viterbi(cfg, lam, off, x, lm=(uni, ex, fin))
            name = f"{kind}_P{P}N{N}D{D}s{seg}"
            nseg = np.array([len(sg[0]) for sg in segs], np.uint32)
            out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/lam": lam, f"{name}/off": off, f"{name}/ftrs": x,
                        f"{name}/lm_start": uni, f"{name}/lm_bigram": ex, f"{name}/lm_final": fin, f"{name}/nseg": nseg, f"{name}/cost": cost,
                        f"{name}/lab": np.concatenate([sg[0] for sg in segs]), f"{name}/dur": np.concatenate([sg[1] for sg in segs]),
                        f"{name}/phn": np.concatenate([sg[2] for sg in segs])})
            n += 1
    # beam pruning (nStateDecode's input_beam > 0), one state per phone, with the free-phone LM and with a bigram LM
    nb = 0
    for (P, D, seg) in [(5, 1, 0), (7, 3, 1), (6, 5, 1), (20, 4, 1), (48, 10, 1)]:
        w = 4 if (D == 1 or not seg) else 8 * 4 + D
        cfg = make_config("stdseg_no_dur_no_segtransftr", n_labs=P, n_base_ftrs=4, max_dur=D, extract_seg_ftrs=seg)
        nl = ref.lambda_len(cfg)
        f = rng.random((int(off[-1]), 4), dtype=np.float32)
        qf = (np.round(f * 2) / 2).astype(np.float32)
        for kind in ("rand", "quant"):
            lam = rng.uniform(-0.5, 0.5, nl)
            st = rng.uniform(0, 3, P).astype(np.float32); bg = rng.uniform(0, 3, (P, P)).astype(np.float32); fin = rng.uniform(0, 2, P).astype(np.float32)
            x = f
            if kind == "quant":
                lam = np.round(rng.uniform(-1, 1, nl) * 2) / 2
                st, bg, fin, x = np.round(st), np.round(bg), np.round(fin), qf
            for use_lm in (0, 1):
                for beam in (0.5, 2.0):
                    lm = (st, bg, fin) if use_lm else None
                    segs, cost, _ = ref.viterbi(cfg, lam, off, x, lm=lm, beam=beam)
                    name = f"beam{beam}_{kind}_P{P}D{D}s{seg}lm{use_lm}"
                    nseg = np.array([len(sg[0]) for sg in segs], np.uint32)
                    out.update({f"{name}/cfg": cfg_to_array(cfg), f"{name}/lam": lam, f"{name}/off": off, f"{name}/ftrs": x, f"{name}/beam": np.array([beam]),
                                f"{name}/lm_start": st if use_lm else np.zeros(0, np.float32), f"{name}/lm_bigram": bg if use_lm else np.zeros(0, np.float32),
                                f"{name}/lm_final": fin if use_lm else np.zeros(0, np.float32), f"{name}/nseg": nseg, f"{name}/cost": cost,
                                f"{name}/lab": np.concatenate([sg[0] for sg in segs] + [np.zeros(0, np.uint32)]), f"{name}/dur": np.concatenate([sg[1] for sg in segs] + [np.zeros(0, np.uint32)]),
                                f"{name}/phn": np.concatenate([sg[2] for sg in segs] + [np.zeros(0, np.uint32)])})
                    nb += 1
    print("beam-pruned Viterbi cases:", nb)
    print("LM-constrained Viterbi cases:", n)
    np.savez_compressed(os.path.join(OUT, "viterbi_lm_golden.npz"), **out)


if __name__ == "__main__":
    main()
