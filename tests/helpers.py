"""Shared test helpers: golden loading and seeded synthetic batches."""
import os

import numpy as np

from oracle.binding import Config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CFG_KEYS = [f[0] for f in Config._fields_]


def cfg_from_array(a):
    vals = [float(v) if k.endswith("_val") else int(v) for k, v in zip(CFG_KEYS, a)]
    return Config(*vals)


def load_cases(fname):
    z = np.load(os.path.join(GOLDEN, fname))
    cases = {}
    for key in z.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = z[key]
    for c in cases.values():
        c["cfg"] = cfg_from_array(c["cfg"])
    return cases


def synth_batch(rng, n_utt, t_lo, t_hi, F, P, seg_lo=1, seg_hi=8, states=1):
    lens = rng.integers(t_lo, t_hi + 1, n_utt)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    ftrs = rng.random((int(off[-1]), F), dtype=np.float32)
    labs = np.zeros(int(off[-1]), np.uint32)
    for u in range(n_utt):
        t, prev = int(off[u]), -1
        while t < off[u + 1]:
            d = int(rng.integers(seg_lo, seg_hi + 1))
            lab = int(rng.integers(0, P))
            while lab == prev and P > 1:
                lab = int(rng.integers(0, P))
            e = min(t + d, int(off[u + 1]))
            if states == 1:
                labs[t:e] = lab
            else:
                n = e - t
                sub = np.minimum(np.arange(n) * states // max(n, 1), states - 1)
                labs[t:e] = lab * states + sub
            prev, t = lab, e
    return off, ftrs, labs


def split_segs(lab, dur, phn, nseg):
    out, b = [], 0
    for k in nseg:
        k = int(k)
        out.append((lab[b:b + k], dur[b:b + k], phn[b:b + k]))
        b += k
    return out
