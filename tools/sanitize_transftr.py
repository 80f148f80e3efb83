#!/usr/bin/env python
"""Tiny pass over the transition-feature recursions (crf_dp_transftr.cu: bulk-copied dense matrices, fused backward step): finite results on
small cases (compute-sanitizer is closed on the GPU pool; parity is the tests job): frame-level with an odd and an even label count and
N states, the segmental kernels with an even phone count; utterances of 1, 2 and more frames."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))
import crf_b200  # noqa: E402

rng = np.random.default_rng(1)
lens = np.array([1, 2, 3, 9, 17])
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
N = int(off[-1])
f = rng.random((N, 5), dtype=np.float32)
for kw in (dict(model="stdframe", n_labs=7), dict(model="stdframe", n_labs=6), dict(model="stdframe", n_labs=6, n_states=3),
           dict(model="stdseg_no_dur_no_segtransftr", n_labs=6, max_dur=3, extract_seg_ftrs=1, trans_fidx=(0, 4))):
    kw = dict(kw)
    model = kw.pop("model")
    P = kw["n_labs"] // kw.get("n_states", 1)
    labs = (np.repeat(rng.integers(0, P, N), 1) * kw.get("n_states", 1)).astype(np.uint32)
    m = crf_b200.CrfGpu(crf_b200.make_config(model, n_base_ftrs=5, use_trans_ftrs=1, **kw))
    m.set_lambda(rng.uniform(-0.05, 0.05, m.lambda_len))
    g, n, z = m.fwdbwd(off, f, labs)
    assert np.all(np.isfinite(z)) and np.all(np.isfinite(g)), (model, kw)
    m.close()
print("sanitize_transftr ok")
