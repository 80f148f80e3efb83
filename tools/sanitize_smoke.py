#!/usr/bin/env python
"""Small end-to-end pass over the paths added in round 2 (virtual windows, joined / context windows, phone LM + beam decoding) for
`compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`; sizes are tiny because every kernel runs instrumented."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))
import crf_b200  # noqa: E402

rng = np.random.default_rng(0)
lens = rng.integers(3, 40, 6)
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
N = int(off[-1])
labs = np.repeat(rng.integers(0, 6, N // 3 + 1), 3)[:N].astype(np.uint32)
# stdseg with segment features: virtual windows, contraction-sliced lattice kernels, TMA-fed GEMMs
f = rng.random((N, 8), dtype=np.float32)
m = crf_b200.CrfGpu(crf_b200.make_config("stdseg", n_labs=30, n_base_ftrs=8, max_dur=5, n_actual_labs=6, extract_seg_ftrs=1))
m.set_lambda(rng.uniform(-0.05, 0.05, m.lambda_len))
g, n, z = m.fwdbwd(off, f, labs)
assert np.all(np.isfinite(z))
m.stage(off, f, labs); m.fwdbwd_staged(); m.prefetch(off, f, labs); m.fetch_fwdbwd(); m.stage(off, f, labs); m.fwdbwd_staged(); m.fetch_fwdbwd()
m.close()
# the recipe's layout: joined second stream with context frames, transition features
f2 = rng.random((N + 4 * (len(off) - 1), 5), dtype=np.float32)
w1, w2 = 8 * 8 + 4, 5 * 5
cfg = crf_b200.make_config("stdseg_no_dur_no_segtransftr", n_labs=6, n_base_ftrs=8, max_dur=4, extract_seg_ftrs=1, n_base_ftrs2=5, left_ctx2=2, right_ctx2=2,
                           use_trans_ftrs=1, state_fidx=(0, w1 - 1), trans_fidx=(w1, w1 + w2 - 1))
m = crf_b200.CrfGpu(cfg)
m.set_lambda(rng.uniform(-0.05, 0.05, m.lambda_len))
g, n, z = m.fwdbwd(off, f, labs, ftrs2=f2)
assert np.all(np.isfinite(z))
segs, cost = m.viterbi(off, f, ftrs2=f2)
m.close()
# decoding with a phone LM and a beam
cfg = crf_b200.make_config("stdseg_no_dur_no_segtransftr", n_labs=6, n_base_ftrs=8, max_dur=4, extract_seg_ftrs=1)
m = crf_b200.CrfGpu(cfg)
m.set_lambda(rng.uniform(-0.5, 0.5, m.lambda_len))
m.set_phone_lm(rng.uniform(0, 2, 6).astype(np.float32), rng.uniform(0, 2, (6, 6)).astype(np.float32), rng.uniform(0, 1, 6).astype(np.float32))
m.set_beam(0.5)
segs, cost = m.viterbi(off, f)
assert np.all(np.isfinite(cost))
m.close()
print("sanitize_smoke ok")
