#!/usr/bin/env python
"""BASELINE cfg5 (stress) on one GPU: stdseg_no_dur_no_segtransftr, 1024 phones, maxDur 30, 64 utterances x 2000 frames.
Prints one JSON line with training (forward-backward + gradient) and Viterbi throughput and the per-phase device times."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))
import crf_b200  # noqa: E402
import workloads  # noqa: E402

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
off, ftrs, labs = workloads.cfg5_batch(n_utt, n_frames)
m = crf_b200.CrfGpu(crf_b200.make_config(**workloads.cfg5_kwargs()))
m.set_lambda(workloads.lam_for("cfg5", m.lambda_len))
N = int(off[-1])
out = {"workload": "cfg5", "phones": 1024, "max_dur": 30, "seg_ftrs": 542, "utts": n_utt, "frames": N, "lambda_len": m.lambda_len}
m.stage(off, ftrs, labs)
best = None
for _ in range(3):
    m.fwdbwd_staged(); m.synchronize()
    ph = {k: m.phase_ms(k) for k in ["score", "forward", "backward", "xi", "grad"]}
    if best is None or sum(ph.values()) < sum(best.values()):
        best = ph
out["train_phases_ms"] = best
out["train_frames_per_s"] = N / (sum(best.values()) / 1e3)
t0 = time.perf_counter()
g, nu, z = m.fwdbwd(off, ftrs, labs)
out["train_e2e_frames_per_s"] = N / (time.perf_counter() - t0)
out["loglik"] = float(np.sum(nu - z))
if "--no-viterbi" not in sys.argv:
    m.stage(off, ftrs)
    m.viterbi_staged(); m.synchronize()
    m.viterbi_staged(); m.synchronize()
    vs, vr = m.phase_ms("viterbi_score"), m.phase_ms("viterbi")
    out["viterbi_phases_ms"] = {"score": vs, "recursion": vr}
    out["viterbi_frames_per_s"] = N / ((vs + vr) / 1e3)
print(json.dumps(out))
m.close()
