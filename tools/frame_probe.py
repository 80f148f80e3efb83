#!/usr/bin/env python
"""BASELINE cfg2 (frame-level CRF, 61 labels, 105 features) on the TIMIT-shaped batch: per-phase device times and frames/s."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))
import numpy as np, crf_b200, workloads
n_utt = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 462
transftr = "--transftr" in sys.argv        # crf_featuremap=stdtrans: 105 transition features per label pair
off, ftrs, labs = workloads.timit_train_batch(0, n_utt)
kw = workloads.cfg2_kwargs()
if transftr:
    kw["use_trans_ftrs"] = 1
m = crf_b200.CrfGpu(crf_b200.make_config(**kw))
m.set_lambda(np.random.default_rng(3).uniform(-0.02, 0.02, m.lambda_len) if transftr else workloads.lam_for("cfg2", m.lambda_len))
for a in sys.argv[1:]:                     # name=value pairs go to crfgpu_set_option
    if "=" in a:
        k, v = a.split("="); m.set_option(k, int(v))
m.stage(off, ftrs, labs)
names = ["score", "forward", "backward", "xi", "grad"]
best = None
for _ in range(6):
    m.fwdbwd_staged(); m.synchronize()
    ph = [m.phase_ms(k) for k in names]
    best = ph if best is None else [min(a, b) for a, b in zip(best, ph)]
N = int(off[-1])
print("cfg2" + (" stdtrans" if transftr else ""), n_utt, "utts", N, "frames |", " ".join(f"{k} {x:.3f}" for k, x in zip(names, best)), f"| sum {sum(best):.3f} ms -> {N / sum(best) / 1e3:.2f} M frames/s")
for _ in range(3):
    t0 = time.perf_counter(); m.fwdbwd(off, ftrs, labs); dt = time.perf_counter() - t0
print(f"e2e {1e3 * dt:.3f} ms -> {N / dt / 1e6:.2f} M frames/s (pageable host buffers)")
