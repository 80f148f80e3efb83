"""Host-side timeline of the trainer's end-to-end step on the cfg4 shard (the loop bench.py times as `e2e`): how long each call of the
step keeps the host, and the step as a whole.  python tools/e2e_steps.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))
import crf_b200  # noqa: E402
import workloads  # noqa: E402

off, ftrs, labs = workloads.timit_train_batch(0, 462)
m = crf_b200.CrfGpu(crf_b200.make_config(**workloads.cfg4_kwargs()))
m.set_lambda(workloads.lam_for("cfg4", m.lambda_len))
pf = crf_b200.PinnedBuffer(ftrs.shape, np.float32); pf.array[...] = ftrs
pl = crf_b200.PinnedBuffer(labs.shape, np.uint32); pl.array[...] = labs
pn = crf_b200.PinnedBuffer((462,), np.float64); pz = crf_b200.PinnedBuffer((462,), np.float64)
names = ["stage", "fwdbwd_staged", "prefetch", "fetch", "sgd_update"]
acc = np.zeros(len(names)); tot = 0.0; n = 0
for it in range(12):
    t = [time.perf_counter()]
    m.stage(off, pf.array, pl.array); t.append(time.perf_counter())
    m.fwdbwd_staged(); t.append(time.perf_counter())
    m.prefetch(off, pf.array, pl.array); t.append(time.perf_counter())
    m.fetch_fwdbwd(out=(None, pn.array, pz.array)); t.append(time.perf_counter())
    m.sgd_update(1.0, lr=1e-13); t.append(time.perf_counter())
    if it >= 4:
        acc += np.diff(t); tot += t[-1] - t[0]; n += 1
print("host ms per call:", {k: round(1e3 * v / n, 3) for k, v in zip(names, acc)}, "step", round(1e3 * tot / n, 3))
print("device phases ms:", {k: round(m.phase_ms(k), 3) for k in ["score", "forward", "backward", "xi", "grad"]})
