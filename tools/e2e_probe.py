import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/asr-craft_b200')
import numpy as np, crf_b200, workloads
off, ftrs, labs = workloads.timit_train_batch(0, 462)
m = crf_b200.CrfGpu(crf_b200.make_config(**workloads.cfg4_kwargs()))
lam = workloads.lam_for("cfg4", m.lambda_len)
t0 = time.perf_counter(); m.set_lambda(lam); print("set_lambda ms", 1e3 * (time.perf_counter() - t0))
t0 = time.perf_counter(); m.set_lambda(lam); print("set_lambda ms", 1e3 * (time.perf_counter() - t0))
pf = crf_b200.PinnedBuffer(ftrs.shape, np.float32); pf.array[...] = ftrs
pl = crf_b200.PinnedBuffer(labs.shape, np.uint32); pl.array[...] = labs
pg = crf_b200.PinnedBuffer((m.lambda_len,), np.float64); pn = crf_b200.PinnedBuffer((462,), np.float64); pz = crf_b200.PinnedBuffer((462,), np.float64)
for i in range(4):
    t0 = time.perf_counter(); m.fwdbwd(off, pf.array, pl.array, out=(pg.array, pn.array, pz.array)); dt = 1e3 * (time.perf_counter() - t0); ph = {k: round(m.phase_ms(k), 3) for k in ["expand", "score", "forward", "backward", "xi", "grad"]}; print("e2e ms", round(dt, 3), ph, "sum", round(sum(ph.values()), 3))
