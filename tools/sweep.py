#!/usr/bin/env python
"""Option sweep on the cfg4 bench workload: prints per-phase device times (CUDA events on the handle's stream) for each
setting.  Usage: python tools/sweep.py name=v1,v2,... [name2=...]   (cartesian product; options of crfgpu_set_option)"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))
import crf_b200  # noqa: E402
import workloads  # noqa: E402

axes = [(a.split("=")[0], [int(v) for v in a.split("=")[1].split(",")]) for a in sys.argv[1:]]
off, ftrs, labs = workloads.timit_train_batch(0, 462)
m = crf_b200.CrfGpu(crf_b200.make_config(**workloads.cfg4_kwargs()))
m.set_lambda(workloads.lam_for("cfg4", m.lambda_len))
names = ["score", "forward", "backward", "xi", "grad"]
for combo in itertools.product(*[v for _, v in axes]) if axes else [()]:
    for (k, _), v in zip(axes, combo):
        m.set_option(k, v)
    m.stage(off, ftrs, labs)
    best = None
    for _ in range(6):
        m.fwdbwd_staged(); m.synchronize()
        ph = [m.phase_ms(k) for k in names]
        best = ph if best is None else [min(a, b) for a, b in zip(best, ph)]
    print(" ".join(f"{k}={v}" for (k, _), v in zip(axes, combo)), "|", " ".join(f"{k} {x:.3f}" for k, x in zip(names, best)),
          f"| sum {sum(best):.3f} ms", flush=True)
m.close()
