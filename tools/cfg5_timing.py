"""Cycle counters of the native no_dur recursion on the cfg5 stress workload (CRFGPU_DP_TIMING=1).  python tools/cfg5_timing.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))
os.environ["CRFGPU_DP_TIMING"] = "1"
import crf_b200  # noqa: E402
import workloads  # noqa: E402

cfg = crf_b200.make_config(**workloads.cfg5_kwargs())
off, ftrs, labs = workloads.cfg5_batch()
m = crf_b200.CrfGpu(cfg, device=0)
m.set_lambda(workloads.lam_for("cfg5", m.lambda_len))
m.stage(off, ftrs, labs)
for _ in range(2):
    m.fwdbwd_staged()
for ph in ("score", "forward", "backward", "xi", "grad"):
    print(ph, m.phase_ms(ph))
