import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "asr-craft_b200"); sys.path.insert(0, "tests")
import numpy as np, crf_b200
from helpers import load_cases
c = load_cases("train_golden.npz")["stdseg_d10_segftr"]
m = crf_b200.CrfGpu(crf_b200.copy_config(c["cfg"]))
m.set_lambda(c["lam"])
g, n, z = m.fwdbwd(c["off"], c["ftrs"], c["labs"])
print(m.plan_info())
print(np.abs(z - c["logZ"]).max())
