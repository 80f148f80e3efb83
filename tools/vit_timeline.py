"""Device / host timeline of the decode seam on the cfg3 workload (CRFGPU_VERBOSE=1): where the end-to-end time of
crfgpu_viterbi_batch goes.  python tools/vit_timeline.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))
os.environ["CRFGPU_VERBOSE"] = "1"
import crf_b200  # noqa: E402
import workloads  # noqa: E402

vcfg = crf_b200.make_config(**workloads.cfg3_kwargs())
voff, vftrs = workloads.cfg3_batch(1680)
vm = crf_b200.CrfGpu(vcfg, device=0)
vm.set_lambda(workloads.lam_for("cfg3", vm.lambda_len))
pin = crf_b200.PinnedBuffer(vftrs.shape, np.float32)
pin.array[...] = vftrs
for _ in range(2):
    vm.viterbi(voff, pin.array, raw=True)
t0 = time.perf_counter()
for _ in range(5):
    vm.viterbi(voff, pin.array, raw=True)
dt = (time.perf_counter() - t0) / 5
print(f"e2e {dt * 1e3:.3f} ms per batch, {float(voff[-1]) / dt / 1e6:.1f} M frames/s")
