"""Phase times of a frame-level model with transition features at the cfg2 geometry (61 labels, 105 features for states and
transitions) on the TIMIT-shaped shard.  python tools/frame_tf_timing.py [n_utt]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))
import crf_b200  # noqa: E402
import workloads  # noqa: E402

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 462
off, ftrs, labs = workloads.timit_train_batch(0, n_utt)
kw = dict(workloads.cfg2_kwargs()); kw.update(use_trans_ftrs=1)
m = crf_b200.CrfGpu(crf_b200.make_config(**kw), device=0)
import numpy as np  # noqa: E402
m.set_lambda(np.random.default_rng(3).uniform(-0.02, 0.02, m.lambda_len))
m.stage(off, ftrs, labs)
for _ in range(3):
    m.fwdbwd_staged()
    m.synchronize()
ph = {k: round(m.phase_ms(k), 3) for k in ("score", "forward", "backward", "xi", "grad")}
print(n_utt, "utterances,", int(off[-1]), "frames;", ph, "->", round(float(off[-1]) / sum(ph.values()) / 1e3, 2), "M frames/s")
