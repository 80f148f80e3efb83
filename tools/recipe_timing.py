"""Phase times of the production-recipe shape (workloads.recipe_*), with the transition-score GEMM on its own.  python tools/recipe_timing.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))
import crf_b200  # noqa: E402
import workloads  # noqa: E402

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 128
off, f1, f2, labs = workloads.recipe_batch(n_utt)
m = crf_b200.CrfGpu(crf_b200.make_config(**workloads.recipe_kwargs()), device=0)
m.set_lambda(workloads.lam_for("recipe", m.lambda_len))
m.stage(off, f1, labs, ftrs2=f2)
for _ in range(3):
    m.fwdbwd_staged()
    m.synchronize()
print(n_utt, "utterances,", int(off[-1]), "frames, longest", int((off[1:] - off[:-1]).max()))
for ph in ("expand", "score", "forward", "trans_score", "backward", "xi", "grad"):
    print(ph, round(m.phase_ms(ph), 3))
