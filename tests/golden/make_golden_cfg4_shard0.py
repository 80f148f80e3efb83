#!/usr/bin/env python
"""Golden for the BENCH workload itself: cfg4 (stdseg, 61 phones x maxDur 10, 850 segment features) on
workloads.timit_train_batch(0, 462) with workloads.lam_for("cfg4") -- the minibatch bench.py times on rank 0.

Produced by the UNMODIFIED reference (oracle/_ref/libcrfref.so, CRF_NewGradBuilder_StdSeg::buildGradient over
CRF_StdSegStateNode) when it is built here, else by the C restatement (the file records which).  ~5 minutes on 8
cores.  Output: tests/golden/cfg4_shard0_golden.npz with per-utterance numerators and logZ (fp64), the gradient
(fp64 sums: ||g||^2, max|g|, sum g; all entries as float32 -- 2^-24 relative, far below the 1e-4 parity bound) and a
small pinned summary tests/golden/cfg4_shard0_pin.json that bench.py compares its log-likelihood against.

    python tests/golden/make_golden_cfg4_shard0.py [n_threads]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import workloads  # noqa: E402
from oracle.binding import OracleLib, RefLib, have_ref, make_config  # noqa: E402


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    lib, kind = (RefLib(), "reference") if have_ref() else (OracleLib(), "port")
    cfg = make_config(**workloads.cfg4_kwargs())
    lam = workloads.lam_for("cfg4", lib.lambda_len(cfg))
    off, ftrs, labs = workloads.timit_train_batch(0, 462)
    t0 = time.time()
    g, n, z = lib.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=threads)
    sec = time.time() - t0
    ll = float((n - z).sum())
    print(f"{kind}: {sec:.0f} s, sum(numer - logZ) = {ll:.10f}, sum numer = {n.sum():.10f}, sum logZ = {z.sum():.10f}, "
          f"||g||^2 = {np.sum(g * g):.10e}, max|g| = {np.abs(g).max():.6f}")
    np.savez_compressed(os.path.join(HERE, "cfg4_shard0_golden.npz"), numer=n, logZ=z, grad32=g.astype(np.float32),
                        grad_sq=np.sum(g * g), grad_absmax=np.abs(g).max(), grad_sum=g.sum(), kind=kind)
    pin = {"workload": "cfg4, workloads.timit_train_batch(0, 462), workloads.lam_for('cfg4')", "producer": kind,
           "n_utt": 462, "frames": int(off[-1]), "loglik": ll, "sum_numer": float(n.sum()), "sum_logZ": float(z.sum()),
           "grad_sq": float(np.sum(g * g)), "grad_absmax": float(np.abs(g).max())}
    json.dump(pin, open(os.path.join(HERE, "cfg4_shard0_pin.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
