#!/usr/bin/env python
"""Pins for every leg of bench.py, so that a timed run can never again be "fast but wrong" (round 1 timed a kernel that
computed logZ 1.5 % low on the bench shard and compared its log-likelihood with nothing).

Produced by the UNMODIFIED reference (oracle/_ref/libcrfref.so) when it is built here, else by the C restatement; the file
records which.  Output tests/golden/bench_pins.npz:

  cfg4/numer, cfg4/logZ   per utterance, all 3696 utterances of workloads.timit_train_batch (stdseg 61 x 10, lam_for("cfg4")):
                          any contiguous rank shard / minibatch of bench.py is a slice of these
  cfg2/numer, cfg2/logZ   the same for the frame-level leg (61 labels, lam_for("cfg2"))
  cfg3/cost, cfg3/crc     per utterance of workloads.cfg3_batch(1680): float path cost and CRC-32 of the (label, duration, phone)
                          segment arrays of the best path (3-state Viterbi, lam_for("cfg3"))

  cfg5/numer, cfg5/logZ,  the stress leg (1024 phones, maxDur 30, 2000-frame utterances): the first CFG5_PINNED utterances of
  cfg5/cost, cfg5/crc     workloads.cfg5_batch() -- training numerator / logZ and the Viterbi path CRC / cost (the whole batch is hours of CPU)
  recipe/numer, .../logZ, the production TIMIT recipe's shape (48 phones, maxDur 10, 1162 state + 1872 transition features, joined streams):
  recipe/cost, recipe/crc the first RECIPE_PINNED utterances of workloads.recipe_batch(), training and Viterbi

Utterances are independent given lambda, so the work is cut into chunks that run in forked worker processes.

    python tests/golden/make_bench_pins.py [n_procs] [cfg4|cfg2|cfg3 ...]      (cfg4: ~45 min on 6 cores, cfg2 / cfg3: minutes)
"""
import multiprocessing as mp
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import workloads  # noqa: E402
from oracle.binding import OracleLib, RefLib, have_ref, make_config  # noqa: E402

OUT = os.path.join(HERE, "bench_pins.npz")


def lib_and_kind():
    """PIN_LIB=port forces the C restatement (about twice as fast as the reference's own node objects; it is itself pinned to the
    reference -- to 1e-13 on every golden and on bench shard 0)."""
    if os.environ.get("PIN_LIB") != "port" and have_ref():
        return RefLib(), "reference"
    return OracleLib(), "port"


def path_crc(lab, dur, phn):
    return zlib.crc32(np.ascontiguousarray(phn, np.uint32).tobytes(),
                      zlib.crc32(np.ascontiguousarray(dur, np.uint32).tobytes(), zlib.crc32(np.ascontiguousarray(lab, np.uint32).tobytes())))


def train_chunk(args):
    name, first, n = args
    lib, _ = lib_and_kind()
    cfg = make_config(**getattr(workloads, name + "_kwargs")())
    lam = workloads.lam_for(name, lib.lambda_len(cfg))
    off, ftrs, labs = workloads.timit_train_batch(first, n)
    _, numer, logz = lib.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=1)
    return first, numer, logz


def vit_chunk(args):
    first, n = args
    lib, _ = lib_and_kind()
    cfg = make_config(**workloads.cfg3_kwargs())
    lam = workloads.lam_for("cfg3", lib.lambda_len(cfg))
    off, ftrs = workloads.cfg3_batch(1680)
    sub_off = (off[first:first + n + 1] - off[first]).astype(np.uint32)
    segs, cost, _ = lib.viterbi(cfg, lam, sub_off, ftrs[int(off[first]):int(off[first + n])])
    return first, cost, np.array([path_crc(*s) for s in segs], np.uint32)


CFG5_PINNED = 4


def cfg5_chunk(u):
    lib, _ = lib_and_kind()
    cfg = make_config(**workloads.cfg5_kwargs())
    lam = workloads.lam_for("cfg5", lib.lambda_len(cfg))
    off, ftrs, labs = workloads.cfg5_batch()
    a, b = int(off[u]), int(off[u + 1])
    sub = np.array([0, b - a], np.uint32)
    _, numer, logz = lib.fwdbwd(cfg, lam, sub, ftrs[a:b], labs[a:b], n_threads=1)
    segs, cost, _ = lib.viterbi(cfg, lam, sub, ftrs[a:b])
    return u, numer[0], logz[0], cost[0], path_crc(*segs[0])


RECIPE_PINNED = 4


def recipe_chunk(u):
    lib, _ = lib_and_kind()
    cfg = make_config(**workloads.recipe_kwargs())
    lam = workloads.lam_for("recipe", lib.lambda_len(cfg))
    sub, f1, f2, labs = workloads.recipe_utt(*workloads.recipe_batch(), u)
    _, numer, logz = lib.fwdbwd(cfg, lam, sub, f1, labs, n_threads=1, ftrs2=f2)
    segs, cost, _ = lib.viterbi(cfg, lam, sub, f1, f2)
    return u, numer[0], logz[0], cost[0], path_crc(*segs[0])


def main():
    procs = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    what = sys.argv[2:] or ["cfg2", "cfg3", "cfg4"]
    out = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    _, kind = lib_and_kind()
    with mp.Pool(procs) as pool:
        for name in what:
            t0 = time.time()
            if name in ("cfg2", "cfg4"):
                step = 8 if name == "cfg4" else 64
                jobs = [(name, f, min(step, 3696 - f)) for f in range(0, 3696, step)]
                numer, logz = np.zeros(3696), np.zeros(3696)
                g0 = os.path.join(HERE, "cfg4_shard0_golden.npz")
                if name == "cfg4" and os.path.exists(g0):      # shard 0 = utterances 0..461 already came from the reference itself
                    z0 = np.load(g0)
                    numer[:462], logz[:462] = z0["numer"], z0["logZ"]
                    jobs = [j for j in jobs if j[1] >= 462 or j[1] + j[2] > 462]
                for k, (first, n, z) in enumerate(pool.imap_unordered(train_chunk, jobs)):
                    numer[first:first + len(n)] = n; logz[first:first + len(z)] = z
                    if k % 16 == 0:
                        print(f"{name}: {k + 1}/{len(jobs)} chunks, {time.time() - t0:.0f} s", flush=True)
                out[name + "/numer"], out[name + "/logZ"] = numer, logz
            elif name == "cfg3":
                jobs = [(f, min(24, 1680 - f)) for f in range(0, 1680, 24)]
                cost, crc = np.zeros(1680, np.float32), np.zeros(1680, np.uint32)
                for first, c, h in pool.imap_unordered(vit_chunk, jobs):
                    cost[first:first + len(c)] = c; crc[first:first + len(h)] = h
                out["cfg3/cost"], out["cfg3/crc"] = cost, crc
            elif name == "cfg5":
                numer, logz = np.zeros(CFG5_PINNED), np.zeros(CFG5_PINNED)
                cost, crc = np.zeros(CFG5_PINNED, np.float32), np.zeros(CFG5_PINNED, np.uint32)
                for u, n, z, c, h in pool.imap_unordered(cfg5_chunk, range(CFG5_PINNED)):
                    numer[u], logz[u], cost[u], crc[u] = n, z, c, h
                out["cfg5/numer"], out["cfg5/logZ"], out["cfg5/cost"], out["cfg5/crc"] = numer, logz, cost, crc
            elif name == "recipe":
                numer, logz = np.zeros(RECIPE_PINNED), np.zeros(RECIPE_PINNED)
                cost, crc = np.zeros(RECIPE_PINNED, np.float32), np.zeros(RECIPE_PINNED, np.uint32)
                for u, n, z, c, h in pool.imap_unordered(recipe_chunk, range(RECIPE_PINNED)):
                    numer[u], logz[u], cost[u], crc[u] = n, z, c, h
                out["recipe/numer"], out["recipe/logZ"], out["recipe/cost"], out["recipe/crc"] = numer, logz, cost, crc
            out[name + "/producer"] = np.array(kind)
            np.savez_compressed(OUT, **out)
            print(f"{name}: done in {time.time() - t0:.0f} s ({kind})", flush=True)


if __name__ == "__main__":
    main()
