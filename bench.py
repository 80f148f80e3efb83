#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on its own config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload = "cfg4"): segmental CRF `stdseg`, 61 phones x maxDur 10 (610 labels), 850 segment
features from 105 base features, TIMIT-shaped utterances (real lengths / segment boundaries, synthetic features
and labels).  One STEP = forward-backward + gradient over one global minibatch of N x --utts-per-gpu
utterances (default 462 = 3696/8 per GPU: the whole corpus is one minibatch at N = 8): the global minibatch is the union
of the ranks' contiguous corpus views (CRF_FeatureStreamManager.cpp:425-464), its utterances are dealt to the ranks
length-balanced (crfgpu_balance_utts: membership of the minibatch unchanged, the gradient sum is order-free; --contiguous
keeps the reference's own placement) and the lambda-gradient (+3 scalars) is combined with ONE NCCL all-reduce per step
issued by the product (crfgpu_allreduce_grad).  `value` = frames of all ranks / max-over-ranks device time with the batch
already resident in HBM; `e2e` = the trainer's step through the C ABI with HOST buffers: crfgpu_stage_batch (H2D from pinned
memory) -> crfgpu_fwdbwd_staged -> crfgpu_prefetch_batch (next minibatch) -> crfgpu_allreduce_grad -> crfgpu_sgd_update
(lambda += lr * grad / N on the device, tables rebuilt) -> D2H of the numerators / logZ, every step inside the timed region.

PARITY GATE: before anything is timed every leg is checked against pins produced by the unmodified reference / the C
restatement (tests/golden/bench_pins.npz, cfg4_shard0_pin.json): per-utterance log-likelihood sums of the very minibatch
being timed (rel 1e-5), Viterbi path CRCs and float costs (bit-exact).  A mismatch prints the evidence and exits 3.

Other lines of the same JSON object: "minibatch_sweep" (64 and 148 utterances per GPU, SURVEY.md 8d's crf_bunch_size =
64 x nGPU), "viterbi" (cfg3: 183 labels, all 1680/N utterances per GPU) with its own roofline and CPU baseline,
"frame_crf" (cfg2), "stress" (cfg5, N = 1 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))

import workloads  # noqa: E402

METRIC = "seg-CRF fwd-bwd+grad frames/s (TIMIT shape)"
UNIT = "frames/s"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.rows, self.proc = dev, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ pins
class Pins:
    """Per-utterance results of the reference / the C restatement for the bench workloads (tests/golden/make_bench_pins.py)."""

    def __init__(self):
        self.z = {}
        p = os.path.join(GOLDEN, "bench_pins.npz")
        if os.path.exists(p):
            self.z = dict(np.load(p))
        g0 = os.path.join(GOLDEN, "cfg4_shard0_golden.npz")
        if "cfg4/logZ" not in self.z and os.path.exists(g0):      # at least bench shard 0, produced by the reference itself
            z0 = np.load(g0)
            numer, logz = np.zeros(3696), np.zeros(3696)
            numer[:462], logz[:462] = z0["numer"], z0["logZ"]
            self.z["cfg4/numer"], self.z["cfg4/logZ"] = numer, logz

    def loglik(self, name, ids):
        """(sum(numer - logZ) over the utterances, True) or (None, False) when a pin is missing for any of them"""
        k = name + "/logZ"
        if k not in self.z:
            return None, False
        ids = np.asarray(list(ids), np.int64)
        z, n = self.z[k][ids], self.z[name + "/numer"][ids]
        if np.any(z == 0.0):
            return None, False
        return float(np.sum(n - z)), True

    def check_loglik(self, name, ids, numer, logz, what, rtol=1e-5):
        got = float(np.sum(numer - logz))
        want, have = self.loglik(name, ids)
        rec = {"what": what, "loglik": got, "pinned": want, "ok": None}
        if have:
            z, n = self.z[name + "/logZ"][np.asarray(list(ids), np.int64)], self.z[name + "/numer"][np.asarray(list(ids), np.int64)]
            worst = float(np.max(np.abs(logz - z) / np.maximum(1.0, np.abs(z))))
            rec.update(ok=bool(abs(got - want) <= rtol * abs(want) and worst <= rtol and np.allclose(numer, n, rtol=rtol, atol=rtol)),
                       worst_rel_logZ=worst)
        return rec


def path_crc(lab, dur, phn):
    return zlib.crc32(np.ascontiguousarray(phn, np.uint32).tobytes(),
                      zlib.crc32(np.ascontiguousarray(dur, np.uint32).tobytes(), zlib.crc32(np.ascontiguousarray(lab, np.uint32).tobytes())))


def fail_parity(records):
    sys.stderr.write("bench.py: PARITY GATE FAILED -- nothing is reported for a path that computes the wrong answer\n")
    for r in records:
        sys.stderr.write(json.dumps(r) + "\n")
    sys.stderr.flush()
    os._exit(3)


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_lib():
    """The reference's own code when it was compiled here (oracle/_ref), else the C port."""
    from oracle.binding import OracleLib, RefLib, have_ref
    if have_ref():
        return RefLib(), "reference"
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return OracleLib(), "port"


def cpu_sample(n_utt, frames_per_utt):
    """Bounded sample of cfg4: the first n_utt utterances of the shard, truncated to frames_per_utt frames."""
    off, ftrs, labs = workloads.timit_train_batch(0, n_utt)
    keep = np.concatenate([np.arange(off[u], min(off[u] + frames_per_utt, off[u + 1])) for u in range(n_utt)])
    lens = [min(frames_per_utt, int(off[u + 1] - off[u])) for u in range(n_utt)]
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32), ftrs[keep], labs[keep]


def time_cpu(lib, cfg, lam, off, ftrs, labs, threads):
    t0 = time.perf_counter()
    lib.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=threads)
    return time.perf_counter() - t0


def cpu_recipe_baseline(threads, fpu=32):
    """The reference's buildGradient at the production recipe's shape on `threads` host threads: the first `threads` utterances of
    workloads.recipe_batch(), each truncated to its first fpu frames (stream 2 keeps its 6 + 6 context frames per utterance)."""
    from oracle.binding import make_config
    lib, kind = cpu_baseline_lib()
    cfg = make_config(**workloads.recipe_kwargs())
    lam = workloads.lam_for("recipe", lib.lambda_len(cfg))
    off, f1, f2, labs = workloads.recipe_batch(threads)
    ctx = 2 * workloads.RECIPE_CTX
    lens = [min(fpu, int(off[u + 1] - off[u])) for u in range(threads)]
    so = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    k1 = np.concatenate([np.arange(int(off[u]), int(off[u]) + lens[u]) for u in range(threads)])
    k2 = np.concatenate([np.arange(int(off[u]) + ctx * u, int(off[u]) + ctx * u + lens[u] + ctx) for u in range(threads)])
    a1, a2, al = np.ascontiguousarray(f1[k1]), np.ascontiguousarray(f2[k2]), np.ascontiguousarray(labs[k1])
    t0 = time.perf_counter()
    lib.fwdbwd(cfg, lam, so, a1, al, n_threads=threads, ftrs2=a2)
    dt = time.perf_counter() - t0
    return {"value": float(so[-1]) / dt, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{threads} utterances x first {fpu} frames ({int(so[-1])} frames), {threads} pthreads, {dt:.1f} s"}


_VIT_DATA = None      # (off, ftrs) of cfg3, set in the parent before the workers are forked (copy-on-write)


def _vit_worker(args):
    first, n, fpu = args
    from oracle.binding import make_config
    lib, _ = cpu_baseline_lib()
    cfg = make_config(**workloads.cfg3_kwargs())
    lam = workloads.lam_for("cfg3", lib.lambda_len(cfg))
    off, ftrs = _VIT_DATA
    keep = np.concatenate([np.arange(off[u], min(off[u] + fpu, off[u + 1])) for u in range(first, first + n)])
    lens = [min(fpu, int(off[u + 1] - off[u])) for u in range(first, first + n)]
    so = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    f = np.ascontiguousarray(ftrs[keep])
    t0 = time.perf_counter()
    lib.viterbi(cfg, lam, so, f)
    return int(so[-1]), time.perf_counter() - t0


def cpu_viterbi_baseline(cores, data, utts_per_core=4, fpu=300):
    """The reference decoder (nStateDecode, one decoder object per utterance, single-threaded as in CRFDecode) on `cores` forked
    processes side by side, a bounded sample of cfg3; value = frames of all processes / the slowest process's decode time."""
    import multiprocessing as mp
    global _VIT_DATA
    _VIT_DATA = data
    _, kind = cpu_baseline_lib()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_vit_worker, [(c * utts_per_core, utts_per_core, fpu) for c in range(cores)])
    frames, slowest = sum(r[0] for r in res), max(r[1] for r in res)
    return {"value": frames / slowest, "unit": UNIT, "cores": cores, "kind": kind,
            "one_thread": {"value": res[0][0] / res[0][1], "unit": UNIT, "cores": 1},
            "sample": f"{cores} processes x {utts_per_core} utterances x first {fpu} frames ({frames} frames), slowest process {slowest:.1f} s"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.binding import make_config
    lib, kind = cpu_baseline_lib()
    cores = cpu_threads()
    cfg = make_config(**workloads.cfg4_kwargs())
    lam = workloads.lam_for("cfg4", lib.lambda_len(cfg))
    fpu = 96
    off, ftrs, labs = cpu_sample(cores, fpu)
    frames = int(off[-1])
    for _ in range(args.warmup):
        time_cpu(lib, cfg, lam, off, ftrs, labs, cores)
    t = [time_cpu(lib, cfg, lam, off, ftrs, labs, cores) for _ in range(args.steps)]
    total = float(np.sum(t))
    value = frames * args.steps / total
    sample = f"{cores} utterances x first {fpu} frames each ({frames} frames/step), {cores} pthreads, reference sharding + serial reduce"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg4", "model_type": "stdseg", "phones": 61, "max_dur": 10, "labels": 610, "base_ftrs": 105,
                   "seg_ftrs": 850, "lambda_len": int(len(lam))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    import crf_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))      # plumbing: barriers, max-over-ranks, the 128-byte communicator id
    peaks = load_peaks()
    pins = Pins()
    gate = []          # parity records of this rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return float(t.item())

    def gather_list(x):
        if world == 1:
            return [x]
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    utt_len, _, _ = workloads.timit_shape()

    def my_utts(upg):
        """utterance ids of this rank for a global minibatch of world x upg utterances"""
        glob = np.arange(world * upg)                       # union of the ranks' next upg utterances = the first world*upg of the corpus views
        if world == 1 or args.contiguous:
            return glob[rank * upg:(rank + 1) * upg]
        rank_of = crf_b200.balance_utts(utt_len[glob].astype(np.uint32), world)
        return glob[rank_of == rank]

    cfg = crf_b200.make_config(**workloads.cfg4_kwargs())
    m = crf_b200.CrfGpu(cfg, device=local)
    lam = workloads.lam_for("cfg4", m.lambda_len)
    m.set_lambda(lam)
    if world > 1:
        # the product's own communicator: rank 0 creates the id, torch.distributed only ships its 128 bytes
        idt = torch.zeros(crf_b200.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt = torch.tensor(list(crf_b200.comm_unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(idt, 0)
        m.comm_init_rank(world, rank, bytes(idt.cpu().numpy().tobytes()))
    stream = torch.cuda.ExternalStream(m.stream, device=local)

    def allreduce_grad():
        if world > 1:
            m.allreduce_grad()

    def timed_resident(model, steps, step_fn):
        st = torch.cuda.ExternalStream(model.stream, device=local)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
        for _ in range(steps):
            step_fn()
        with torch.cuda.stream(st):
            e1.record(st)
        barrier()
        return reduce_max(e0.elapsed_time(e1))

    phase_names = ["score", "forward", "backward", "xi", "grad"]

    placement = {}

    def cfg4_leg(upg, steps, warmup, headline):
        ids = my_utts(upg)
        off, ftrs, labs = workloads.timit_train_utts(ids)
        frames_local = int(off[-1])
        if world > 1 and not args.contiguous and not args.count_balanced:
            # Placement by a time model (crfgpu_balance_utts_cost).  The lattice kernels are a dependent chain whose length is the LONGEST
            # utterance a rank holds (or its largest slot load), the rest of the step streams the rank's frames: a rank that holds one of
            # the corpus' longest utterances should get fewer frames.  The model's one constant -- the cost of a lock-step in units of
            # the per-frame cost -- is bracketed by this minibatch's own phase times (average cost per lock-step / per frame on the
            # count-balanced placement = an upper bound of the marginal ratio) and the candidates are TIMED: the fastest placement of
            # the global minibatch wins (membership, hence the gradient, is the same for all of them).
            def trial(ids_t):
                o, f, l = workloads.timit_train_utts(ids_t)
                m.stage(o, f, l)
                for _ in range(2):
                    m.fwdbwd_staged(); allreduce_grad()
                return timed_resident(m, 3, lambda: (m.fwdbwd_staged(), allreduce_grad())) / 3.0
            t_count = trial(ids)
            ph = {k: m.phase_ms(k) for k in phase_names}
            plan = m.plan_info()
            ls = int(plan.split("locksteps=")[1].split(";")[0])
            slots = 16 * int(plan.split(" clusters x 16 slots")[0].split()[-1]) if " clusters x 16 slots" in plan else 240
            upper = ((ph["forward"] + ph["backward"]) / max(ls, 1)) / ((ph["score"] + ph["xi"] + ph["grad"]) / max(frames_local, 1))
            upper = reduce_sum(upper) / world
            glob = np.arange(world * upg)
            best = (t_count, None, ids)
            tried = {"count_balanced": t_count}
            for frac in (0.3, 0.5, 0.7, 1.0):
                rank_of = crf_b200.balance_utts_cost(utt_len[glob].astype(np.uint32), world, slots, frac * upper)
                ids_c = glob[rank_of == rank]
                t_c = trial(ids_c)
                tried[f"step_frames={frac * upper:.0f}"] = t_c
                if t_c < best[0]:
                    best = (t_c, frac * upper, ids_c)
            ids = best[2]
            off, ftrs, labs = workloads.timit_train_utts(ids)
            frames_local = int(off[-1])
            placement[upg] = {"chosen": "count_balanced" if best[1] is None else f"step_frames={best[1]:.0f}", "slots": slots, "trial_ms_per_step": tried}
        # ---- parity gate: the minibatch about to be timed, through the host-buffer call ----
        _, n, z = m.fwdbwd(off, ftrs, labs)
        gate.append(pins.check_loglik("cfg4", ids, n, z, f"cfg4 fwd-bwd, {upg} utterances/GPU, rank {rank}"))
        if gate[-1]["ok"] is False:
            fail_parity(gate)
        m.stage(off, ftrs, labs)
        for _ in range(warmup):
            m.fwdbwd_staged(); allreduce_grad()
        sampler = ClockSampler(local)
        if rank == 0 and headline:
            sampler.start()
        l0 = m.launch_count
        ms = timed_resident(m, steps, lambda: (m.fwdbwd_staged(), allreduce_grad()))
        launches = m.launch_count - l0
        clocks = sampler.stop() if rank == 0 and headline else None
        phases = {k: m.phase_ms(k) for k in phase_names}
        frames_total = reduce_sum(frames_local)
        plans = gather_list(m.plan_info())
        fr = gather_list(frames_local)
        res = {"utts_per_gpu": upg, "value": frames_total * steps / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
               "frames_per_gpu": fr, "utts_on_gpu": gather_list(len(ids)), "locksteps_per_gpu": [int(p.split("locksteps=")[1].split(";")[0]) for p in plans],
               "phases_ms": phases, "loglik": gate[-1]["loglik"], "loglik_pinned": gate[-1]["pinned"], "parity_ok": gate[-1]["ok"]}
        return res, (ids, off, ftrs, labs, frames_local, frames_total, launches, clocks, plans)

    upg = args.utts_per_gpu
    head, (ids, off, ftrs, labs, frames_local, frames_total, launches, clocks, plans) = cfg4_leg(upg, args.steps, args.warmup, True)
    ms_per_step, value, phase_acc = head["ms_per_step"], head["value"], head["phases_ms"]

    # ---- end-to-end: the trainer's step through the C ABI with host buffers ----
    pin_f = crf_b200.PinnedBuffer(ftrs.shape, np.float32); pin_f.array[...] = ftrs
    pin_l = crf_b200.PinnedBuffer(labs.shape, np.uint32); pin_l.array[...] = labs
    pin_n = crf_b200.PinnedBuffer((len(ids),), np.float64); pin_z = crf_b200.PinnedBuffer((len(ids),), np.float64)
    e2e_steps = max(2, min(args.steps, 5))
    # a real update every step (update kernel + rebuild of every lambda-derived table); the rate is tiny so that the timed workload stays
    # the pinned one: one step moves the log-likelihood by lr * |grad|^2 / N = 1e-13 * 2.7e10 = 0.003 of -571 721
    e2e_lr = 1e-13

    def e2e_step():
        m.stage(off, pin_f.array, pin_l.array)                 # H2D of this minibatch (taken over from the read-ahead after the first step)
        m.fwdbwd_staged()
        if not args.no_prefetch:
            m.prefetch(off, pin_f.array, pin_l.array)          # read-ahead of the NEXT minibatch: H2D, window expansion and label tables on side streams
        allreduce_grad()                                       # ONE ncclAllReduce on the handle's stream (N > 1)
        m.fetch_fwdbwd(out=(None, pin_n.array, pin_z.array))   # D2H of the step's result: per-utterance numerators and logZ
        m.sgd_update(float(world), lr=e2e_lr)                  # lambda += lr * grad / nStreams_active on the device, every table rebuilt

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_dt = reduce_max(time.perf_counter() - t0)
    e2e_value = frames_total * e2e_steps / e2e_dt
    e2e_loglik = float(np.sum(pin_n.array - pin_z.array))
    gate.append(pins.check_loglik("cfg4", ids, pin_n.array, pin_z.array, f"cfg4 e2e last step, rank {rank}", rtol=1e-5))   # lambda moved by 7 updates of lr 1e-13
    if gate[-1]["ok"] is False:
        fail_parity(gate)
    h2d = int(ftrs.nbytes + labs.nbytes + off.nbytes + 2 * 4 * frames_local)  # + the label-derived per-frame tables (the index tables are built on the device)
    d2h = int(8 * 2 * len(ids))
    m.set_lambda(lam)

    # ---- minibatch sweep (crf_bunch_size = 64 x nGPU) ----
    sweep = []
    for u in [int(x) for x in args.sweep.split(",") if x]:
        if u == upg:
            continue
        r, _ = cfg4_leg(u, max(3, args.steps // 2), 3, False)
        sweep.append(r)

    # ---- Viterbi leg (cfg3), all 1680/N utterances per rank, no collective ----
    vit = None
    if not args.no_viterbi:
        vcfg = crf_b200.make_config(**workloads.cfg3_kwargs())
        voff, vftrs = workloads.cfg3_batch(1680)
        per = 1680 // world
        lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else 1680
        sub_off = (voff[lo:hi + 1] - voff[lo]).astype(np.uint32)
        sub_f = vftrs[int(voff[lo]):int(voff[hi])]
        vm = crf_b200.CrfGpu(vcfg, device=local)
        vm.set_lambda(workloads.lam_for("cfg3", vm.lambda_len))
        segs, cost = vm.viterbi(sub_off, sub_f)
        rec = {"what": f"cfg3 Viterbi, utterances {lo}..{hi - 1}, rank {rank}", "ok": None}
        if "cfg3/crc" in pins.z:
            crc = np.array([path_crc(*s) for s in segs], np.uint32)
            rec["paths_equal"] = int(np.sum(crc == pins.z["cfg3/crc"][lo:hi])); rec["n"] = hi - lo
            rec["costs_bit_equal"] = bool(np.array_equal(cost.view(np.uint32), pins.z["cfg3/cost"][lo:hi].view(np.uint32)))
            rec["ok"] = bool(rec["paths_equal"] == hi - lo and rec["costs_bit_equal"])
            rec["cost_sum"] = float(cost.astype(np.float64).sum())
        gate.append(rec)
        if rec["ok"] is False:
            fail_parity(gate)
        vm.stage(sub_off, sub_f)
        for _ in range(3):
            vm.viterbi_staged()
        vms = timed_resident(vm, args.steps, vm.viterbi_staged)
        vfr = reduce_sum(float(sub_off[-1]))
        vpin = crf_b200.PinnedBuffer(sub_f.shape, np.float32); vpin.array[...] = sub_f      # e2e: H2D from pinned host memory
        vm.viterbi(sub_off, vpin.array, raw=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            vm.viterbi(sub_off, vpin.array, raw=True)
        ve2e = 3 * vfr / reduce_max(time.perf_counter() - t0)
        vpin.free()
        sc_ms, rc_ms = vm.phase_ms("viterbi_score"), vm.phase_ms("viterbi")
        Lv, Fv, nfr = 183, 105, float(sub_off[-1])
        # algorithmic bytes per frame: scoring reads the 4F-byte frame and writes 4L scores; the recursion reads 4L scores,
        # writes a 2-byte back pointer + 1-byte duration per label and reads T of them back in the traceback
        vb = {"vit_scores_kernel": (4.0 * Fv + 4.0 * Lv, sc_ms), "viterbi_kernel": (4.0 * Lv + 3.0 * Lv + 3.0, rc_ms)}
        dom = max(vb, key=lambda k: vb[k][1])
        vroof = {"kernel": dom, "bound": "hbm", "achieved": vb[dom][0] * nfr / (vb[dom][1] / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "traffic": None, "launch_ms": vb[dom][1],
                 "note": "latency chain: one CTA per utterance walks its frames in order (778 dependent frames in the longest utterance); "
                         "the fp64 scoring kernel is bound by the DMUL+DADD rate the reference's summation order imposes"}
        vroof["frac"] = vroof["achieved"] / peaks["hbm_gbs"]
        vit = {"metric": "Viterbi frames/s (cfg3: 183 labels = 61 phones x 3 states, all 1680 utterances / N per GPU)",
               "value": vfr * args.steps / (vms / 1000.0), "unit": UNIT, "e2e": ve2e, "score_ms": sc_ms, "recursion_ms": rc_ms,
               "frames_per_gpu": int(sub_off[-1]), "utts_per_gpu": hi - lo, "parity": rec, "roofline": vroof}
        if rank == 0 and world == 1 and not args.no_cpu:
            vit["cpu_baseline"] = cpu_viterbi_baseline(cpu_threads(), (voff, vftrs))
        vm.close()

    # ---- frame-level leg (cfg2: 61 labels, 105 features, the same TIMIT-shaped minibatch), no collective ----
    frame = None
    if not args.no_frame:
        fm = crf_b200.CrfGpu(crf_b200.make_config(**workloads.cfg2_kwargs()), device=local)
        fm.set_lambda(workloads.lam_for("cfg2", fm.lambda_len))
        _, n2, z2 = fm.fwdbwd(off, ftrs, labs)
        gate.append(pins.check_loglik("cfg2", ids, n2, z2, f"cfg2 frame-level fwd-bwd, rank {rank}"))
        if gate[-1]["ok"] is False:
            fail_parity(gate)
        fm.stage(off, ftrs, labs)
        for _ in range(3):
            fm.fwdbwd_staged()
        fms = timed_resident(fm, args.steps, fm.fwdbwd_staged)
        pin_g = crf_b200.PinnedBuffer((fm.lambda_len,), np.float64)
        fout = (pin_g.array, pin_n.array, pin_z.array)
        def f_step():      # the trainer's loop as in the headline leg: this minibatch taken over from the read-ahead, the next one requested
            fm.stage(off, pin_f.array, pin_l.array)
            fm.fwdbwd_staged()
            if not args.no_prefetch:
                fm.prefetch(off, pin_f.array, pin_l.array)
            fm.fetch_fwdbwd(out=fout)

        for _ in range(2):
            f_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            f_step()
        torch.cuda.synchronize()
        fe2e = 5 * frames_total / reduce_max(time.perf_counter() - t0)
        gate.append(pins.check_loglik("cfg2", ids, pin_n.array, pin_z.array, f"cfg2 frame-level e2e last step, rank {rank}"))
        if gate[-1]["ok"] is False:
            fail_parity(gate)
        frame = {"metric": f"frame-level CRF fwd-bwd+grad frames/s (cfg2: 61 labels, 105 features, {upg} utterances per GPU)",
                 "value": frames_total * args.steps / (fms / 1e3), "unit": UNIT, "e2e": fe2e,
                 "phases_ms": {k: fm.phase_ms(k) for k in phase_names}, "lambda_len": fm.lambda_len, "parity": gate[-1]}
        pin_g.free()
        fm.close()

    # ---- stress leg (cfg5: 1024 phones, maxDur 30, 64 utterances x 2000 frames), one GPU, reported beside the headline ----
    stress = None
    if world == 1 and not args.no_stress:
        soff, sftrs, slabs = workloads.cfg5_batch()
        sm = crf_b200.CrfGpu(crf_b200.make_config(**workloads.cfg5_kwargs()), device=local)
        sm.set_lambda(workloads.lam_for("cfg5", sm.lambda_len))
        sm.stage(soff, sftrs, slabs)
        best = None
        for _ in range(3):
            sm.fwdbwd_staged(); sm.synchronize()
            ph = {k: sm.phase_ms(k) for k in phase_names}
            best = ph if best is None or sum(ph.values()) < sum(best.values()) else best
        # parity of the stress leg: the pinned utterances (the first few of the batch; the whole batch is hours of CPU) against the oracle
        _, sn, sz = sm.fetch_fwdbwd()
        npin = len(pins.z["cfg5/logZ"]) if "cfg5/logZ" in pins.z else 0
        if npin:
            gate.append(pins.check_loglik("cfg5", range(npin), sn[:npin], sz[:npin], f"cfg5 fwd-bwd, pinned utterances 0..{npin - 1} of 64"))
            if gate[-1]["ok"] is False:
                fail_parity(gate)
        sm.stage(soff, sftrs)
        for _ in range(2):
            sm.viterbi_staged(); sm.synchronize()
        vs_, vr_ = sm.phase_ms("viterbi_score"), sm.phase_ms("viterbi")
        if npin and "cfg5/crc" in pins.z:
            ssegs, scost = sm.fetch_viterbi(soff)
            crc = np.array([path_crc(*sg) for sg in ssegs[:npin]], np.uint32)
            rec = {"what": f"cfg5 Viterbi, pinned utterances 0..{npin - 1} of 64", "paths_equal": int(np.sum(crc == pins.z["cfg5/crc"])), "n": npin,
                   "costs_bit_equal": bool(np.array_equal(scost[:npin].view(np.uint32), pins.z["cfg5/cost"].view(np.uint32)))}
            rec["ok"] = bool(rec["paths_equal"] == npin and rec["costs_bit_equal"])
            gate.append(rec)
            if not rec["ok"]:
                fail_parity(gate)
        sN = float(soff[-1])
        stress = {"workload": "cfg5 (stdseg_no_dur_no_segtransftr, 1024 phones, maxDur 30, 542 segment features, 64 utterances x 2000 frames)",
                  "train_frames_per_s": sN / (sum(best.values()) / 1e3), "train_phases_ms": best,
                  "viterbi_frames_per_s": sN / ((vs_ + vr_) / 1e3), "viterbi_phases_ms": {"score": vs_, "recursion": vr_},
                  "lambda_len": sm.lambda_len, "plan": sm.plan_info()}
        sm.close()

    # ---- production-recipe leg (stdseg_no_dur_no_segtransftr + stdtrans at the TIMIT demo's shape), one GPU ----
    recipe = None
    if world == 1 and not args.no_recipe:
        roff, rf1, rf2, rlabs = workloads.recipe_batch(args.recipe_utts)      # the headline's minibatch shape by default: the first 462 TIMIT-shaped utterances
        rm = crf_b200.CrfGpu(crf_b200.make_config(**workloads.recipe_kwargs()), device=local)
        rm.set_lambda(workloads.lam_for("recipe", rm.lambda_len))
        rm.stage(roff, rf1, rlabs, ftrs2=rf2)
        rbest = None
        for _ in range(3):
            rm.fwdbwd_staged(); rm.synchronize()
            ph = {k: rm.phase_ms(k) for k in phase_names}
            rbest = ph if rbest is None or sum(ph.values()) < sum(rbest.values()) else rbest
        _, rn, rz = rm.fetch_fwdbwd()
        npin = len(pins.z["recipe/logZ"]) if "recipe/logZ" in pins.z else 0
        if npin:
            gate.append(pins.check_loglik("recipe", range(npin), rn[:npin], rz[:npin], f"recipe fwd-bwd, pinned utterances 0..{npin - 1} of {len(roff) - 1}"))
            if gate[-1]["ok"] is False:
                fail_parity(gate)
        # end to end: the trainer's step with host buffers -- both streams over PCIe, joined windows built on the device, fwd-bwd + gradient,
        # D2H of numerators / logZ, lambda update on the device, read-ahead of the next minibatch (crfgpu_prefetch_train_batch2)
        rp1 = crf_b200.PinnedBuffer(rf1.shape, np.float32); rp1.array[...] = rf1
        rp2 = crf_b200.PinnedBuffer(rf2.shape, np.float32); rp2.array[...] = rf2
        rpl = crf_b200.PinnedBuffer(rlabs.shape, np.uint32); rpl.array[...] = rlabs
        rpn = crf_b200.PinnedBuffer((len(roff) - 1,), np.float64); rpz = crf_b200.PinnedBuffer((len(roff) - 1,), np.float64)

        def r_step():
            rm.stage(roff, rp1.array, rpl.array, ftrs2=rp2.array)     # taken over from the read-ahead after the first step
            rm.fwdbwd_staged()
            if not args.no_prefetch:
                rm.prefetch(roff, rp1.array, rpl.array, ftrs2=rp2.array)   # the NEXT minibatch: copies, joined windows, label tables on the side stream
            rm.fetch_fwdbwd(out=(None, rpn.array, rpz.array))
            rm.sgd_update(1.0, lr=1e-16)      # a real update (every lambda-derived table rebuilt); the rate keeps the timed workload the pinned one

        for _ in range(2):
            r_step()
        t0 = time.perf_counter()
        for _ in range(3):
            r_step()
        torch.cuda.synchronize()
        re2e = 3 * float(roff[-1]) / (time.perf_counter() - t0)
        if npin:
            gate.append(pins.check_loglik("recipe", range(npin), rpn.array[:npin], rpz.array[:npin], "recipe e2e last step, pinned utterances", rtol=1e-5))
            if gate[-1]["ok"] is False:
                fail_parity(gate)
        rh2d = int(rf1.nbytes + rf2.nbytes + rlabs.nbytes + roff.nbytes + 3 * 4 * int(roff[-1]))
        for pb in (rp1, rp2, rpl, rpn, rpz):
            pb.free()
        rm.set_lambda(workloads.lam_for("recipe", rm.lambda_len))
        rm.stage(roff, rf1, ftrs2=rf2)
        for _ in range(2):
            rm.viterbi_staged(); rm.synchronize()
        rvs, rvr = rm.phase_ms("viterbi_score"), rm.phase_ms("viterbi")
        if npin and "recipe/crc" in pins.z:
            rsegs, rcost = rm.fetch_viterbi(roff)
            crc = np.array([path_crc(*sg) for sg in rsegs[:npin]], np.uint32)
            rec = {"what": f"recipe Viterbi, pinned utterances 0..{npin - 1} of {len(roff) - 1}", "paths_equal": int(np.sum(crc == pins.z["recipe/crc"])), "n": npin,
                   "costs_bit_equal": bool(np.array_equal(rcost[:npin].view(np.uint32), pins.z["recipe/cost"].view(np.uint32)))}
            rec["ok"] = bool(rec["paths_equal"] == npin and rec["costs_bit_equal"])
            gate.append(rec)
            if not rec["ok"]:
                fail_parity(gate)
        rN = float(roff[-1])
        recipe = {"workload": "TIMIT recipe shape (stdseg_no_dur_no_segtransftr + stdtrans, 48 phones, maxDur 10, 1162 state features from stream 1, "
                              f"1872 transition features from 13 context frames of stream 2; {len(roff) - 1} utterances = {int(rN)} frames, device-resident)",
                  "train_frames_per_s": rN / (sum(rbest.values()) / 1e3), "train_phases_ms": rbest,
                  "train_e2e_frames_per_s": re2e, "h2d_bytes_per_step": rh2d,
                  "viterbi_frames_per_s": rN / ((rvs + rvr) / 1e3), "viterbi_phases_ms": {"score": rvs, "recursion": rvr},
                  "lambda_len": rm.lambda_len, "plan": rm.plan_info()}
        # the leg's dominant kernel is tensor-bound ("the only genuinely tensor-bound GEMM"): the transition-gradient
        # reduce-GEMM [P^2 x frames] . [frames x (Ft + 1)], every product three bf16 MMAs (hi*hi + hi*lo + lo*hi)
        rP, rFt = workloads.RECIPE_PHONES, (2 * workloads.RECIPE_CTX + 1) * workloads.RECIPE_FTRS
        xi_flop = 2.0 * rN * rP * rP * (rFt + 1)
        xi_ms = rbest["xi"]
        recipe["roofline"] = {"kernel": "reduce_gemm_tiled_kernel (+ its two tiling passes)", "bound": "tensor", "achieved": 3.0 * xi_flop / (xi_ms / 1e3) / 1e12,
                              "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "traffic": None, "launch_ms": xi_ms,
                              "fp32_equivalent_tflops": xi_flop / (xi_ms / 1e3) / 1e12, "peak_source": peaks["source"],
                              "note": "achieved counts the three bf16 MMAs of every fp32-equivalent product; operands split into bf16 hi / lo and tiled "
                                      "once per batch, ring stages by bulk copies (DESIGN.md section 4); launch_ms is the whole xi phase"}
        recipe["roofline"]["frac"] = recipe["roofline"]["achieved"] / peaks["bf16_tflops_sustained"]
        rm.close()
        if rank == 0 and not args.no_cpu:
            recipe["cpu_baseline"] = cpu_recipe_baseline(cpu_threads())

    all_gate = [r for g in gather_list(gate) for r in g]
    if rank == 0:
        # per-frame ALGORITHMIC work of cfg4 (SURVEY.md 8d / DESIGN.md section 4).  Every phase sits below the tensor ridge
        # (1388 TFLOP/s / 6.55 TB/s = 212 flop/B): score and state-gradient GEMMs move 15.7 kB of window data per 1.04 Mflop
        # (66 flop/B), the transition-gradient GEMM 2*L^2 flop per 8*L bytes (152 flop/B), the lattice recursions are streams
        # of 4*L-byte frame vectors -> the bound of every kernel is HBM, and `achieved` is algorithmic bytes / measured time.
        L, P, D, Fs, F = 610, 61, 10, 850, 105
        flops = {"score": 2.0 * (Fs + 1) * L, "forward": 2.0 * L * L, "backward": 2.0 * L * L, "xi": 2.0 * L * L,
                 "grad": 2.0 * L * (Fs + 1)}
        # window GEMMs over the VIRTUAL windows: per frame D rows of the aggregate array (avg | max | min, 3F + 1
        # floats padded to 32) + one row of the padded base stream (the five sampled-frame blocks are row shifts of it) + the label vector
        win = 4.0 * D * ((3 * F + 1 + 31) // 32 * 32) + 4.0 * ((F + 31) // 32 * 32)
        bytes_ = {"score": win + 4 * L, "forward": 8.0 * L, "backward": 12.0 * L, "xi": 8.0 * L, "grad": 4.0 * L + win}
        lat = "dp_ks_kernel" if "dp_ks_kernel" in plans[0] else ("dp_tc_kernel" if "dp_tc_kernel" in plans[0] else "lattice kernel")
        kernel_of = {"score": "score_gemm_tmem_kernel", "forward": lat + "<0>", "backward": lat + "<1>",
                     "xi": "frame_gemm_tmem_kernel<1>", "grad": "frame_gemm_tmem_kernel<0>"}
        dom = max(phase_names, key=lambda k: phase_acc[k])
        rooflines = {}
        for k in phase_names:
            sec = max(phase_acc[k], 1e-6) / 1000.0
            rooflines[k] = {"ms": phase_acc[k], "tflops": flops[k] * frames_local / sec / 1e12,
                            "gbs": bytes_[k] * frames_local / sec / 1e9, "kernel": kernel_of[k], "launches": 1}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):      # DRAM read+write bytes of ONE launch from the committed ncu --set full capture
            traffic = json.load(open(tpath)).get(kernel_of[dom])
        ach = rooflines[dom]["gbs"]
        roof = {"kernel": kernel_of[dom], "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic,
                "algorithmic_bytes_per_launch": bytes_[dom] * frames_local,
                "launch_ms": phase_acc[dom], "launches_per_step": 1,
                "peak_source": peaks["source"], "tensor_tflops": rooflines[dom]["tflops"],
                "note": "a dependent chain of lock-steps (locksteps_per_gpu), not a stream: latency-bound, see DESIGN.md section 4"}
        # CPU baseline beside it (bounded sample, N=1 only): the reference's own pthread fan-out on all host threads, and one thread
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle.binding import make_config as omake
            lib, kind = cpu_baseline_lib()
            cores = cpu_threads()
            ocfg = omake(**workloads.cfg4_kwargs())
            fpu = 96
            coff, cftrs, clabs = cpu_sample(cores, fpu)
            sec = time_cpu(lib, ocfg, lam, coff, cftrs, clabs, cores)
            o1, f1, l1 = cpu_sample(1, 160)
            sec1 = time_cpu(lib, ocfg, lam, o1, f1, l1, 1)
            cpu = {"value": int(coff[-1]) / sec, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"{cores} utterances x first {fpu} frames ({int(coff[-1])} frames), {cores} pthreads, {sec:.1f} s",
                   "one_thread": {"value": int(o1[-1]) / sec1, "unit": UNIT, "cores": 1, "sample": f"1 utterance x first {int(o1[-1])} frames, {sec1:.1f} s"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3->f32 (every GEMM-shaped product is three bf16 tcgen05 MMAs on hi/lo splits with fp32 accumulation, ~16 mantissa bits; "
                     "log scales, logZ and the gradient sums in f64)",
            "data": "synthetic",
            "config": {"workload": "cfg4", "model_type": "stdseg", "phones": 61, "max_dur": 10, "labels": 610, "base_ftrs": 105,
                       "seg_ftrs": 850, "lambda_len": m.lambda_len, "utts_per_gpu": upg, "global_minibatch": upg * world,
                       "frames_per_gpu": head["frames_per_gpu"], "utts_on_gpu": head["utts_on_gpu"], "locksteps_per_gpu": head["locksteps_per_gpu"],
                       "placement_model": placement.get(upg),
                       "sharding": ("contiguous corpus views" if (args.contiguous or world == 1) else
                                    "global minibatch = union of the ranks' contiguous views, dealt to ranks " +
                                    ("with equal counts and balanced frame totals (crfgpu_balance_utts)" if args.count_balanced else
                                     "by the time model step_frames x lock-steps + frames, its constant measured on this minibatch (crfgpu_balance_utts_cost)"))
                                   + "; one ncclAllReduce of lambda_len+4 doubles per step (crfgpu_allreduce_grad)",
                       "l2": "inputs_exceed_l2 (1.8 GB of window aggregates + 5 x 0.36 GB lattice arrays per step vs 126 MB L2)",
                       "plan": plans[0], "nccl_ranks": m.comm_size if world > 1 else 1},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "step": "stage(H2D) + fwdbwd + prefetch(next) + allreduce + D2H(numer, logZ) + sgd_update(lr 1e-13) through the C ABI",
                    "loglik_check": e2e_loglik},
            "parity_gate": {"ok": all(r["ok"] is not False for r in all_gate), "unpinned": [r["what"] for r in all_gate if r["ok"] is None],
                            "cfg4_loglik": head["loglik"], "cfg4_loglik_pinned": head["loglik_pinned"], "records": all_gate if world == 1 else all_gate[:8]},
            "gpu_launches": int(launches), "roofline": roof, "phases": rooflines, "cpu_baseline": cpu, "minibatch_sweep": sweep,
            "viterbi": vit, "frame_crf": frame, "stress": stress, "recipe": recipe}
        print(json.dumps(line))
    for pb in (pin_f, pin_l, pin_n, pin_z):
        pb.free()
    m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts-per-gpu", type=int, default=462)
    ap.add_argument("--sweep", default="64,148", help="further minibatch sizes (utterances per GPU) reported under minibatch_sweep")
    ap.add_argument("--contiguous", action="store_true", help="keep the reference's contiguous placement instead of the length-balanced one")
    ap.add_argument("--count-balanced", action="store_true", help="N > 1: equal utterance counts per rank with balanced frame totals (crfgpu_balance_utts) instead of the time-model placement (crfgpu_balance_utts_cost)")
    ap.add_argument("--no-viterbi", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-stress", action="store_true")
    ap.add_argument("--no-recipe", action="store_true", help="skip the production-recipe leg (transition features, joined streams)")
    ap.add_argument("--recipe-utts", type=int, default=462, help="utterances of the production-recipe leg's minibatch")
    ap.add_argument("--no-frame", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
