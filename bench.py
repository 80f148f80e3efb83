#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on its own config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload = "cfg4"): segmental CRF `stdseg`, 61 phones x maxDur 10 (610 labels), 850 segment
features from 105 base features, TIMIT-shaped utterances (real lengths / segment boundaries, synthetic features
and labels).  One STEP = forward-backward + gradient over one minibatch of --utts-per-gpu (462 =
3696/8) utterances per GPU; rank r owns the contiguous utterance range [r*462,(r+1)*462) (the reference's
contiguous-view rule, CRF_FeatureStreamManager.cpp:425-464) and the lambda-gradient (+3 scalars) is combined
with ONE NCCL all-reduce per step.  `value` = frames of all ranks / max-over-ranks device time with the batch
already resident in HBM; `e2e` = the same metric through the host-buffer C-ABI call crfgpu_fwdbwd_batch with
pinned host inputs (H2D of base features + labels, D2H of gradient/numerator/logZ inside the timed region).
The JSON line also carries the Viterbi leg (cfg3: 183 labels, 1680 utterances) under "viterbi".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "asr-craft_b200"))

import workloads  # noqa: E402

METRIC = "seg-CRF fwd-bwd+grad frames/s (TIMIT shape)"
UNIT = "frames/s"


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


# ------------------------------------------------------------------------------------------------
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


def cpu_sample(cores, frames_per_utt):
    """Bounded sample of cfg4: the first `cores` utterances of the shard, truncated to frames_per_utt frames."""
    off, ftrs, labs = workloads.timit_train_batch(0, cores)
    keep = np.concatenate([np.arange(off[u], min(off[u] + frames_per_utt, off[u + 1])) for u in range(cores)])
    lens = [min(frames_per_utt, int(off[u + 1] - off[u])) for u in range(cores)]
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32), ftrs[keep], labs[keep]


def time_cpu(lib, cfg, lam, off, ftrs, labs, threads):
    t0 = time.perf_counter()
    lib.fwdbwd(cfg, lam, off, ftrs, labs, n_threads=threads)
    return time.perf_counter() - t0


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


# ------------------------------------------------------------------------------------------------
class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


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
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("CRFGPU_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = load_peaks()

    upg = args.utts_per_gpu
    off, ftrs, labs = workloads.timit_train_batch(rank * upg, upg)
    frames_local = int(off[-1])
    cfg = crf_b200.make_config(**workloads.cfg4_kwargs())
    m = crf_b200.CrfGpu(cfg, device=local)
    lam = workloads.lam_for("cfg4", m.lambda_len)
    m.set_lambda(lam)
    if args.slots:
        m.set_option("slots", args.slots)
    stream = torch.cuda.ExternalStream(m.stream, device=local)
    n_ext = m.lambda_len + 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce_grad():
        if world > 1:
            gptr, _, _ = m.device_results()
            g = torch.as_tensor(_DevArray(gptr, n_ext), device=f"cuda:{local}")
            with torch.cuda.stream(stream):
                dist.all_reduce(g)

    # ---- device-resident timing ("value") ----
    m.stage(off, ftrs, labs)
    for _ in range(args.warmup):
        m.fwdbwd_staged(); allreduce_grad()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = m.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase_names = ["score", "forward", "backward", "xi", "grad"]
    phase_acc = {k: 0.0 for k in phase_names}
    with torch.cuda.stream(stream):
        e0.record(stream)
    for _ in range(args.steps):
        m.fwdbwd_staged(); allreduce_grad()
    with torch.cuda.stream(stream):
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = m.launch_count - l0
    for k in phase_names:   # per-phase device time of the last step (CUDA events on the launching stream)
        phase_acc[k] = m.phase_ms(k)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    fr = torch.tensor([frames_local], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(fr)
    ms_max, frames_total = float(t.item()), float(fr.item())
    value = frames_total * args.steps / (ms_max / 1000.0)

    # ---- end-to-end through the host-buffer C ABI ("e2e") ----
    pin_f = crf_b200.PinnedBuffer(ftrs.shape, np.float32); pin_f.array[...] = ftrs
    pin_l = crf_b200.PinnedBuffer(labs.shape, np.uint32); pin_l.array[...] = labs
    pin_g = crf_b200.PinnedBuffer((m.lambda_len,), np.float64)
    pin_n = crf_b200.PinnedBuffer((upg,), np.float64); pin_z = crf_b200.PinnedBuffer((upg,), np.float64)
    out = (pin_g.array, pin_n.array, pin_z.array)
    e2e_steps = max(2, min(args.steps, 5))
    def e2e_step():
        # the host-buffer calls of one training step: H2D (crfgpu_stage_batch), kernels (crfgpu_fwdbwd_staged), the read-ahead of the
        # NEXT minibatch's features (crfgpu_prefetch_batch: its H2D + window expansion run on side streams under this step's kernels
        # and are taken over by the next crfgpu_stage_batch), ONE NCCL all-reduce of the gradient (+ scalars) on the handle's stream
        # when N > 1, D2H (crfgpu_fetch_fwdbwd).  Every step's H2D and D2H are inside the timed region.
        m.stage(off, pin_f.array, pin_l.array)
        m.fwdbwd_staged()
        if not args.no_prefetch:
            m.prefetch(off, pin_f.array)
        allreduce_grad()
        m.fetch_fwdbwd(out=out)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = frames_total * e2e_steps / float(te.item())
    h2d = int(ftrs.nbytes + labs.nbytes + off.nbytes + 3 * 4 * frames_local + 2 * 4 * frames_local)  # + derived per-frame index/label tables
    d2h = int(8 * (m.lambda_len + 2 * upg))
    ll = float(np.sum(pin_n.array - pin_z.array))

    # ---- Viterbi leg (cfg3), N=1 shard per rank, no collective ----
    vit = None
    if not args.no_viterbi:
        vcfg = crf_b200.make_config(**workloads.cfg3_kwargs())
        voff, vftrs = workloads.cfg3_batch(1680)
        per = 1680 // 8
        lo, hi = rank * per, (rank + 1) * per
        sub_off = (voff[lo:hi + 1] - voff[lo]).astype(np.uint32)
        sub_f = vftrs[int(voff[lo]):int(voff[hi])]
        vm = crf_b200.CrfGpu(vcfg, device=local)
        vm.set_lambda(workloads.lam_for("cfg3", vm.lambda_len))
        vstream = torch.cuda.ExternalStream(vm.stream, device=local)
        vm.stage(sub_off, sub_f)
        for _ in range(3):
            vm.viterbi_staged()
        barrier()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(vstream):
            v0.record(vstream)
        for _ in range(args.steps):
            vm.viterbi_staged()
        with torch.cuda.stream(vstream):
            v1.record(vstream)
        barrier()
        vms = v0.elapsed_time(v1)
        vt = torch.tensor([vms], dtype=torch.float64, device=f"cuda:{local}")
        vfr = torch.tensor([float(sub_off[-1])], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(vt, op=dist.ReduceOp.MAX); dist.all_reduce(vfr)
        vpin = crf_b200.PinnedBuffer(sub_f.shape, np.float32); vpin.array[...] = sub_f      # e2e: H2D from pinned host memory
        vm.viterbi(sub_off, vpin.array, raw=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            vm.viterbi(sub_off, vpin.array, raw=True)
        vdt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(vdt, op=dist.ReduceOp.MAX)
        ve2e = 3 * float(vfr.item()) / float(vdt.item())
        vpin.free()
        vit = {"metric": "Viterbi frames/s (cfg3: 183 labels = 61 phones x 3 states, 210 utterances per GPU)",
               "value": float(vfr.item()) * args.steps / (float(vt.item()) / 1000.0), "unit": UNIT, "e2e": ve2e,
               "score_ms": vm.phase_ms("viterbi_score"), "recursion_ms": vm.phase_ms("viterbi"),
               "frames_per_gpu": int(sub_off[-1])}
        vm.close()

    # ---- frame-level leg (cfg2: 61 labels, 105 features, the same TIMIT-shaped shard), no collective ----
    frame = None
    if not args.no_frame:
        fm = crf_b200.CrfGpu(crf_b200.make_config(**workloads.cfg2_kwargs()), device=local)
        fm.set_lambda(workloads.lam_for("cfg2", fm.lambda_len))
        fstream = torch.cuda.ExternalStream(fm.stream, device=local)
        fm.stage(off, ftrs, labs)
        for _ in range(3):
            fm.fwdbwd_staged()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(fstream):
            f0.record(fstream)
        for _ in range(args.steps):
            fm.fwdbwd_staged()
        with torch.cuda.stream(fstream):
            f1.record(fstream)
        barrier()
        ft = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(ft, op=dist.ReduceOp.MAX)
        fpin_g = crf_b200.PinnedBuffer((fm.lambda_len,), np.float64)
        fout = (fpin_g.array, pin_n.array, pin_z.array)
        for _ in range(2):
            fm.fwdbwd(off, pin_f.array, pin_l.array, out=fout)
        t0 = time.perf_counter()
        for _ in range(3):
            fm.fwdbwd(off, pin_f.array, pin_l.array, out=fout)
        fe2e = 3 * frames_total / (time.perf_counter() - t0)
        frame = {"metric": "frame-level CRF fwd-bwd+grad frames/s (cfg2: 61 labels, 105 features, 462 utterances per GPU)",
                 "value": frames_total * args.steps / (float(ft.item()) / 1e3), "unit": UNIT, "e2e": fe2e,
                 "phases_ms": {k: fm.phase_ms(k) for k in phase_names}, "lambda_len": fm.lambda_len}
        fpin_g.free()
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
        sm.stage(soff, sftrs)
        for _ in range(2):
            sm.viterbi_staged(); sm.synchronize()
        vs_, vr_ = sm.phase_ms("viterbi_score"), sm.phase_ms("viterbi")
        sN = float(soff[-1])
        stress = {"workload": "cfg5 (stdseg_no_dur_no_segtransftr, 1024 phones, maxDur 30, 542 segment features, 64 utterances x 2000 frames)",
                  "train_frames_per_s": sN / (sum(best.values()) / 1e3), "train_phases_ms": best,
                  "viterbi_frames_per_s": sN / ((vs_ + vr_) / 1e3), "viterbi_phases_ms": {"score": vs_, "recursion": vr_},
                  "lambda_len": sm.lambda_len}
        sm.close()

    if rank == 0:
        # per-frame ALGORITHMIC work of cfg4 (SURVEY.md 8d / DESIGN.md section 4).  Every phase sits below the tensor ridge
        # (1388 TFLOP/s / 6.55 TB/s = 212 flop/B): score and state-gradient GEMMs move 34 kB of window features per 1.04 Mflop
        # (29 flop/B), the transition-gradient GEMM 2*L^2 flop per 8*L bytes (152 flop/B), the lattice recursions are streams
        # of 4*L-byte frame vectors -> the bound of every kernel is HBM, and `achieved` is algorithmic bytes / measured time.
        L, P, D, Fs = 610, 61, 10, 850
        flops = {"score": 2.0 * (Fs + 1) * L, "forward": 2.0 * L * L, "backward": 2.0 * L * L, "xi": 2.0 * L * L,
                 "grad": 2.0 * L * (Fs + 1)}
        bytes_ = {"score": 4.0 * D * Fs + 4 * L, "forward": 8.0 * L, "backward": 12.0 * L, "xi": 8.0 * L, "grad": 4.0 * L + 4.0 * D * Fs}
        kernel_of = {"score": "score_gemm_tmem_kernel", "forward": "dp_tc_kernel<0>", "backward": "dp_tc_kernel<1>",
                     "xi": "frame_gemm_tmem_kernel<1>", "grad": "frame_gemm_tmem_kernel<0>"}
        launches_of = {"score": 1, "forward": 1, "backward": 1, "xi": 1, "grad": 1}
        dom = max(phase_names, key=lambda k: phase_acc[k])
        rooflines = {}
        for k in phase_names:
            sec = max(phase_acc[k], 1e-6) / 1000.0
            rooflines[k] = {"ms": phase_acc[k], "tflops": flops[k] * frames_local / sec / 1e12,
                            "gbs": bytes_[k] * frames_local / sec / 1e9, "kernel": kernel_of[k], "launches": launches_of[k]}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):      # DRAM read+write bytes of ONE launch from the committed ncu --set full capture
            traffic = json.load(open(tpath)).get(kernel_of[dom])
        ach = rooflines[dom]["gbs"]
        roof = {"kernel": kernel_of[dom], "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": traffic,
                "algorithmic_bytes_per_launch": bytes_[dom] * frames_local / launches_of[dom],
                "launch_ms": phase_acc[dom] / launches_of[dom], "launches_per_step": launches_of[dom],
                "peak_source": peaks["source"], "tensor_tflops": rooflines[dom]["tflops"]}
        # CPU baseline beside it (bounded sample, N=1 only)
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle.binding import make_config as omake
            lib, kind = cpu_baseline_lib()
            cores = cpu_threads()
            ocfg = omake(**workloads.cfg4_kwargs())
            fpu = 96
            coff, cftrs, clabs = cpu_sample(cores, fpu)
            sec = time_cpu(lib, ocfg, lam, coff, cftrs, clabs, cores)
            cpu = {"value": int(coff[-1]) / sec, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"{cores} utterances x first {fpu} frames ({int(coff[-1])} frames), {cores} pthreads, {sec:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg4", "model_type": "stdseg", "phones": 61, "max_dur": 10, "labels": 610, "base_ftrs": 105,
                       "seg_ftrs": 850, "lambda_len": m.lambda_len, "utts_per_gpu": upg, "frames_per_gpu": frames_local,
                       "sharding": "contiguous utterance ranges, one NCCL all-reduce of lambda_len+4 doubles per step",
                       "l2": "inputs_exceed_l2 (4.8 GB of window features + 5 x 0.36 GB lattice arrays per step vs 126 MB L2)",
                       "slots_per_cta_option": int(args.slots or 0)},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                      "steps": e2e_steps, "loglik_check": ll},
            "gpu_launches": int(launches), "roofline": roof, "phases": rooflines, "cpu_baseline": cpu, "viterbi": vit, "frame_crf": frame, "stress": stress}
        print(json.dumps(line))
    for pb in (pin_f, pin_l, pin_g, pin_n, pin_z):
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
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--no-viterbi", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-stress", action="store_true")
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
