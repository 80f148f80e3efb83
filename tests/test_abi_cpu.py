"""CPU tests of the drop-in boundary: libcrfgpu.so loads, exports every symbol include/crfgpu.h
declares, its host-only entry points (window width, label grouping) agree with the oracle, and the
compute entry points fail loudly -- never fall back -- when no CUDA device is usable."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import crf_b200
from helpers import load_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "crfgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crfgpu_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = crf_b200.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(crf_b200.SYMBOLS) == names


def test_config_struct_matches_oracle_layout():
    from oracle.binding import Config as OConfig
    assert [f[0] for f in OConfig._fields_] == [f[0] for f in crf_b200.Config._fields_]
    assert C.sizeof(OConfig) == C.sizeof(crf_b200.Config) == 15 * 4 + 4 + 16 + 8 * 4   # 15 words, pad, 2 doubles, 8 context / joined-stream words


def test_window_width():
    lib = crf_b200.load_library()
    for (F, D, seg, want) in [(105, 1, 0, 105), (105, 10, 1, 850), (64, 30, 1, 542), (9, 3, 0, 9)]:
        cfg = crf_b200.make_config("stdseg", n_labs=D * 2, n_base_ftrs=F, max_dur=D, extract_seg_ftrs=seg)
        assert lib.crfgpu_window_width(C.byref(cfg)) == want
    # the TIMIT recipe (demo/segmental-timit-demo.cfg.in:16-33): 1162 state + 1872 transition features
    cfg = crf_b200.make_config("stdseg_no_dur_no_segtransftr", n_labs=48, n_base_ftrs=144, max_dur=10, extract_seg_ftrs=1,
                               n_base_ftrs2=144, left_ctx2=6, right_ctx2=6)
    assert lib.crfgpu_window_width(C.byref(cfg)) == 1162 + 1872
    for kw, want in [(dict(max_dur=4, extract_seg_ftrs=1, left_ctx=2, right_ctx=3), 56), (dict(max_dur=3, left_ctx=3, right_ctx=1, boundary_delta=1), 8),
                     (dict(max_dur=1, left_ctx=2, right_ctx=2), 20)]:
        cfg = crf_b200.make_config("stdseg_no_dur_no_segtransftr" if kw["max_dur"] > 1 else "stdframe", n_labs=5, n_base_ftrs=4, **kw)
        assert lib.crfgpu_window_width(C.byref(cfg)) == want


@pytest.mark.parametrize("name", sorted(k for k in load_cases("window_golden.npz") if k.startswith("lab_")))
def test_group_labels_matches_reference_golden(name):
    c = load_cases("window_golden.npz")[name]
    lib = crf_b200.load_library()
    cfg = crf_b200.copy_config(c["cfg"])
    labs = np.ascontiguousarray(c["labs"], np.uint32)
    out = np.zeros((len(labs), 4), np.uint32)
    rc = lib.crfgpu_group_labels(C.byref(cfg), C.c_uint32(len(labs)), labs.ctypes.data_as(C.POINTER(C.c_uint32)),
                                 out.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert rc == 0
    assert np.array_equal(out, c["out"])


def test_no_gpu_means_loud_failure_not_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(crf_b200.CrfGpuError) as ei:
        crf_b200.CrfGpu(crf_b200.make_config("stdframe", n_labs=4, n_base_ftrs=3))
    assert ei.value.code == 3 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    """The package under asr-craft_b200/ must not reference oracle/ (checker only)."""
    pkg = os.path.join(ROOT, "asr-craft_b200")
    for dp, _, files in os.walk(pkg):
        if "build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "crf_oracle" not in text and "libcrforacle" not in text and "libcrfref" not in text, os.path.join(dp, f)


def test_balance_rules_of_one_global_minibatch():
    """crfgpu_balance_utts (equal counts, balanced frames) and crfgpu_balance_utts_cost (time model: step_frames x lock-steps + frames):
    both keep the membership of the minibatch; the time model gives the rank that holds the longest utterance the fewest frames and never
    a worse modelled step than the count-balanced deal."""
    import workloads
    utt_len, _, _ = workloads.timit_shape()
    for world in (2, 4, 8):
        lens = utt_len[:world * 462].astype(np.uint32)
        by_count = crf_b200.balance_utts(lens, world)
        by_cost = crf_b200.balance_utts_cost(lens, world, 240, 440.0)
        for ro in (by_count, by_cost):
            assert ro.min() == 0 and ro.max() == world - 1 and len(ro) == len(lens)
        assert all((by_count == r).sum() == 462 for r in range(world))

        def model(ro):
            return max(440.0 * max(int(lens[ro == r].max()), -(-int(lens[ro == r].sum()) // 240)) + int(lens[ro == r].sum()) for r in range(world))
        assert model(by_cost) <= model(by_count)
        frames = [int(lens[by_cost == r].sum()) for r in range(world)]
        longest_rank = int(by_cost[np.argmax(lens)])
        assert frames[longest_rank] == min(frames)
    # degenerate inputs
    assert list(crf_b200.balance_utts_cost(np.array([5, 9, 2], np.uint32), 1, 16, 100.0)) == [0, 0, 0]
    assert len(crf_b200.balance_utts_cost(np.zeros(0, np.uint32), 3, 16, 100.0)) == 0


def test_recipe_workload_and_pins():
    """the production-recipe leg of bench.py: geometry of the workload and the pins the leg is gated on (made by the reference itself,
    tests/golden/make_bench_pins.py recipe)"""
    import workloads
    kw = workloads.recipe_kwargs()
    assert kw["state_fidx"] == (0, 1161) and kw["trans_fidx"] == (1162, 3033)
    off, f1, f2, labs = workloads.recipe_batch(3)
    assert f1.shape == (int(off[-1]), 144) and f2.shape == (int(off[-1]) + 12 * 3, 144) and labs.max() < 48
    so, a1, a2, al = workloads.recipe_utt(off, f1, f2, labs, 1)
    assert a1.shape[0] == int(so[-1]) and a2.shape[0] == int(so[-1]) + 12 and np.array_equal(a1, f1[int(off[1]):int(off[2])])
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bench_pins.npz"))
    for k in ("recipe/numer", "recipe/logZ", "recipe/cost", "recipe/crc"):
        assert k in z and len(z[k]) == 4
