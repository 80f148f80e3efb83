"""crf_b200 -- thin Python host layer over libcrfgpu.so (the C ABI declared in include/crfgpu.h).

The product path is the CUDA library; this module only marshals numpy / pinned buffers through ctypes.
Importing it never touches oracle/ and there is no CPU fallback: if libcrfgpu.so is missing, or no CUDA
device is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libcrfgpu.so")

MODEL_TYPES = {"stdframe": 0, "stdseg": 1, "stdseg_no_dur": 2,
               "stdseg_no_dur_no_transftr": 3, "stdseg_no_dur_no_segtransftr": 4}
LAB_BAD = 0xFFFFFFFF
ERR_NAMES = {1: "ARG", 2: "UNSUPPORTED", 3: "CUDA", 4: "NUMERIC"}


class CrfGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"crfgpu error {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Config(C.Structure):
    """crfgpu_config (include/crfgpu.h), mirrors CRF_FeatureMap_config + CRF_Model geometry."""
    _fields_ = [("model_type", C.c_uint32), ("n_labs", C.c_uint32), ("n_base_ftrs", C.c_uint32),
                ("n_states", C.c_uint32), ("max_dur", C.c_uint32), ("n_actual_labs", C.c_uint32),
                ("extract_seg_ftrs", C.c_uint32),
                ("use_state_ftrs", C.c_uint32), ("state_fidx_start", C.c_uint32), ("state_fidx_end", C.c_uint32),
                ("use_trans_ftrs", C.c_uint32), ("trans_fidx_start", C.c_uint32), ("trans_fidx_end", C.c_uint32),
                ("use_state_bias", C.c_uint32), ("use_trans_bias", C.c_uint32),
                ("state_bias_val", C.c_double), ("trans_bias_val", C.c_double),
                # context frames of stream 1, optional joined second stream (ftr1_* / ftr2_* of CRFTrain)
                ("left_ctx", C.c_uint32), ("right_ctx", C.c_uint32), ("boundary_delta", C.c_uint32),
                ("n_base_ftrs2", C.c_uint32), ("extract_seg_ftrs2", C.c_uint32), ("left_ctx2", C.c_uint32),
                ("right_ctx2", C.c_uint32), ("boundary_delta2", C.c_uint32)]


def window_width(n_base_ftrs, max_dur, extract_seg_ftrs, left_ctx=0, right_ctx=0, boundary_delta=0):
    """width of one stream's window vector (CRF_InFtrStream_SeqMultiWindow ctor, .cpp:47-117)"""
    if max_dur == 1:
        return (left_ctx + 1 + right_ctx) * n_base_ftrs
    if extract_seg_ftrs:
        return 8 * n_base_ftrs + max_dur + (left_ctx + right_ctx) * n_base_ftrs
    if boundary_delta:
        return min(left_ctx, right_ctx + 1) * n_base_ftrs
    return (left_ctx + 1 + right_ctx) * n_base_ftrs


def make_config(model_type="stdframe", n_labs=0, n_base_ftrs=0, n_states=1, max_dur=1, n_actual_labs=None,
                extract_seg_ftrs=0, use_trans_ftrs=0, state_fidx=None, trans_fidx=None,
                use_state_bias=1, use_trans_bias=1, state_bias_val=1.0, trans_bias_val=1.0,
                left_ctx=0, right_ctx=0, boundary_delta=0, n_base_ftrs2=0, extract_seg_ftrs2=0, left_ctx2=0, right_ctx2=0,
                boundary_delta2=0):
    """Same defaults as CRFTrain's set_fmap_config (CRFTrain/src/Main.cpp:372-430)."""
    w = window_width(n_base_ftrs, max_dur, extract_seg_ftrs, left_ctx, right_ctx, boundary_delta)
    if n_base_ftrs2:
        w += window_width(n_base_ftrs2, max_dur, extract_seg_ftrs2, left_ctx2, right_ctx2, boundary_delta2)
    if n_actual_labs is None:
        n_actual_labs = n_labs // max_dur if model_type == "stdseg" else n_labs
    s0, s1 = state_fidx if state_fidx is not None else (0, w - 1)
    t0, t1 = trans_fidx if trans_fidx is not None else (0, w - 1)
    return Config(MODEL_TYPES[model_type], n_labs, n_base_ftrs, n_states, max_dur, n_actual_labs,
                  int(extract_seg_ftrs), 1, s0, s1, int(use_trans_ftrs), t0, t1,
                  int(use_state_bias), int(use_trans_bias), state_bias_val, trans_bias_val,
                  int(left_ctx), int(right_ctx), int(boundary_delta), int(n_base_ftrs2), int(extract_seg_ftrs2),
                  int(left_ctx2), int(right_ctx2), int(boundary_delta2))


def copy_config(cfg):
    """Field-wise copy from any ctypes struct with the same field names (e.g. the oracle's Config)."""
    return Config(*[getattr(cfg, f[0]) for f in Config._fields_])


_lib = None

SYMBOLS = ["crfgpu_last_error", "crfgpu_create", "crfgpu_destroy", "crfgpu_window_width", "crfgpu_lambda_len",
           "crfgpu_index_maps", "crfgpu_set_lambda", "crfgpu_fwdbwd_batch", "crfgpu_viterbi_batch",
           "crfgpu_expand_windows", "crfgpu_group_labels", "crfgpu_stage_batch", "crfgpu_fwdbwd_staged",
           "crfgpu_viterbi_staged", "crfgpu_device_results", "crfgpu_fetch_fwdbwd", "crfgpu_fetch_viterbi",
           "crfgpu_synchronize", "crfgpu_stream", "crfgpu_launch_count", "crfgpu_phase_ms",
           "crfgpu_fetch_alpha_beta", "crfgpu_set_option", "crfgpu_host_alloc", "crfgpu_host_free",
           "crfgpu_sgd_update", "crfgpu_get_lambda", "crfgpu_set_train_state", "crfgpu_prefetch_batch",
           "crfgpu_comm_unique_id", "crfgpu_comm_init_rank", "crfgpu_comm_init_all", "crfgpu_comm_destroy", "crfgpu_comm_size",
           "crfgpu_group_start", "crfgpu_group_end", "crfgpu_allreduce_grad", "crfgpu_fetch_tail",
           "crfgpu_shard_views", "crfgpu_minibatch_share", "crfgpu_balance_utts", "crfgpu_plan_info",
           "crfgpu_fetch_posterior_mass", "crfgpu_stage_batch2", "crfgpu_fwdbwd_batch2", "crfgpu_viterbi_batch2",
           "crfgpu_expand_windows2", "crfgpu_prefetch_train_batch", "crfgpu_prefetch_train_batch2", "crfgpu_set_phone_lm", "crfgpu_set_phone_unigram_lm", "crfgpu_set_beam", "crfgpu_balance_utts_cost"]
COMM_ID_BYTES = 128


class Sgd(C.Structure):
    """crfgpu_sgd (include/crfgpu.h): the per-minibatch update of CRF_SGTrainer.cpp:299-325."""
    _fields_ = [("lr", C.c_double), ("use_gvar", C.c_uint32), ("inv_square_var", C.c_double),
                ("use_adagrad", C.c_uint32), ("eta", C.c_double), ("eps", C.c_double)]


def load_library(path=None):
    """Loads libcrfgpu.so (raises if absent -- there is no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found: build it with `make -C asr-craft_b200` (or __graft_entry__.build())")
    lib = C.CDLL(path)
    lib.crfgpu_last_error.restype = C.c_char_p
    lib.crfgpu_lambda_len.restype = C.c_uint32
    lib.crfgpu_window_width.restype = C.c_uint32
    lib.crfgpu_stream.restype = C.c_void_p
    lib.crfgpu_launch_count.restype = C.c_uint64
    lib.crfgpu_phase_ms.restype = C.c_double
    for name in ("crfgpu_destroy", "crfgpu_lambda_len", "crfgpu_fwdbwd_staged", "crfgpu_viterbi_staged",
                 "crfgpu_synchronize", "crfgpu_stream", "crfgpu_launch_count"):
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.crfgpu_phase_ms.argtypes = [C.c_void_p, C.c_char_p]
    lib.crfgpu_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    lib.crfgpu_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
    lib.crfgpu_host_free.argtypes = [C.c_void_p]
    lib.crfgpu_sgd_update.argtypes = [C.c_void_p, C.POINTER(Sgd), C.c_double]
    lib.crfgpu_get_lambda.argtypes = [C.c_void_p] + [C.POINTER(C.c_double)] * 4
    lib.crfgpu_set_train_state.argtypes = [C.c_void_p] + [C.POINTER(C.c_double)] * 3
    lib.crfgpu_comm_unique_id.argtypes = [C.c_void_p]
    lib.crfgpu_comm_init_rank.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.crfgpu_comm_init_all.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    for name in ("crfgpu_comm_destroy", "crfgpu_comm_size", "crfgpu_allreduce_grad"):
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.crfgpu_fetch_tail.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.crfgpu_shard_views.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.crfgpu_minibatch_share.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    lib.crfgpu_minibatch_share.restype = C.c_uint32
    lib.crfgpu_balance_utts.argtypes = [C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32)]
    lib.crfgpu_plan_info.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
    lib.crfgpu_plan_info.restype = C.c_uint32
    _lib = lib
    return lib


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class PinnedBuffer:
    """Page-locked host array (cudaMallocHost) exposed as numpy."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        rc = self.lib.crfgpu_host_alloc(C.byref(p), C.c_uint64(max(n, 1)))
        if rc:
            raise CrfGpuError(rc, self.lib.crfgpu_last_error().decode())
        self.ptr = p
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr is not None:
            self.array = None
            self.lib.crfgpu_host_free(self.ptr)
            self.ptr = None


def _host_check(rc):
    if rc:
        raise CrfGpuError(rc, load_library().crfgpu_last_error().decode())


def shard_views(n_utt, n_streams):
    """crfgpu_shard_views: the reference's contiguous corpus views (first, count) per stream."""
    first = np.zeros(n_streams, np.uint32); count = np.zeros(n_streams, np.uint32)
    _host_check(load_library().crfgpu_shard_views(n_utt, n_streams, _ptr(first, C.c_uint32), _ptr(count, C.c_uint32)))
    return first, count


def minibatch_share(minibatch, n_streams, stream):
    return int(load_library().crfgpu_minibatch_share(minibatch, n_streams, stream))


def balance_utts(n_frames, n_ranks):
    """crfgpu_balance_utts: rank of every utterance of one global minibatch (equal counts, frames balanced, longest first)."""
    n_frames = np.ascontiguousarray(n_frames, np.uint32)
    out = np.zeros(len(n_frames), np.uint32)
    _host_check(load_library().crfgpu_balance_utts(len(n_frames), _ptr(n_frames, C.c_uint32), n_ranks, _ptr(out, C.c_uint32)))
    return out


def balance_utts_cost(n_frames, n_ranks, n_slots, step_frames):
    """crfgpu_balance_utts_cost: rank of every utterance of one global minibatch by the time model step_frames * lock-steps + frames"""
    n_frames = np.ascontiguousarray(n_frames, np.uint32)
    out = np.zeros(len(n_frames), np.uint32)
    lib = load_library()
    lib.crfgpu_balance_utts_cost.argtypes = [C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.c_double, C.POINTER(C.c_uint32)]
    _host_check(lib.crfgpu_balance_utts_cost(len(n_frames), _ptr(n_frames, C.c_uint32), n_ranks, n_slots, float(step_frames), _ptr(out, C.c_uint32)))
    return out


def comm_unique_id():
    """128 bytes rank 0 hands to the other ranks (crfgpu_comm_unique_id)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _host_check(load_library().crfgpu_comm_unique_id(buf))
    return buf.raw


def comm_init_all(models):
    """One process, one handle per device: crfgpu_comm_init_all."""
    arr = (C.c_void_p * len(models))(*[m.h for m in models])
    _host_check(load_library().crfgpu_comm_init_all(arr, len(models)))


def group_start():
    _host_check(load_library().crfgpu_group_start())


def group_end():
    _host_check(load_library().crfgpu_group_end())


class CrfGpu:
    """One device handle: CRF_Model + CRF_FeatureMap + grad builder + decoder for one geometry."""

    def __init__(self, cfg, device=0):
        self.lib = load_library()
        self.cfg = copy_config(cfg)
        h = C.c_void_p()
        rc = self.lib.crfgpu_create(C.byref(self.cfg), C.c_int(device), C.byref(h))
        if rc:
            raise CrfGpuError(rc, self.lib.crfgpu_last_error().decode())
        self.h = h
        self.lambda_len = self.lib.crfgpu_lambda_len(self.h)
        self._n_utt = 0
        self._n_frames = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.crfgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise CrfGpuError(rc, self.lib.crfgpu_last_error().decode())

    def set_option(self, name, value):
        self._check(self.lib.crfgpu_set_option(self.h, name.encode(), int(value)))

    def index_maps(self):
        L = self.cfg.n_labs
        s = np.zeros(L, np.uint32)
        t = np.zeros((L, L), np.uint32)
        self._check(self.lib.crfgpu_index_maps(self.h, _ptr(s, C.c_uint32), _ptr(t, C.c_uint32)))
        return s, t

    def set_lambda(self, lam):
        lam = np.ascontiguousarray(lam, np.float64)
        self._check(self.lib.crfgpu_set_lambda(self.h, _ptr(lam, C.c_double), C.c_uint32(len(lam))))

    def sgd_update(self, n_active, lr=0.0, use_gvar=0, inv_square_var=0.0, use_adagrad=0, eta=0.0, eps=0.0):
        """lambda += lr * grad / n_active (or AdaGrad) on the device from the staged gradient; rebuilds the device tables."""
        opt = Sgd(lr, use_gvar, inv_square_var, use_adagrad, eta, eps)
        self._check(self.lib.crfgpu_sgd_update(self.h, C.byref(opt), C.c_double(n_active)))

    def get_lambda(self, with_state=False):
        n = self.lambda_len
        lam = np.zeros(n, np.float64)
        if not with_state:
            self._check(self.lib.crfgpu_get_lambda(self.h, _ptr(lam, C.c_double), None, None, None))
            return lam
        acc, sqr, gsq = np.zeros(n), np.zeros(n), np.zeros(n)
        self._check(self.lib.crfgpu_get_lambda(self.h, _ptr(lam, C.c_double), _ptr(acc, C.c_double), _ptr(sqr, C.c_double),
                                               _ptr(gsq, C.c_double)))
        return lam, acc, sqr, gsq

    def set_train_state(self, lambda_acc=None, lambda_sqr_acc=None, grad_sqr_acc=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (lambda_acc, lambda_sqr_acc, grad_sqr_acc)]
        self._check(self.lib.crfgpu_set_train_state(self.h, *[None if a is None else _ptr(a, C.c_double) for a in arrs]))

    def set_beam(self, beam):
        """crfgpu_set_beam: beam pruning of the decoder (0 = off)"""
        self._check(self.lib.crfgpu_set_beam(self.h, C.c_double(beam)))

    def set_phone_lm(self, start=None, bigram=None, final=None):
        """crfgpu_set_phone_lm: complete phone-bigram LM (costs) for decoding, one state per phone; no arguments drop it.
        With N > 1 states per phone the arguments are (unigram[P], exit[P], final[P]): crfgpu_set_phone_unigram_lm."""
        if start is None:
            self._check(self.lib.crfgpu_set_phone_lm(self.h, None, None, None))
            return
        a = [np.ascontiguousarray(x, np.float32) for x in (start, bigram, final)]
        fn = self.lib.crfgpu_set_phone_lm if self.cfg.n_states == 1 else self.lib.crfgpu_set_phone_unigram_lm
        self._check(fn(self.h, *[_ptr(x, C.c_float) for x in a]))

    # ---- host-buffer calls --------------------------------------------------------------------
    def fwdbwd(self, off, ftrs, labs, out=None, ftrs2=None):
        """ftrs2: the joined second stream (crfgpu_fwdbwd_batch2); with context frames stream s has
        off[u] + u * (left_ctx_s + right_ctx_s) rows before utterance u"""
        off = np.ascontiguousarray(off, np.uint32)
        ftrs = np.ascontiguousarray(ftrs, np.float32)
        labs = np.ascontiguousarray(labs, np.uint32)
        n = len(off) - 1
        if out is None:
            out = (np.zeros(self.lambda_len, np.float64), np.zeros(n, np.float64), np.zeros(n, np.float64))
        grad, numer, logz = out
        if ftrs2 is not None:
            ftrs2 = np.ascontiguousarray(ftrs2, np.float32)
            self._check(self.lib.crfgpu_fwdbwd_batch2(self.h, C.c_uint32(n), _ptr(off, C.c_uint32), _ptr(ftrs, C.c_float), _ptr(ftrs2, C.c_float),
                                                      _ptr(labs, C.c_uint32), _ptr(grad, C.c_double),
                                                      _ptr(numer, C.c_double), _ptr(logz, C.c_double)))
        else:
            self._check(self.lib.crfgpu_fwdbwd_batch(self.h, C.c_uint32(n), _ptr(off, C.c_uint32), _ptr(ftrs, C.c_float),
                                                     _ptr(labs, C.c_uint32), _ptr(grad, C.c_double),
                                                     _ptr(numer, C.c_double), _ptr(logz, C.c_double)))
        self._n_utt, self._n_frames = n, int(off[-1])
        return grad, numer, logz

    def viterbi(self, off, ftrs, raw=False, ftrs2=None):
        off = np.ascontiguousarray(off, np.uint32)
        ftrs = np.ascontiguousarray(ftrs, np.float32)
        f2 = None if ftrs2 is None else np.ascontiguousarray(ftrs2, np.float32)
        n, tot = len(off) - 1, int(off[-1])
        lab = np.zeros(tot, np.uint32)
        dur = np.zeros(tot, np.uint32)
        phn = np.zeros(tot, np.uint32)
        nseg = np.zeros(n, np.uint32)
        cost = np.zeros(n, np.float32)
        self._check(self.lib.crfgpu_viterbi_batch2(self.h, C.c_uint32(n), _ptr(off, C.c_uint32), _ptr(ftrs, C.c_float),
                                                   None if f2 is None else _ptr(f2, C.c_float),
                                                   _ptr(lab, C.c_uint32), _ptr(dur, C.c_uint32), _ptr(phn, C.c_uint32),
                                                   _ptr(nseg, C.c_uint32), _ptr(cost, C.c_float)))
        self._n_utt, self._n_frames = n, tot
        if raw:
            return lab, dur, phn, nseg, cost
        segs = []
        for u in range(n):
            b, k = int(off[u]), int(nseg[u])
            segs.append((lab[b:b + k].copy(), dur[b:b + k].copy(), phn[b:b + k].copy()))
        return segs, cost

    def expand_windows(self, ftrs, ftrs2=None):
        """ftrs: [left_ctx + T + right_ctx][n_base_ftrs]; ftrs2 likewise for the joined second stream -> [T][max_dur][width]"""
        ftrs = np.ascontiguousarray(ftrs, np.float32)
        f2 = None if ftrs2 is None else np.ascontiguousarray(ftrs2, np.float32)
        T = ftrs.shape[0] - self.cfg.left_ctx - self.cfg.right_ctx
        w = int(self.lib.crfgpu_window_width(C.byref(self.cfg)))
        out = np.zeros((T, self.cfg.max_dur, w), np.float32)
        self._check(self.lib.crfgpu_expand_windows2(self.h, C.c_uint32(T), _ptr(ftrs, C.c_float), None if f2 is None else _ptr(f2, C.c_float),
                                                    _ptr(out, C.c_float)))
        return out

    def group_labels(self, labs):
        labs = np.ascontiguousarray(labs, np.uint32)
        out = np.zeros((len(labs), 4), np.uint32)
        self._check(self.lib.crfgpu_group_labels(C.byref(self.cfg), C.c_uint32(len(labs)), _ptr(labs, C.c_uint32),
                                                 _ptr(out, C.c_uint32)))
        return out

    # ---- device-resident calls ----------------------------------------------------------------
    def stage(self, off, ftrs, labs=None, ftrs2=None):
        off = np.ascontiguousarray(off, np.uint32)
        ftrs = np.ascontiguousarray(ftrs, np.float32)
        f2 = None if ftrs2 is None else np.ascontiguousarray(ftrs2, np.float32)
        lp = None
        if labs is not None:
            labs = np.ascontiguousarray(labs, np.uint32)
            lp = _ptr(labs, C.c_uint32)
        n = len(off) - 1
        self._check(self.lib.crfgpu_stage_batch2(self.h, C.c_uint32(n), _ptr(off, C.c_uint32), _ptr(ftrs, C.c_float),
                                                 None if f2 is None else _ptr(f2, C.c_float), lp))
        self._n_utt, self._n_frames = n, int(off[-1])

    def prefetch(self, off, ftrs, labs=None, ftrs2=None):
        """crfgpu_prefetch_batch / crfgpu_prefetch_train_batch[2]: copy + window-expand the NEXT batch on side streams (with labs: its
        label tables too); `ftrs` (`ftrs2`, `labs`) must be the very arrays later staged."""
        off = np.ascontiguousarray(off, np.uint32)
        assert ftrs.dtype == np.float32 and ftrs.flags["C_CONTIGUOUS"]
        if ftrs2 is not None:
            assert ftrs2.dtype == np.float32 and ftrs2.flags["C_CONTIGUOUS"]
            assert labs is None or (labs.dtype == np.uint32 and labs.flags["C_CONTIGUOUS"])
            self._check(self.lib.crfgpu_prefetch_train_batch2(self.h, C.c_uint32(len(off) - 1), _ptr(off, C.c_uint32), _ptr(ftrs, C.c_float),
                                                              _ptr(ftrs2, C.c_float), None if labs is None else _ptr(labs, C.c_uint32)))
        elif labs is None:
            self._check(self.lib.crfgpu_prefetch_batch(self.h, C.c_uint32(len(off) - 1), _ptr(off, C.c_uint32), _ptr(ftrs, C.c_float)))
        else:
            assert labs.dtype == np.uint32 and labs.flags["C_CONTIGUOUS"]
            self._check(self.lib.crfgpu_prefetch_train_batch(self.h, C.c_uint32(len(off) - 1), _ptr(off, C.c_uint32), _ptr(ftrs, C.c_float),
                                                             _ptr(labs, C.c_uint32)))

    def fwdbwd_staged(self):
        self._check(self.lib.crfgpu_fwdbwd_staged(self.h))

    def viterbi_staged(self):
        self._check(self.lib.crfgpu_viterbi_staged(self.h))

    def fetch_viterbi(self, off):
        """crfgpu_fetch_viterbi after viterbi_staged(): (segments per utterance, path costs) like viterbi()"""
        off = np.ascontiguousarray(off, np.uint32)
        n, tot = len(off) - 1, int(off[-1])
        lab = np.zeros(tot, np.uint32); dur = np.zeros(tot, np.uint32); phn = np.zeros(tot, np.uint32)
        nseg = np.zeros(n, np.uint32); cost = np.zeros(n, np.float32)
        self._check(self.lib.crfgpu_fetch_viterbi(self.h, _ptr(lab, C.c_uint32), _ptr(dur, C.c_uint32), _ptr(phn, C.c_uint32),
                                                  _ptr(nseg, C.c_uint32), _ptr(cost, C.c_float)))
        segs = []
        for u in range(n):
            b, k = int(off[u]), int(nseg[u])
            segs.append((lab[b:b + k].copy(), dur[b:b + k].copy(), phn[b:b + k].copy()))
        return segs, cost

    def synchronize(self):
        self._check(self.lib.crfgpu_synchronize(self.h))

    def fetch_fwdbwd(self, out=None):
        if out is None:
            out = (np.zeros(self.lambda_len, np.float64), np.zeros(self._n_utt, np.float64),
                   np.zeros(self._n_utt, np.float64))
        grad, numer, logz = out          # any may be None: that array stays on the device
        self._check(self.lib.crfgpu_fetch_fwdbwd(self.h, *[None if a is None else _ptr(a, C.c_double) for a in (grad, numer, logz)]))
        return grad, numer, logz

    def device_results(self):
        g, n, z = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.lib.crfgpu_device_results(self.h, C.byref(g), C.byref(n), C.byref(z)))
        return g.value, n.value, z.value

    def fetch_alpha_beta(self):
        L = self.cfg.n_labs
        a = np.zeros((self._n_frames, L), np.float64)
        b = np.zeros((self._n_frames, L), np.float64)
        self._check(self.lib.crfgpu_fetch_alpha_beta(self.h, _ptr(a, C.c_double), _ptr(b, C.c_double)))
        return a, b

    # ---- multi-GPU ----------------------------------------------------------------------------
    def comm_init_rank(self, n_ranks, rank, id128):
        buf = C.create_string_buffer(bytes(id128), COMM_ID_BYTES)
        self._check(self.lib.crfgpu_comm_init_rank(self.h, n_ranks, rank, buf))

    def comm_destroy(self):
        self._check(self.lib.crfgpu_comm_destroy(self.h))

    @property
    def comm_size(self):
        return int(self.lib.crfgpu_comm_size(self.h))

    def allreduce_grad(self):
        """ONE ncclAllReduce(sum) over the staged gradient + [sum numer, sum logZ, n_utt, 0], in place, on the handle's stream."""
        self._check(self.lib.crfgpu_allreduce_grad(self.h))

    def fetch_tail(self):
        t = np.zeros(4, np.float64)
        self._check(self.lib.crfgpu_fetch_tail(self.h, _ptr(t, C.c_double)))
        return t

    def plan_info(self):
        buf = C.create_string_buffer(2048)
        n = self.lib.crfgpu_plan_info(self.h, buf, 2048)
        return buf.raw[:n].decode()

    def fetch_posterior_mass(self):
        out = np.zeros(self._n_frames, np.float32)
        self._check(self.lib.crfgpu_fetch_posterior_mass(self.h, _ptr(out, C.c_float)))
        return out

    def phase_ms(self, name):
        return float(self.lib.crfgpu_phase_ms(self.h, name.encode()))

    @property
    def stream(self):
        return self.lib.crfgpu_stream(self.h)

    @property
    def launch_count(self):
        return int(self.lib.crfgpu_launch_count(self.h))
