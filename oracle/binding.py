"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the two CPU checkers.

* ``RefLib``    -> oracle/_ref/libcrfref.so : the unmodified ASR-CRaFT hot path (see oracle/ref_driver.cpp)
* ``OracleLib`` -> oracle/libcrforacle.so   : our C restatement (oracle/crf_oracle.c)

Both expose the same call shapes so tests can swap one for the other.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

MODEL_TYPES = {"stdframe": 0, "stdseg": 1, "stdseg_no_dur": 2,
               "stdseg_no_dur_no_transftr": 3, "stdseg_no_dur_no_segtransftr": 4}


class Config(C.Structure):
    """Field order shared by crfref_config (ref_driver.cpp) and crforacle_config (crf_oracle.h)."""
    _fields_ = [("model_type", C.c_uint32), ("n_labs", C.c_uint32), ("n_base_ftrs", C.c_uint32),
                ("n_states", C.c_uint32), ("max_dur", C.c_uint32), ("n_actual_labs", C.c_uint32),
                ("extract_seg_ftrs", C.c_uint32),
                ("use_state_ftrs", C.c_uint32), ("state_fidx_start", C.c_uint32), ("state_fidx_end", C.c_uint32),
                ("use_trans_ftrs", C.c_uint32), ("trans_fidx_start", C.c_uint32), ("trans_fidx_end", C.c_uint32),
                ("use_state_bias", C.c_uint32), ("use_trans_bias", C.c_uint32),
                ("state_bias_val", C.c_double), ("trans_bias_val", C.c_double),
                # context frames of stream 1 and the optional joined second stream (CRFTrain/src/Main.cpp:508-526)
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


def config_width(cfg):
    w = window_width(cfg.n_base_ftrs, cfg.max_dur, cfg.extract_seg_ftrs, cfg.left_ctx, cfg.right_ctx, cfg.boundary_delta)
    if cfg.n_base_ftrs2:
        w += window_width(cfg.n_base_ftrs2, cfg.max_dur, cfg.extract_seg_ftrs2, cfg.left_ctx2, cfg.right_ctx2, cfg.boundary_delta2)
    return w


def make_config(model_type="stdframe", n_labs=0, n_base_ftrs=0, n_states=1, max_dur=1, n_actual_labs=None,
                extract_seg_ftrs=0, use_trans_ftrs=0, state_fidx=None, trans_fidx=None,
                use_state_bias=1, use_trans_bias=1, state_bias_val=1.0, trans_bias_val=1.0,
                left_ctx=0, right_ctx=0, boundary_delta=0, n_base_ftrs2=0, extract_seg_ftrs2=0, left_ctx2=0, right_ctx2=0,
                boundary_delta2=0):
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


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class _Lib:
    prefix = ""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self._fn("last_error").restype = C.c_char_p

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self._fn("last_error")().decode())

    def lambda_len(self, cfg):
        out = C.c_uint32(0)
        self._check(self._fn("lambda_len")(C.byref(cfg), C.byref(out)))
        return out.value

    def fwdbwd(self, cfg, lam, off, ftrs, labs, n_threads=1, ftrs2=None):
        """ftrs / ftrs2: utterance u's rows of stream s start at off[u] + u * (left_ctx_s + right_ctx_s) and number
        T_u + left_ctx_s + right_ctx_s (off are the offsets of the LABELLED frames)"""
        lam = np.ascontiguousarray(lam, np.float64)
        off = np.ascontiguousarray(off, np.uint32)
        ftrs = np.ascontiguousarray(ftrs, np.float32)
        labs = np.ascontiguousarray(labs, np.uint32)
        n = len(off) - 1
        grad = np.zeros(len(lam), np.float64)
        numer = np.zeros(n, np.float64)
        logz = np.zeros(n, np.float64)
        f2 = np.ascontiguousarray(ftrs2, np.float32) if ftrs2 is not None else None
        self._check(self._fn("fwdbwd_mt2")(C.byref(cfg), _p(lam, C.c_double), C.c_uint32(len(lam)), C.c_uint32(n),
                                           _p(off, C.c_uint32), _p(ftrs, C.c_float), _p(f2, C.c_float) if f2 is not None else None,
                                           _p(labs, C.c_uint32),
                                           _p(grad, C.c_double), _p(numer, C.c_double), _p(logz, C.c_double),
                                           C.c_uint32(n_threads)))
        return grad, numer, logz

    def viterbi(self, cfg, lam, off, ftrs, ftrs2=None, lm=None, beam=0.0):
        """Returns list of (labels, durs, phones) per utterance, path costs, logZ.  lm = (start[P], bigram[P][P], final[P]) float32:
        a phone-bigram LM in the topology of the decoder's free-phone LM (one state per phone)."""
        f2 = np.ascontiguousarray(ftrs2, np.float32) if ftrs2 is not None else None
        lam = np.ascontiguousarray(lam, np.float64)
        off = np.ascontiguousarray(off, np.uint32)
        ftrs = np.ascontiguousarray(ftrs, np.float32)
        n = len(off) - 1
        tot = int(off[-1])
        lab = np.zeros(tot, np.uint32)
        dur = np.zeros(tot, np.uint32)
        phn = np.zeros(tot, np.uint32)
        nseg = np.zeros(n, np.uint32)
        cost = np.zeros(n, np.float32)
        logz = np.zeros(n, np.float64)
        lms = [None, None, None] if lm is None else [np.ascontiguousarray(a, np.float32) for a in lm]
        self._check(self._fn("viterbi_beam")(C.byref(cfg), _p(lam, C.c_double), C.c_uint32(len(lam)), C.c_uint32(n),
                                        _p(off, C.c_uint32), _p(ftrs, C.c_float), _p(f2, C.c_float) if f2 is not None else None,
                                        *[None if a is None else _p(a, C.c_float) for a in lms], C.c_double(beam),
                                        _p(lab, C.c_uint32),
                                        _p(dur, C.c_uint32), _p(phn, C.c_uint32), _p(nseg, C.c_uint32),
                                        _p(cost, C.c_float), _p(logz, C.c_double)))
        segs = []
        for u in range(n):
            b, k = int(off[u]), int(nseg[u])
            segs.append((lab[b:b + k].copy(), dur[b:b + k].copy(), phn[b:b + k].copy()))
        return segs, cost, logz

    def window_ftrs(self, cfg, ftrs, ftrs2=None):
        """ftrs: [left_ctx + T + right_ctx][n_base_ftrs] (ftrs2 likewise with its own context lengths) -> [T][max_dur][width]"""
        ftrs = np.ascontiguousarray(ftrs, np.float32)
        f2 = np.ascontiguousarray(ftrs2, np.float32) if ftrs2 is not None else None
        T = ftrs.shape[0] - cfg.left_ctx - cfg.right_ctx
        w = config_width(cfg)
        out = np.full((T, cfg.max_dur, w), np.nan, np.float32)
        self._check(self._fn("window_ftrs2")(C.byref(cfg), C.c_uint32(T), _p(ftrs, C.c_float),
                                             _p(f2, C.c_float) if f2 is not None else None, _p(out, C.c_float)))
        return out

    def window_labs(self, cfg, labs):
        labs = np.ascontiguousarray(labs, np.uint32)
        out = np.zeros((len(labs), 4), np.uint32)
        self._check(self._fn("window_labs")(C.byref(cfg), C.c_uint32(len(labs)), _p(labs, C.c_uint32),
                                            _p(out, C.c_uint32)))
        return out


class RefLib(_Lib):
    prefix = "crfref_"

    def __init__(self, path=None):
        super().__init__(path or os.path.join(HERE, "_ref", "libcrfref.so"))

    def viterbi_old(self, cfg, lam, off, ftrs):
        lam = np.ascontiguousarray(lam, np.float64)
        off = np.ascontiguousarray(off, np.uint32)
        ftrs = np.ascontiguousarray(ftrs, np.float32)
        lab = np.zeros(int(off[-1]), np.uint32)
        self._check(self.lib.crfref_viterbi_old(C.byref(cfg), _p(lam, C.c_double), C.c_uint32(len(lam)),
                                                C.c_uint32(len(off) - 1), _p(off, C.c_uint32),
                                                _p(ftrs, C.c_float), _p(lab, C.c_uint32)))
        return lab


class OracleLib(_Lib):
    prefix = "crforacle_"

    def __init__(self, path=None):
        super().__init__(path or os.path.join(HERE, "libcrforacle.so"))

    def fwdbwd(self, cfg, lam, off, ftrs, labs, n_threads=1, tied=False, ftrs2=None):
        """stdseg_no_dur* training has two restatements here.  Default: the native O(P^2 + D*P) recursion of
        crf_oracle.c::fb_nodur.  tied=True: stdseg_no_dur* training (CRF_StdSegStateNode_WithoutDurLab*, CRF_NewGradBuilder_StdSeg[_NoDur_NoTrans].cpp) is
        restated through its equivalence with `stdseg` on the (duration, phone) label set with TIED weights: state rows of
        label (d,y) = rows of phone y, transition (d',y')->(d,y) = transition y'->y (no transition FEATURES).  The reference's
        own stdseg run with tied lambda reproduces the no_dur logZ / numerators (tests/test_oracle.py pins this restatement to
        goldens produced by the reference's no_dur node classes)."""
        if tied and cfg.model_type in (2, 3, 4) and cfg.max_dur > 1 and cfg.n_states == 1 and not cfg.use_trans_ftrs:
            P, D = cfg.n_labs, cfg.max_dur
            w = window_width(cfg.n_base_ftrs, D, cfg.extract_seg_ftrs)
            nS = (cfg.state_fidx_end - cfg.state_fidx_start + 1) + (1 if cfg.use_state_bias else 0)
            nT = 1 if cfg.use_trans_bias else 0
            big = Config(1, P * D, cfg.n_base_ftrs, 1, D, P, cfg.extract_seg_ftrs, 1, cfg.state_fidx_start, cfg.state_fidx_end,
                         0, cfg.trans_fidx_start, cfg.trans_fidx_end, cfg.use_state_bias, cfg.use_trans_bias,
                         cfg.state_bias_val, cfg.trans_bias_val, *[getattr(cfg, f[0]) for f in Config._fields_[17:]])
            L = P * D
            lam = np.asarray(lam, np.float64)
            # 1-state layout (CRF_StdFeatureMap.cpp:295,365): label c owns [c*(nS+L*nT), +nS) state weights then L*nT transition weights
            src = np.empty(L * (nS + L * nT), np.int64)
            for c in range(L):
                y = c % P
                base_s, base_b = y * (nS + P * nT), c * (nS + L * nT)
                src[base_b:base_b + nS] = np.arange(base_s, base_s + nS)
                if nT:
                    src[base_b + nS:base_b + nS + L] = base_s + nS + (np.arange(L) % P)
            gb, numer, logz = _Lib.fwdbwd(self, big, lam[src], off, ftrs, labs, n_threads, ftrs2)
            grad = np.zeros(len(lam), np.float64)
            np.add.at(grad, src, gb)
            return grad, numer, logz
        if cfg.model_type in (2, 3, 4) and cfg.max_dur == 1 and not cfg.use_trans_ftrs and (tied or cfg.n_states > 1):
            # one-frame segments: the no_dur nodes reduce to the frame-level recursions (1 or N states per phone)
            frame = Config(0, *[getattr(cfg, f[0]) for f in Config._fields_[1:]])
            return _Lib.fwdbwd(self, frame, lam, off, ftrs, labs, n_threads, ftrs2)
        return _Lib.fwdbwd(self, cfg, lam, off, ftrs, labs, n_threads, ftrs2)


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libcrfref.so"))
