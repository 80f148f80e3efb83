"""Seeded synthetic workloads of BASELINE.json's configs, shared by bench.py, the tests and
__graft_entry__.smoke().  Pure numpy; reads only the committed shape fixture tests/golden/timit_shape.npz
(utterance lengths and segment durations of the TIMIT train set), never /root/reference.

Every function returns (cfg_kwargs, lam_seed_info, off, ftrs, labs) pieces as a dict so that both the
oracle binding and crf_b200 can build their own Config from the same kwargs.
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
N_PHONES = 61
N_BASE_FTRS = 105


def timit_shape():
    z = np.load(os.path.join(ROOT, "tests", "golden", "timit_shape.npz"))
    return z["utt_len"].astype(np.int64), z["seg_cnt"].astype(np.int64), z["seg_dur"].astype(np.int64)


def _features(rng, frame_phone, n_ftrs=N_BASE_FTRS, n_post=N_PHONES):
    """First n_post dims: softmax(N(0,1) + 4*onehot(true phone)) (posterior-like); rest U[0,1) (attribute-like)."""
    n = len(frame_phone)
    logits = rng.standard_normal((n, n_post), dtype=np.float32)
    logits[np.arange(n), frame_phone % n_post] += 4.0
    logits -= logits.max(axis=1, keepdims=True)
    np.exp(logits, out=logits)
    logits /= logits.sum(axis=1, keepdims=True)
    out = np.empty((n, n_ftrs), np.float32)
    out[:, :n_post] = logits
    if n_ftrs > n_post:
        out[:, n_post:] = rng.random((n, n_ftrs - n_post), dtype=np.float32)
    return out


def timit_train_batch(first_utt=0, n_utt=3696, n_phones=N_PHONES):
    """Utterances [first_utt, first_utt+n_utt) of the TIMIT-shaped train set: real lengths and segment
    boundaries, phone ids U{0..n_phones-1} with no two adjacent segments equal (seed 9), features seed 2.
    Labels and features of an utterance do not depend on which slice is requested."""
    return timit_train_utts(range(first_utt, first_utt + n_utt), n_phones)


def timit_train_utts(ids, n_phones=N_PHONES):
    """The same utterances by id (any subset, in the given order): what a rank of a length-balanced global minibatch stages."""
    utt_len, seg_cnt, seg_dur = timit_shape()
    seg_start = np.concatenate([[0], np.cumsum(seg_cnt)])
    sel = list(ids)
    lens = utt_len[np.asarray(sel, np.int64)] if len(sel) else np.zeros(0, np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    labs = np.empty(int(off[-1]), np.uint32)
    ftrs = np.empty((int(off[-1]), N_BASE_FTRS), np.float32)
    for k, u in enumerate(sel):
        rl = np.random.default_rng([9, u])
        durs = seg_dur[seg_start[u]:seg_start[u + 1]]
        phones = np.empty(len(durs), np.int64)
        prev = -1
        for i in range(len(durs)):
            p = int(rl.integers(0, n_phones))
            while p == prev:
                p = int(rl.integers(0, n_phones))
            phones[i] = prev = p
        fl = np.repeat(phones, durs)
        labs[off[k]:off[k + 1]] = fl
        ftrs[off[k]:off[k + 1]] = _features(np.random.default_rng([2, u]), fl)
    return off, ftrs, labs


def cfg2_kwargs():
    """frame-level CRF, 61 labels, 1 state/phone, 105 features, state+transition bias (dim(lambda)=10 187)."""
    return dict(model_type="stdframe", n_labs=N_PHONES, n_base_ftrs=N_BASE_FTRS)


def cfg3_kwargs():
    """3-state/phone frame CRF decoded through stdseg_no_dur_no_segtransftr with maxDur 1 (dim(lambda)=23 424)."""
    return dict(model_type="stdseg_no_dur_no_segtransftr", n_labs=3 * N_PHONES, n_base_ftrs=N_BASE_FTRS, n_states=3)


def cfg4_kwargs():
    """segmental stdseg, 61 phones x maxDur 10 = 610 labels, 850 segment features (dim(lambda)=891 210)."""
    return dict(model_type="stdseg", n_labs=10 * N_PHONES, n_base_ftrs=N_BASE_FTRS, max_dur=10,
                n_actual_labs=N_PHONES, extract_seg_ftrs=1)


def cfg3_batch(n_utt=1680):
    """1680 utterances, lengths round(N(304,80^2)) clipped to [92,778] (seed 4), features seed 5."""
    rng = np.random.default_rng(4)
    lens = np.clip(np.rint(rng.normal(304, 80, n_utt)), 92, 778).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    r5 = np.random.default_rng(5)
    phone = r5.integers(0, N_PHONES, int(off[-1]))
    return off, _features(r5, phone)


def cfg5_kwargs(n_phones=1024, max_dur=30, n_base_ftrs=64):
    """stress: segmental CRF without duration labels, 1024 phones, maxDur 30, 8*64+30 = 542 segment features
    (stdseg_no_dur_no_segtransftr: the stdseg transition table would be (P*D)^2 weights; dim(lambda) = 1024*(543+1024))."""
    return dict(model_type="stdseg_no_dur_no_segtransftr", n_labs=n_phones, n_base_ftrs=n_base_ftrs, max_dur=max_dur,
                n_actual_labs=n_phones, extract_seg_ftrs=1)


def cfg5_batch(n_utt=64, n_frames=2000, n_phones=1024, max_dur=30, n_base_ftrs=64):
    """64 utterances x 2000 frames, reference segments dur ~ U{1..30} with no two adjacent phones equal, features U[0,1) (seed 8)."""
    rng = np.random.default_rng(8)
    off = (np.arange(n_utt + 1) * n_frames).astype(np.uint32)
    ftrs = rng.random((n_utt * n_frames, n_base_ftrs), dtype=np.float32)
    labs = np.empty(n_utt * n_frames, np.uint32)
    for u in range(n_utt):
        t, prev = 0, -1
        while t < n_frames:
            d = int(rng.integers(1, max_dur + 1))
            lab = int(rng.integers(0, n_phones))
            while lab == prev:
                lab = int(rng.integers(0, n_phones))
            labs[u * n_frames + t:u * n_frames + min(t + d, n_frames)] = lab
            prev, t = lab, t + d
    return off, ftrs, labs


RECIPE_PHONES, RECIPE_FTRS, RECIPE_DUR, RECIPE_CTX = 48, 144, 10, 6


def recipe_kwargs():
    """The production TIMIT recipe (demo/segmental-timit-demo.cfg.in:11-48): stdseg_no_dur_no_segtransftr + stdtrans,
    48 phones, maxDur 10; stream 1 = 144 inputs -> 8*144 + 10 = 1162 segment features (state features), stream 2 = the same inputs
    padded by 6 frames on each side -> 13*144 = 1872 context features (transition features); dim(lambda) = 48*1163 + 48^2*1873."""
    w1, w2 = 8 * RECIPE_FTRS + RECIPE_DUR, (2 * RECIPE_CTX + 1) * RECIPE_FTRS
    return dict(model_type="stdseg_no_dur_no_segtransftr", n_labs=RECIPE_PHONES, n_base_ftrs=RECIPE_FTRS, max_dur=RECIPE_DUR,
                n_actual_labs=RECIPE_PHONES, extract_seg_ftrs=1, n_base_ftrs2=RECIPE_FTRS, left_ctx2=RECIPE_CTX, right_ctx2=RECIPE_CTX,
                use_trans_ftrs=1, state_fidx=(0, w1 - 1), trans_fidx=(w1, w1 + w2 - 1))


def recipe_batch(n_utt=128):
    """The first n_utt utterances of the TIMIT-shaped train set (real lengths and segment boundaries) with 48 phone ids drawn like the
    other workloads (seed 9), stream 1 U[0,1) (seed 11), stream 2 U[0,1) with 6 context frames on each side of every utterance (seed 12).
    Returns off, ftrs1, ftrs2, labs."""
    utt_len, seg_cnt, seg_dur = timit_shape()
    seg_start = np.concatenate([[0], np.cumsum(seg_cnt)])
    lens = utt_len[:n_utt]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    labs = np.empty(int(off[-1]), np.uint32)
    for u in range(n_utt):
        rl = np.random.default_rng([9, u])
        durs = seg_dur[seg_start[u]:seg_start[u + 1]]
        phones = np.empty(len(durs), np.int64)
        prev = -1
        for i in range(len(durs)):
            p = int(rl.integers(0, RECIPE_PHONES))
            while p == prev:
                p = int(rl.integers(0, RECIPE_PHONES))
            phones[i] = prev = p
        labs[off[u]:off[u + 1]] = np.repeat(phones, durs)
    f1 = np.random.default_rng(11).random((int(off[-1]), RECIPE_FTRS), dtype=np.float32)
    f2 = np.random.default_rng(12).random((int(off[-1]) + 2 * RECIPE_CTX * n_utt, RECIPE_FTRS), dtype=np.float32)
    return off, f1, f2, labs


def recipe_utt(off, f1, f2, labs, u):
    """utterance u of a recipe batch as a batch of its own (stream 2 carries its own context frames)"""
    a, b = int(off[u]), int(off[u + 1])
    a2, b2 = a + 2 * RECIPE_CTX * u, b + 2 * RECIPE_CTX * (u + 1)
    return np.array([0, b - a], np.uint32), f1[a:b], f2[a2:b2], labs[a:b]


def lam_for(name, n):
    seed, scale = {"cfg2": (3, 0.25), "cfg3": (6, 0.25), "cfg4": (7, 0.01), "cfg5": (8, 0.01), "recipe": (10, 0.01)}[name]
    return np.random.default_rng(seed).uniform(-scale, scale, n)
