"""Derives the TIMIT train SHAPE (utterance lengths and segment durations in 10 ms frames, nothing else)
from the reference's bundled transcription demo/timit-aux/timit_train.mlf, for the 3696 utterances of
demo/timit-aux/timit_sisx_train.olist in list order.  Phone identities are NOT kept: the
synthetic workloads draw their own labels.  Output: tests/golden/timit_shape.npz (committed).

    python tests/golden/make_timit_shape.py
"""
import os

import numpy as np

AUX = "/root/reference/demo/timit-aux"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "timit_shape.npz")


def main():
    ids = [line.strip().replace(".lat", "") for line in open(os.path.join(AUX, "timit_sisx_train.olist")) if line.strip()]
    segs, cur = {}, None
    for line in open(os.path.join(AUX, "timit_train.mlf")):
        line = line.strip()
        if line.startswith('"'):
            cur = line.strip('"').split("/")[-1].replace(".lab", "")
            segs[cur] = []
        elif line and line[0].isdigit():
            s, e, _ = line.split()[:3]
            segs[cur].append((int(s), int(e)))
    utt_len, seg_cnt, seg_dur = [], [], []
    for i in ids:
        bounds = [0] + [(e + 50000) // 100000 for (_, e) in segs[i]]      # frame index of each segment end (100 ns -> 10 ms, rounded)
        durs = [b - a for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        utt_len.append(sum(durs)); seg_cnt.append(len(durs)); seg_dur.extend(durs)
    utt_len = np.array(utt_len, np.uint16); seg_cnt = np.array(seg_cnt, np.uint16); seg_dur = np.array(seg_dur, np.uint16)
    np.savez_compressed(OUT, utt_len=utt_len, seg_cnt=seg_cnt, seg_dur=seg_dur)
    print("utts", len(utt_len), "frames", int(utt_len.sum()), "mean", utt_len.mean(), "min", utt_len.min(), "max", utt_len.max(),
          "segments", len(seg_dur), "mean dur", seg_dur.mean(), ">10:", (seg_dur > 10).mean(), ">30:", (seg_dur > 30).mean())


if __name__ == "__main__":
    main()
