"""Writes tests/golden/stack_golden.npz and kats.json.

The reference (Go) cannot be built in this environment (no Go toolchain, un-vendored modules), so the
golden vectors are (1) the hand-derivable known-answer vectors of SURVEY.md section 8c, which were derived
with an independent float32 transliteration and are typed in literally below, and (2) outputs of the
oracle (oracle/nl_oracle.c, itself pinned by (1), by the reference's own qsort test and by tests/pyref.py)
on seeded inputs.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# SURVEY.md section 8c, literal
C10 = [100.5, 101.25, 99.75, 100, 250, 100.25, 99.5, 3, 100.75, 101]
KATS = {
    "columns": [
        {"col": [10, 11, 9, 10, 12, 100, 10, 9], "mode": "sigma", "sig": 2.0, "hex": "411d5555", "clip": [0, 2]},
        {"col": ["nan", "nan", "nan"], "mode": "sigma", "sig": 2.0, "hex": "00000000", "clip": [0, 0]},
        {"col": ["nan", 4, "nan", 6], "mode": "sigma", "sig": 2.0, "hex": "40a00000", "clip": [0, 0]},
        {"col": C10, "mode": "sigma", "sig": 2.75, "hex": "42d33333", "clip": [0, 0]},
        {"col": C10, "mode": "winsor", "sig": 2.75, "hex": "42c8c000", "clip": [1, 1]},
        {"col": C10, "mode": "mad", "sig": 2.75, "hex": "42c8c000", "clip": [1, 1]},
        {"col": C10, "mode": "linfit", "sig": 2.75, "hex": "42d33333", "clip": [0, 0]},
        {"col": C10, "mode": "sigma", "sig": -1.0, "hex": "42d33333", "clip": [9, 1]},
    ],
    "qselect": [
        {"col": [5, 1, 4, 2, 3], "median": 3.0, "after": [1, 2, 3, 4, 5]},
        {"col": [7, 3, 9, 1, 8, 2], "median": 5.0, "after": [1, 2, 3, 7, 8, 9]},
    ],
    # generator KATs: seed 12345, sigma 2.75, weights w[k] = 1/(1+4*((k%7)/6)); value/clipLow/clipHigh
    "generator": {
        "order": ["median", "mean", "sigma", "winsor", "mad", "linfit", "sigma_w", "winsor_w"],
        "rows": [
            [16, 0, "447cc660", "449a91ce", "44717c6f/0/1", "447fb26c/2/1", "44791d37/1/1", "4484222c/4/1", "4469264c/0/1", "4480545b/2/1"],
            [16, 1, "447524c0", "44bf56ec", "4478978e/0/2", "4471222b/0/4", "4478978e/0/2", "4478978e/0/2", "447452af/0/2", "4470c0ae/0/4"],
            [16, 3, "44815dc0", "447d1128", "447d1128/0/0", "447d1128/0/0", "447d1128/0/0", "4481e4e5/2/0", "447fe4cd/0/0", "447fe4cd/0/0"],
            [16, 7, "447284a0", "4475b8ec", "44718740/0/1", "44718740/0/1", "44718740/0/1", "44718740/0/1", "4472416c/0/1", "4472416c/0/1"],
            [64, 0, "4480e800", "44854a54", "4480e0ae/4/1", "4480e0ae/4/1", "4480e0ae/4/1", "4482525d/22/5", "4480f9b3/4/1", "44813a47/4/1"],
            [64, 3, "4484b4a0", "4480fcdb", "4483435e/2/0", "4483435e/2/0", "4483435e/2/0", "44826dae/6/16", "4482840e/2/0", "4482840e/2/0"],
            [64, 7, "44784080", "44857ff5", "447a9e25/0/1", "447a9e25/0/1", "447a9e25/0/1", "44791e86/5/8", "447b6c8d/0/1", "447b6c8d/0/1"],
        ],
        "first_column_n16_p0": "44238980 45a57de8 448d3400 44484d80 4406b180 4458d240 447d9540 447b9e00 44781400 "
                               "44931360 4488fe20 446720c0 447bf780 4484aaa0 448cbd80 44889a80",
    },
    "project": {
        "w": 4, "h": 4,
        "trans_hex": ["3f7ff605", "bc8ef859", "3f000000", "3c8ef859", "3f7ff605", "3e800000"],
        "inverse_hex": ["3f7ff605", "3c8ef859", "bf0118f3", "bc8ef859", "3f7ff605", "be77067f"],
        "rows_hex": [["nan", "nan", "nan", "nan"],
                     ["nan", "40fd9666", "410bffb9", "41193440"],
                     ["nan", "418f8639", "4196207c", "419cbac0"],
                     ["nan", "41dfa6d8", "41e6411c", "41ecdb60"]],
    },
}


def weights_for(n):
    k = np.arange(n) % 7
    return (np.float32(1) / (np.float32(1) + np.float32(4) * (k.astype(np.float32) / np.float32(6)))).astype(np.float32)


def main():
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump(KATS, f, indent=1)
    out = {}
    # oracle outputs on the synthetic generator: config-1 shape columns and a ragged small case
    for n, p0, count in ((16, 0, 4096), (37, 777, 1031), (256, 4096 * 4096 - 512, 512)):
        frames = O.synth_frames(n, p0, count)
        w = weights_for(n)
        for mode in ("median", "mean", "sigma", "winsor", "mad", "linfit"):
            for weighted in (False, True):
                if weighted and mode in ("median", "mad", "linfit"):
                    continue
                res, cl, ch = O.stack(frames, mode, 2.75, 2.75, weights=w if weighted else None)
                key = "n%d_p%d_c%d_%s%s" % (n, p0, count, mode, "_w" if weighted else "")
                out[key] = res
                out[key + "_clip"] = np.array([cl, ch], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "stack_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
