#!/usr/bin/env python
"""Offline parity check of rows saved by `bench.py --dump-rows R` (gpurun_out/rows_<config>_rank<r>.npy + .json):
regenerates the same synthetic frames with the oracle's generator, stacks them with the CPU restatement of the
reference and compares every pixel bit for bit.  Needs no GPU: the large configurations (1024 x 8192-pixel rows,
linear fit: ~10 Mpx/s on 8 host threads) are checked here instead of while eight B200s wait.

    python tools/verify_rows.py gpurun_out/rows_c4_rank*.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import cpu_stack, host_synth_frames  # noqa: E402


def main():
    bad = 0
    for path in sys.argv[1:]:
        with open(path) as f:
            m = json.load(f)
        got = np.load(path[:-5] + ".npy")
        px = m["rows"] * m["width"]
        assert got.size == px
        t0 = time.time()
        frames = host_synth_frames(m["n_frames"], m["row0"] * m["width"], px)
        w = np.array(m["weights"], np.float32) if m["weights"] is not None else None
        want = cpu_stack(frames, m["mode"], w, m["sigma"], os.cpu_count() or 1)[0]
        gn, wn = np.isnan(got), np.isnan(want)
        same = bool(np.array_equal(gn, wn) and np.array_equal(got.view(np.uint32)[~gn], want.view(np.uint32)[~wn]))
        bad += 0 if same else 1
        print(json.dumps({"file": os.path.basename(path), "config": m["config"], "rank": m["rank"], "world": m["world"], "row0": m["row0"],
                          "rows": m["rows"], "pixels": px, "bit_exact": same, "cpu_seconds": round(time.time() - t0, 1)}), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
