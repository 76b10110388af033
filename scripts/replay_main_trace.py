#!/usr/bin/env python
"""Replays the recorded main.py run (tests/golden/reference_main_trace.json) on the device and prints, per recorded snapshot,
how far rho / u / phase / phi / body_force / solid are from the reference's recorded fields.  Diagnostic twin of
tests/test_main_trace.py::test_gpu_replay_of_the_recorded_main_py_run (same replay function)."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_main_trace as T  # noqa: E402

if __name__ == "__main__":
    for which in T.TRACE_IDS:
        t = T.load_trace(which); fields = T.load_fields(which)
        report = []
        try:
            T.replay(t, fields, compare="--no-compare" not in sys.argv, report=report)
            print(f"{t['command']}: all", len(t["trace"]), "recorded calls executed")
        except BaseException:
            traceback.print_exc()
            print(f"{t['command']}: replay FAILED after", len(report), "snapshot comparisons")
        worst = {}
        for tag, k, err, ref, same in report:
            print(f"{tag:>8s} {k:>10s} max|diff| = {err:.3e}  max|ref| = {ref:.3e}  {'bit-exact' if same else ''}")
            worst[k] = max(worst.get(k, 0.0), err / max(1.0, ref))
        print("worst scaled difference per field:", worst, "| bit-exact comparisons:", sum(1 for r in report if r[4]), "of", len(report))
