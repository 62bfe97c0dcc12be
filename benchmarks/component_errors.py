#!/usr/bin/env python
"""Per-component accuracy of the fp32 mode on the GPU against the committed reference outputs: for every
Kelsall-family golden case the largest error of each component relative to max(|component|, floor * |total|)
for several floors.  Used to set the per-component gates of tests/helpers.py.  One JSON line per case."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
from helpers import case_ids, golden_case  # noqa: E402
from zodipy_b200 import engine  # noqa: E402


def main():
    worst = {}
    for cid in case_ids():
        case, a = golden_case(cid)
        if case["spec"]["kind"] != "kelsall":
            continue
        dm = engine.DeviceModel(case["spec"], 0)
        em = dm.evaluate(a["u"], a["obs"], a["earth"], return_comps=True, precision="fp32")
        ref = a["emission"]
        tot = np.abs(ref.sum(axis=0))[None, :]
        row = {"case": cid, "kernel": dm.kernel_name_for(a["u"].shape[1], "fp32")}
        for floor in (1.0, 1e-2, 1e-3, 1e-4):
            scale = np.maximum(np.abs(ref), floor * tot)
            e = np.nanmax(np.abs(em - ref) / scale, axis=1)
            row[f"floor_{floor:g}"] = [float(f"{v:.3g}") for v in e]
            w = worst.setdefault(floor, np.zeros(6))
            w[:e.size] = np.maximum(w[:e.size], e)
        print(json.dumps(row), flush=True)
    print(json.dumps({"case": "WORST", **{f"floor_{f:g}": [float(f"{v:.3g}") for v in w] for f, w in worst.items()}}))


if __name__ == "__main__":
    main()
