#!/usr/bin/env python
"""Per-loop and per-instruction view of an `ncu --page source --csv --print-source sass` export.

    python tools/ncu_source_hot.py gpurun_out/foo.source.csv [--top 25] [--range 0x1b570 0x1ba10]

Prints the warp-stall sample share and executed-instruction share of address ranges (loops found from
backward branches), then the hottest instructions with their dominant stall reasons.
"""
import argparse
import csv
import re
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--range", nargs=2, default=None)
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ins = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0]:
            continue
        try:
            addr = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
        except ValueError:
            continue

        def num(name):
            try:
                return float(r[col[name]].replace(",", "") or 0)
            except ValueError:
                return 0.0
        ins.append({"addr": addr, "src": r[col["Source"]], "samples": num("# Samples"),
                    "exec": num("Instructions Executed"), "stalls": {s: num(s) for s in stall_cols}})
    base = ins[0]["addr"]
    for i in ins:
        i["off"] = i["addr"] - base
    tot_s = sum(i["samples"] for i in ins) or 1.0
    tot_e = sum(i["exec"] for i in ins) or 1.0
    off_idx = {i["off"]: k for k, i in enumerate(ins)}
    print(f"{len(ins)} instructions, {tot_s:.0f} samples, {tot_e:.4g} warp instructions executed")
    loops = []
    for k, i in enumerate(ins):
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", i["src"])
        if m:
            tgt = int(m.group(1), 16)
            if tgt >= base:  # the export prints absolute addresses
                tgt -= base
            if tgt <= i["off"] and tgt in off_idx:
                loops.append((tgt, i["off"]))
    if args.range:
        loops = [(int(args.range[0], 16), int(args.range[1], 16))]
    print("loops (offset range): share of samples, share of executed instructions, samples per executed instr")
    for lo, hi in loops:
        body = [i for i in ins if lo <= i["off"] <= hi]
        s, e = sum(i["samples"] for i in body), sum(i["exec"] for i in body)
        if s / tot_s < 0.01 and not args.range:
            continue
        st = {}
        for i in body:
            for k2, v in i["stalls"].items():
                st[k2] = st.get(k2, 0.0) + v
        top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
        n_iter = max(i["exec"] for i in body)
        print(f"  0x{lo:05x}-0x{hi:05x}: {len(body):4d} instr  samples {100 * s / tot_s:5.1f} %  executed {100 * e / tot_e:5.1f} %"
              f"  exec/iter {e / n_iter:6.1f}  stalls " + ", ".join(f"{k2[6:]} {100 * v / max(s, 1):.0f}%" for k2, v in top))
    sel = ins if not args.range else [i for i in ins if loops[0][0] <= i["off"] <= loops[0][1]]
    print("hottest instructions:")
    for i in sorted(sel, key=lambda x: -x["samples"])[:args.top]:
        top = sorted(i["stalls"].items(), key=lambda kv: -kv[1])[:3]
        print(f"  0x{i['off']:05x} {100 * i['samples'] / tot_s:5.2f} %  {i['src'][:70]:70s} " +
              ", ".join(f"{k2[6:]} {v:.0f}" for k2, v in top if v))


if __name__ == "__main__":
    main()
