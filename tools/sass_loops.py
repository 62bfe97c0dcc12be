#!/usr/bin/env python
"""Static instruction mix of the loops of one kernel in a built library (cuobjdump -sass).

    python tools/sass_loops.py [lib.so] <substring of the mangled kernel name>

For every backward branch (a loop) prints the number of SASS instructions between its target and
itself, split by pipe class: MUFU (XU pipe), packed fp32 (FFMA2/FMUL2/FADD2), scalar fp32, FP64,
LDS, votes/shuffles and the rest.  Nested loops are reported separately (inner bodies included in
the outer count).  Used to budget issue slots / XU cycles before spending GPU time.
"""
import re
import subprocess
import sys


def main():
    args = sys.argv[1:]
    lib = "zodipy_b200/libzodi_b200.so"
    if args and args[0].endswith(".so"):
        lib = args.pop(0)
    want = args[0]
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, funcs = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    for name, ins in funcs.items():
        if want not in name:
            continue
        print(f"== {name}: {len(ins)} instructions")
        addr_idx = {a: i for i, (a, _) in enumerate(ins)}
        for i, (a, op) in enumerate(ins):
            m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", op)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt > a or tgt not in addr_idx:
                continue
            body = [o for _, o in ins[addr_idx[tgt]:i + 1]]
            cls = {"MUFU": 0, "F32x2": 0, "F32": 0, "F64": 0, "LDS": 0, "VOTE/SHFL": 0, "BRA": 0, "other": 0}
            for o in body:
                o2 = re.sub(r"^@!?U?P\d+\s+", "", o)
                mn = o2.split()[0]
                if mn.startswith("MUFU"):
                    cls["MUFU"] += 1
                elif mn.startswith(("FFMA2", "FMUL2", "FADD2")):
                    cls["F32x2"] += 1
                elif mn.startswith(("FFMA", "FMUL", "FADD", "FMNMX", "FSEL", "FSETP")):
                    cls["F32"] += 1
                elif mn.startswith(("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")):
                    cls["F64"] += 1
                elif mn.startswith("LDS"):
                    cls["LDS"] += 1
                elif mn.startswith(("VOTE", "SHFL")):
                    cls["VOTE/SHFL"] += 1
                elif mn.startswith(("BRA", "BSSY", "BSYNC", "WARPSYNC")):
                    cls["BRA"] += 1
                else:
                    cls["other"] += 1
            print(f"loop 0x{tgt:05x}-0x{a:05x}: {len(body):4d} instr  " + "  ".join(f"{k}={v}" for k, v in cls.items()))


if __name__ == "__main__":
    main()
