#!/usr/bin/env python
"""Fold an `ncu --set full --import-source on` report per CUDA source line: share of stall samples, of executed warp
instructions, and the dominant stall reasons.

    python scripts/ncu_hotspots.py report.ncu-rep [kernel-substring] [top-n]
"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                          capture_output=True, text=True).stdout
    rows = csv.reader(text.splitlines())
    path = func = None
    hdr = None
    lines = defaultdict(lambda: defaultdict(float))
    src = {}
    seen_funcs = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            if func not in seen_funcs:
                seen_funcs.append(func)
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].isdigit():
            continue
        if want and want not in func:
            continue
        if func != seen_funcs[0] and not want:
            continue
        key = (path, int(r[0]))
        src[key] = r[1].strip()
        for i, name in enumerate(hdr):
            if i < 4:
                continue
            try:
                lines[key][name] += float(r[i])
            except ValueError:
                pass
    tot_s = sum(v["# Samples"] for v in lines.values()) or 1.0
    tot_i = sum(v["Instructions Executed"] for v in lines.values()) or 1.0
    stall_names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    tot_st = {n: sum(v[n] for v in lines.values()) for n in stall_names}
    print(f"# kernel: {seen_funcs[0] if not want else want}; samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
    print("# stall reasons: " + ", ".join(f"{n[6:]} {100 * c / tot_s:.1f}%" for n, c in sorted(tot_st.items(), key=lambda x: -x[1]) if c > 0.01 * tot_s))
    byfile = defaultdict(lambda: [0.0, 0.0])
    for (p, _), v in lines.items():
        byfile[p][0] += v["# Samples"]
        byfile[p][1] += v["Instructions Executed"]
    print("# by file: " + ", ".join(f"{p} samples {100 * a / tot_s:.1f}% instr {100 * b / tot_i:.1f}%" for p, (a, b) in sorted(byfile.items(), key=lambda x: -x[1][1])))
    ranked = sorted(lines.items(), key=lambda kv: -(kv[1]["# Samples"] / tot_s + kv[1]["Instructions Executed"] / tot_i))
    for (p, ln), v in ranked[:top]:
        st = sorted(((v[n], n[6:]) for n in stall_names), reverse=True)[:2]
        print(f"{p:16s}:{ln:5d} samples {100 * v['# Samples'] / tot_s:5.1f}% instr {100 * v['Instructions Executed'] / tot_i:5.1f}%  "
              f"[{', '.join(f'{n} {c:.0f}' for c, n in st if c > 0)}]  {src[(p, ln)][:110]}")


if __name__ == "__main__":
    main()
