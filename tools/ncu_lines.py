#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line (needs -lineinfo builds and `--import-source on`).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [--launch 0] [--top 40] [--regions]

Prints the stall-reason totals, the opcode mix and the source lines that collect the most warp-stall
samples for one profiled launch.  With --regions the samples are bucketed by the `// ====` section
banners of nm_kernels.cu (kinematics, CRBA, collision, PGS, ...)."""
import argparse
import os
import collections
import csv
import io
import re
import subprocess


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    return rows, heads


def _i(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--regions", action="store_true")
    ap.add_argument("--src", default="nightmare_rl_b200/csrc/nm_kernels.cu")
    a = ap.parse_args()
    rows, heads = load(a.rep)
    # sections come in (file, kernel) pairs; keep those of nm_kernels.cu for the requested launch
    secs = []
    for k, h in enumerate(heads):
        end = heads[k + 1] - 2 if k + 1 < len(heads) else len(rows)
        fpath = rows[h - 2][1] if h >= 2 else ""
        fn = rows[h - 1][1] if h >= 1 else ""
        secs.append((fpath, fn, h, end))
    kernels = []
    for s in secs:
        if s[1] not in kernels:
            kernels.append(s[1])
    mine = [s for s in secs if s[0].endswith(os.path.basename(a.src))]
    per_launch = len(mine) // max(1, len([1 for s in secs if s[0].endswith(os.path.basename(a.src))]) // max(1, len(set(s[1] for s in mine)))) if mine else 0
    sec = mine[a.launch] if a.launch < len(mine) else mine[0]
    hdr = rows[sec[2]]
    ci = {c: i for i, c in enumerate(hdr)}
    stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    print(f"kernel: {sec[1]}  (section {a.launch} of {len(mine)})")
    line_tot = collections.OrderedDict()
    stall_tot = collections.Counter()
    op = collections.Counter()
    ninst = nsamp = 0
    cur = None
    for r in rows[sec[2] + 1: sec[3]]:
        if len(r) < len(hdr):
            continue
        if r[0] != "":
            cur = int(r[0])
            line_tot.setdefault(cur, dict(src=r[1], samp=0, inst=0, st=collections.Counter()))
            continue                      # the per-line summary row repeats the SASS totals below it
        if cur is None:
            continue
        s = _i(r[ci["# Samples"]])
        n = _i(r[ci["Instructions Executed"]])
        d = line_tot[cur]
        d["samp"] += s; d["inst"] += n
        nsamp += s; ninst += n
        for c in stall_cols:
            v = _i(r[ci[c]])
            if v:
                d["st"][c] += v; stall_tot[c] += v
        toks = r[3].strip().split()
        if toks:
            o = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            op[o.split(".")[0]] += n
    print(f"warp instructions executed: {ninst}   stall samples: {nsamp}")
    print("stall reasons: " + ", ".join(f"{k[6:]} {100 * v / max(1, nsamp):.1f}%" for k, v in stall_tot.most_common(8)))
    print("opcode mix:    " + ", ".join(f"{k} {100 * v / max(1, ninst):.1f}%" for k, v in op.most_common(14)))
    if a.regions:
        banners = []
        for i, l in enumerate(open(a.src), 1):
            m = re.search(r"// =+ (.*)$", l)
            if m:
                banners.append((i, m.group(1).strip()))
        reg = collections.OrderedDict()
        for ln, d in line_tot.items():
            name = "helpers (inlined)"
            for b, nm in banners:
                if ln >= b:
                    name = nm
            if ln < 320:
                name = "helpers (inlined small algebra / factorisations)"
            e = reg.setdefault(name, [0, 0])
            e[0] += d["samp"]; e[1] += d["inst"]
        print("\nregion                                                         samples    %   warp-inst    %")
        for nm, (s, n) in reg.items():
            print(f"{nm[:60]:60s} {s:8d} {100 * s / max(1, nsamp):5.1f} {n:10d} {100 * n / max(1, ninst):5.1f}")
    print(f"\ntop {a.top} source lines by stall samples")
    for ln, d in sorted(line_tot.items(), key=lambda kv: -kv[1]["samp"])[: a.top]:
        st = ",".join(f"{k[6:]}:{v}" for k, v in d["st"].most_common(3))
        print(f"{ln:5d} samp {d['samp']:5d} ({100 * d['samp'] / max(1, nsamp):4.1f}%) inst {d['inst']:8d}  [{st}]  {d['src'].strip()[:90]}")


if __name__ == "__main__":
    main()
