#!/usr/bin/env python
"""Static SASS size per source line / per region of nm_kernels.cu (no GPU needed).

    python tools/sass_lines.py [--kernel nm_step_kernelILb1] [--top 30]
Compiles nothing: expects nightmare_rl_b200/csrc/nm_kernels.o (python -m nightmare_rl_b200.build)."""
import argparse
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--obj", default=os.path.join(ROOT, "nightmare_rl_b200", "csrc", "nm_kernels.o"))
    ap.add_argument("--kernel", default="nm_step_kernelILb1")
    ap.add_argument("--top", type=int, default=30)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", a.obj], cwd=td, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    src = open(os.path.join(ROOT, "nightmare_rl_b200", "csrc", "nm_kernels.cu")).read().splitlines()
    banners = [(i, m.group(1).strip()) for i, l in enumerate(src, 1) if (m := re.search(r"// =+ (.*)$", l))]
    cnt = collections.Counter()
    cur_fn, cur_line, on = None, None, False
    for l in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            cur_fn = m.group(1); on = a.kernel in cur_fn
            continue
        m = re.search(r'//## File ".*nm_kernels.cu", line (\d+)', l)
        if m:
            cur_line = int(m.group(1))
            continue
        if on and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            cnt[cur_line] += 1
    tot = sum(cnt.values())
    reg = collections.OrderedDict()
    for ln in sorted(cnt):
        name = "helpers (inlined algebra)"
        if ln >= 320:
            for b, nm in banners:
                if ln >= b:
                    name = nm
        reg[name] = reg.get(name, 0) + cnt[ln]
    print(f"{a.kernel}: {tot} SASS instructions ({tot * 16 / 1024:.0f} KB)")
    for k, v in reg.items():
        print(f"  {k[:64]:64s} {v:6d} {100 * v / tot:5.1f}%")
    print("top lines:")
    for ln, c in cnt.most_common(a.top):
        print(f"  {ln:5d} {c:5d}  {src[ln - 1].strip()[:110]}")


if __name__ == "__main__":
    main()
