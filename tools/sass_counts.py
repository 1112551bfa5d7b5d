#!/usr/bin/env python
"""Per-kernel SASS instruction counts of the built library: FP64 tensor-core MMAs (DMMA), TMA bulk copies (UBLKCP),
cp.async (LDGSTS), mbarrier operations (SYNCS), named barriers, rsqrt seeds.

    python tools/sass_counts.py > profiles/rNN_sass_counts.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jax_fdm_b200", "lib", "libfdm_b200.so")
PAT = {"DMMA": r"\bDMMA\b", "UBLKCP": r"\bUBLKCP", "LDGSTS": r"\bLDGSTS", "SYNCS": r"\bSYNCS",
       "BAR.SYNC(named)": r"BAR\.SYNC\S*\s+R|BAR\.SYNC\S*\s+0x[1-9]", "MUFU.RSQ64H": r"MUFU\.RSQ64H",
       "total": r"/\*[0-9a-f]{4,5}\*/"}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        for k, p in PAT.items():
            if re.search(p, line):
                counts[fn][k] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    total = collections.Counter()
    rows = []
    for (fn, c), name in zip(counts.items(), names):
        total.update(c)
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        rows.append((re.sub(r"\(.*", "", name), c))
    print("# SASS instruction counts of libfdm_b200.so (cuobjdump -sass, sm_100a), per kernel; tools/sass_counts.py")
    print(f"# whole library: {dict(total)}\n")
    for name, c in sorted(rows, key=lambda r: -r[1]["total"]):
        if c["total"] < 50:
            continue
        keys = [f"{k}={c[k]}" for k in PAT if k != "total" and c[k]]
        print(f"{name[:95]:95s} instr={c['total']:6d}  " + " ".join(keys))


if __name__ == "__main__":
    main()
