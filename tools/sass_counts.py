#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the built library (what proves a Blackwell-native kernel: UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA, HMMA = legacy mma.sync).  Writes profiles/<tag>_sass_counts.md.
  python tools/sass_counts.py <tag>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "phendiff_b200", "libphendiff_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "MUFU", "SYNCS"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "sass"
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = cur.replace("(anonymous namespace)::", "")
            cur = re.sub(r"\(.*", "", cur).replace("void pd::", "").replace("pd::", "")
            counts.setdefault(cur, collections.Counter())
            continue
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for k in MNEMONICS:
                if op.startswith(k):
                    counts[cur][k] += 1
            counts[cur]["total"] += 1
    # merge template instantiations of the same kernel (keep the max per mnemonic and the number of instantiations)
    merged = collections.OrderedDict()
    for name, c in counts.items():
        base = re.sub(r"<.*", "", name)
        e = merged.setdefault(base, {"n": 0, "max": collections.Counter()})
        e["n"] += 1
        for k, v in c.items():
            e["max"][k] = max(e["max"][k], v)
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_counts.md")
    with open(path, "w") as f:
        f.write(f"# SASS mnemonic counts per kernel of `phendiff_b200/libphendiff_b200.so` (`cuobjdump -sass`; max over the template "
                f"instantiations of a kernel)\n\n| kernel | instantiations | instructions | " + " | ".join(MNEMONICS) + " |\n|---|---:|---:|" + "---:|" * len(MNEMONICS) + "\n")
        for base, e in sorted(merged.items(), key=lambda kv: -kv[1]["max"]["UTCHMMA"] * 1000 - kv[1]["max"]["total"]):
            f.write(f"| `{base}` | {e['n']} | {e['max']['total']} | " + " | ".join(str(e["max"][k]) for k in MNEMONICS) + " |\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
