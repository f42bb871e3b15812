#!/usr/bin/env python
"""Per-kernel SASS opcode census of the shipped library (runs here, no GPU needed): which kernels carry tcgen05 / TMEM / TMA
instructions, and how many.

    python tools/sass_census.py [--out profiles/r02_sass_census.md]

`cuobjdump -sass 3dfeatnet_b200/lib3dfeatnet_b200.so`, split per function; counted mnemonics (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): UTCHMMA (tcgen05.mma kind::f16), UTCCP (tcgen05.cp), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit),
UTCATOMSWS (tcgen05.alloc / dealloc), UBLKCP (cp.async.bulk), UTMALDG / UTMASTG (cp.async.bulk.tensor), SYNCS (mbarrier), CREDUX (redux.sync),
UCGABAR (cluster barrier), HMMA (legacy mma.sync -- expected 0), LDGSTS (cp.async)."""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "CREDUX", "UCGABAR", "HMMA", "LDGSTS"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "3dfeatnet_b200", "lib3dfeatnet_b200.so"))
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
    counts, total, arch = collections.OrderedDict(), collections.Counter(), set()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            counts[cur]["_all"] += 1
            for o in OPS:
                if op == o or op.startswith(o):
                    counts[cur][o] += 1
                    total[o] += 1
                    break
    names = demangle(list(counts))
    rows = []
    for fn, c in counts.items():
        if any(c[o] for o in OPS if o != "LDGSTS"):
            nm = re.sub(r"^void ", "", names[fn])
            nm = re.sub(r"\(.*$", "", nm).replace("f3d::", "")
            rows.append((nm, c))
    rows.sort(key=lambda r: -(r[1]["UTCHMMA"] * 1000 + r[1]["UBLKCP"] + r[1]["LDTM"]))
    out = ["# SASS opcode census of lib3dfeatnet_b200.so (tools/sass_census.py; cuobjdump -sass, arch %s, %d kernels)" % (", ".join(sorted(arch)) or "?", len(counts)),
           "",
           "Kernels that carry at least one tensor-core / TMEM / TMA / mbarrier / redux instruction (static counts per kernel; loops are not unrolled in this count).",
           "tcgen05.mma = UTCHMMA, tcgen05.cp = UTCCP, tcgen05.ld/st = LDTM/STTM, tcgen05.commit = UTCBAR, tcgen05.alloc/dealloc = UTCATOMSWS,",
           "cp.async.bulk = UBLKCP (1-D bulk copies on the TMA engine; the activations arrive by index gather, so there is no tensor-map",
           "UTMALDG), mbarrier = SYNCS, redux.sync = CREDUX, cluster barrier = UCGABAR.  HMMA (legacy mma.sync) is expected to be 0 everywhere.",
           "",
           "| kernel | instructions | " + " | ".join(OPS) + " |",
           "|---|---|" + "---|" * len(OPS)]
    for nm, c in rows:
        out.append("| `%s` | %d | " % (nm[:90], c["_all"]) + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
    out.append("| **whole library** | %d | " % sum(c["_all"] for c in counts.values()) + " | ".join(str(total[o]) for o in OPS) + " |")
    text = "\n".join(out) + "\n"
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
