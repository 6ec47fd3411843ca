#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use in the shipped library.

    python tools/sass_counts.py > profiles/r2_sass_counts.txt

UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory), UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk (1-D TMA bulk copy), UTMALDG = cp.async.bulk.tensor (tensor-map TMA), SYNCS = mbarrier ops,
MUFU = special-function unit, FFMA2 / FADD2 / FMUL2 = packed fp32x2 arithmetic.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "aline_b200", "lib", "libaline_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "MUFU", "FFMA2", "FADD2", "FMUL2", "HFMA2",
             "FFMA", "LDG", "STG", "LDS", "STS", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in MNEMONICS:
                if op == k or op.startswith(k + ".") or (k in ("LDG", "STG", "LDS", "STS", "BAR", "MUFU", "SYNCS") and op.startswith(k)):
                    counts[cur][k] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} -- instruction counts per kernel (sm_100a)")
    print("# " + " ".join(f"{k:>8s}" for k in ["total"] + MNEMONICS) + "  kernel")
    for name, dn in zip(order, demangle):
        c = counts[name]
        short = re.sub(r"\(.*", "", dn)
        print("  " + " ".join(f"{c[k]:8d}" for k in ["_total"] + MNEMONICS) + "  " + short)
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("# library totals: " + ", ".join(f"{k}={tot[k]}" for k in MNEMONICS[:6]))


if __name__ == "__main__":
    sys.exit(main())
