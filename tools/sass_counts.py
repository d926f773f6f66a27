#!/usr/bin/env python
"""Per-kernel counts of the SASS instructions that prove the tcgen05 / TMEM / TMA paths in libmsfno_b200.so
(mnemonics from /opt/skills/guides/B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG
(TMA tensor load / store), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), FFMA for contrast.
usage: python tools/sass_counts.py > profiles/r02_sass_tcgen05.txt     (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "modulated-spherical-fourier-neural-operator_b200", "libmsfno_b200.so")
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "FFMA", "HMMA", "MUFU"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    names = dict(zip(re.findall(r"Function : (\S+)", sass), demangle))
    counts = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names.get(m.group(1), m.group(1))
            cur = re.sub(r"\((bool|int|unsigned int)\)", "", cur)          # template-argument casts
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("msfno::", "")
            counts.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["total"] += 1
        base = op.split(".")[0]
        if base in MNEMONICS:
            counts[cur][base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
    print("# SASS instruction counts per kernel of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(lib, ROOT))
    print("%-64s %7s " % ("kernel", "total") + " ".join("%7s" % m[:7] for m in MNEMONICS))
    tot = collections.Counter()
    for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 10 ** 6 - kv[1]["total"]):
        print("%-64s %7d " % (k[:64], c["total"]) + " ".join("%7d" % c[m] for m in MNEMONICS))
        tot.update(c)
    print("%-64s %7d " % ("ALL KERNELS", tot["total"]) + " ".join("%7d" % tot[m] for m in MNEMONICS))


if __name__ == "__main__":
    main()
