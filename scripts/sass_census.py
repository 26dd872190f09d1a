#!/usr/bin/env python
"""Per-kernel census of the Blackwell-specific SASS in the shipped library.

  python scripts/sass_census.py [path/to/libchessai_b200.so] > profiles/sass_census.txt

Runs `cuobjdump -sass` on the built .so and counts, per kernel, the mnemonics that prove which hardware path a kernel
uses (B200_PROFILING.md): DMMA (fp64 tensor pipe; tcgen05 has no fp64 kind), UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM
(tcgen05.ld / st: tensor memory), UTMALDG (TMA tensor loads), UBLKCP (1-D TMA bulk copies), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), plus HMMA / IMMA (legacy mma.sync, expected absent) and the register count from the ELF headers.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "chess-ai_b200", "lib", "libchessai_b200.so")
MNEMONICS = ["DMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "IMMA",
             "DFMA", "MUFU"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    counts[cur][mn] += 1
    regs = {}
    elf = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    fn = None
    for line in elf.split("\n"):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            regs[fn] = (int(m.group(1)), int(m.group(2)))
    names = demangle(order)
    print("# SASS census of %s (sm_100a), `cuobjdump -sass`; regenerate with scripts/sass_census.py" % os.path.relpath(LIB, ROOT))
    print("# columns: instructions, registers, then counts of " + " ".join(MNEMONICS))
    hdr = "%-78s %7s %4s " % ("kernel", "instr", "regs") + " ".join("%7s" % m for m in MNEMONICS)
    print(hdr)
    for fnm in sorted(order, key=lambda f: names[f]):
        c = counts[fnm]
        short = re.sub(r"\(.*", "", names[fnm])
        short = short.replace("void ", "").replace("cab::", "")
        print("%-78s %7d %4s " % (short[:78], c["_total"], regs.get(fnm, ("?", 0))[0]) + " ".join("%7d" % c[m] for m in MNEMONICS))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("%-78s %7d %4s " % ("TOTAL", tot["_total"], "") + " ".join("%7d" % tot[m] for m in MNEMONICS))


if __name__ == "__main__":
    main()
