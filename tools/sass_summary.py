"""Instruction histogram per kernel of libkgan.so (cuobjdump -sass): the mnemonics that prove the Blackwell-native path
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,
UTCATOMSWS = TMEM alloc, SYNCS = mbarrier, LDGSTS = cp.async, REDG/RED = vector atomics) next to the SIMT ones.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kinetic-gan_b200", "csrc", "libkgan.so")
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "REDG", "RED", "ATOMG", "FFMA", "HMMA",
       "LDG", "STG", "LDS", "STS", "ELECT", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(?:\.|\s|;)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("libkgan.so: %d kernels, cubin architectures: %s" % (len(kernels), ", ".join(arch)))
    print("%-64s %6s  %s" % ("kernel", "instrs", "  ".join(KEY)))
    tot = collections.Counter()
    for (name, c), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm).replace("kgan::", "")
        tot.update(c)
        print("%-64s %6d  %s" % (short[:64], sum(c.values()), "  ".join("%*d" % (len(k), c.get(k, 0)) for k in KEY)))
    print("%-64s %6d  %s" % ("TOTAL", sum(tot.values()), "  ".join("%*d" % (len(k), tot.get(k, 0)) for k in KEY)))


if __name__ == "__main__":
    sys.exit(main())
