#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/: counts of the Blackwell-native mnemonics (B200_PROFILING.md "What proves a
Blackwell-native kernel") in every kernel of higashi_b200/libhsg.so.

    python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "higashi_b200", "libhsg.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "MUFU", "LDG", "STG", "LDS", "STS", "BAR", "USETMAXREG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1).split(".")[0]
            kernels[cur]["_total"] += 1
            kernels[cur][op] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    print("# SASS mnemonic counts per kernel of higashi_b200/libhsg.so (cuobjdump -sass, sm_100a); tcgen05.mma -> UTCHMMA, tcgen05.ld/st ->")
    print("# LDTM/STTM, cp.async.bulk -> UBLKCP, cp.async.bulk.tensor -> UTMALDG/UTMASTG, mbarrier -> SYNCS, setmaxnreg -> USETMAXREG")
    print("kernel\ttotal\t" + "\t".join(KEYS))
    for k, c in kernels.items():
        short = re.sub(r"\(.*", "", names[k])[:70]
        print(short + "\t" + str(c["_total"]) + "\t" + "\t".join(str(c[x]) for x in KEYS))


if __name__ == "__main__":
    sys.exit(main())
