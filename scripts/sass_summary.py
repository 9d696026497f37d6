"""SASS evidence table: per kernel of liboat.so, the count of Blackwell tensor / TMA / TMEM instructions
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG/UBLKCP,
mma.sync -> HMMA, cp.async -> LDGSTS). `python scripts/sass_summary.py > profiles/r2_sass_summary.txt` (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oa_transformer_b200", "liboat.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "HMMA", "LDGSTS", "MUFU", "total"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("oat::", "")
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            op = m.group(1)
            counts[name]["total"] += 1
            for k in KEYS[:-1]:
                if op.startswith(k):
                    counts[name][k] += 1
    print("SASS instruction counts per kernel of liboat.so (sm_100a), `python scripts/sass_summary.py`")
    print("%-78s" % "kernel" + "".join("%9s" % k for k in KEYS))
    for name, c in counts.items():
        if c["total"] < 40:
            continue
        print("%-78s" % name[:78] + "".join("%9d" % c[k] for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())
