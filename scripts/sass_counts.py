"""Per-kernel counts of the Blackwell-native SASS mnemonics in libtt_b200.so (evidence that the contractions run on
tcgen05 / TMEM / TMA, not on recompiled mma.sync): python scripts/sass_counts.py > profiles/rNN_sass_tcgen05.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "two-towers-overlords_b200", "lib", "libtt_b200.so")
OPS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "SYNCS", "ELECT", "HMMA")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(anonymous namespace\)::", "", cur).split("(")[0]
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and m.group(1) in OPS:
            counts[cur][m.group(1)] += 1
    print("# cuobjdump -sass two-towers-overlords_b200/lib/libtt_b200.so, per-kernel count of Blackwell-native mnemonics")
    print("# UTCHMMA = tcgen05.mma kind::f16 | LDTM / STTM = tcgen05.ld / st (TMEM) | UTMALDG = TMA tensor load |")
    print("# UTCBAR = tcgen05.commit -> mbarrier | SYNCS = mbarrier arrive/try_wait | ELECT = elect.sync | HMMA = legacy mma.sync (must be 0)")
    for k, c in sorted(counts.items()):
        if any(c[o] for o in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "HMMA")):
            print(f"{k}: " + ", ".join(f"{o}={c[o]}" for o in OPS if c[o]))
    total = collections.Counter()
    for c in counts.values():
        total.update(c)
    print("TOTAL: " + ", ".join(f"{o}={total[o]}" for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
