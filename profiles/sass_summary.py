#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (no GPU needed):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA engine),
HMMA = legacy mma.sync (must be 0), plus the instruction count of every kernel.
usage: python profiles/sass_summary.py [libb2048.so] > profiles/r02_sass_summary.csv"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rl-2048-with-reinforce-and-actor-critic_b200", "libb2048.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "ACQBULK", "SYNCS", "ATOMG", "REDG", "RED"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        total[cur] += 1
        for k in KEYS:
            if op == k or (k == "RED" and op == "RED"):
                counts[cur][k] += 1
print("kernel,instructions," + ",".join(KEYS))
for k, c in counts.items():
    print(f'"{k}",{total[k]},' + ",".join(str(c[x]) for x in KEYS))
