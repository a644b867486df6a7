"""Per-kernel count of the SASS opcodes that prove the Blackwell paths (profiles/README.md):
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.   python tools/sass_summary.py > profiles/r02_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "plangen_b200", "libplangen_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "FFMA", "HMMA"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(.*", "", cur).replace("pg::", "").replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur:
        for o in ops:
            if re.search(r"\b" + o + r"\b|\b" + o + r"\.", line):
                counts[cur][o] += 1
print(f"# SASS opcode counts per kernel of plangen_b200/libplangen_b200.so (sm_100a); {len(counts)} kernels")
print(f"{'kernel':58s} " + " ".join(f"{o:>8s}" for o in ops))
for k, c in counts.items():
    if any(c[o] for o in ops[:5]) or "attn" in k or "gemm" in k:
        print(f"{k[:58]:58s} " + " ".join(f"{c[o]:8d}" for o in ops))
