"""SASS opcode census of the in-tree objects: proof that the hot kernels are tcgen05 / TMA / TMEM code.
   python scripts/sass_summary.py > profiles/rNN_sass_opcodes.txt   (runs anywhere cuobjdump is installed)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "sid_lsg_b200", "_C", "obj")
OPS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "LDTM", "STTM", "ELECT", "BRA.U.ANY", "MUFU.EX2", "FFMA2", "SYNCS", "HMMA", "RED"]
print("%-14s %-58s %s" % ("object", "kernel", " ".join("%9s" % o for o in OPS)))
for f in sorted(os.listdir(OBJ)):
    if not f.endswith(".o"):
        continue
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    counts[cur][o] += 1
    for k, c in counts.items():
        if not any(c[o] for o in OPS[:8]):
            continue
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
        print("%-14s %-58s %s" % (f, name[-58:], " ".join("%9d" % c[o] for o in OPS)))
