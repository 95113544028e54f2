"""Dynamic SASS instruction mix (per thread) from an ncu report's source page."""
import csv, subprocess, sys
from collections import Counter
rep, cells = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows):
    if "Source" in r and "Instructions Executed" in r:
        hdr, start = r, i + 1
        break
si, ei, ti = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
c, tot = Counter(), 0
for r in rows[start:]:
    if len(r) <= ti:
        continue
    try:
        n = int(r[ti])
    except ValueError:
        continue
    toks = r[si].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    c[op.split(".")[0]] += n
    tot += n
print("thread-instr per cell-update: %.1f" % (tot / cells))
for op, n in c.most_common(30):
    print("  %-8s %7.1f" % (op, n / cells))
