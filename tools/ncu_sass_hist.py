"""Per-opcode executed-instruction histogram + hottest instructions of an ncu --import-source capture (SASS view):
    python tools/ncu_sass_hist.py rep.ncu-rep [top_n]"""
import collections, csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
b = out.split('"Kernel Name"')[1]
lines = b.split("\n")
print("Kernel", lines[0][:100])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
iS, iSrc, iX = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = [r for r in rows[1:] if len(r) == len(hdr)]
tot_x = sum(int(r[iX]) for r in data)
tot_s = sum(int(r[iS]) for r in data)
ops = collections.Counter()
samp = collections.Counter()
for r in data:
    toks = r[iSrc].strip().split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    ops[op] += int(r[iX])
    samp[op] += int(r[iS])
print(f"static instructions {len(data)}, executed warp-instructions {tot_x}, samples {tot_s}")
for op, c in ops.most_common(n):
    print(f"  {op:12s} exec {c:12d} {100*c/tot_x:5.1f}%   samples {100*samp[op]/max(tot_s,1):5.1f}%")
