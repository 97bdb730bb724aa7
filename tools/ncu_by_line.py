"""Executed warp-instructions and stall samples per CUDA source line of one kernel: joins the SASS rows of an ncu
--import-source capture (in address order) with `nvdisasm -g` line annotations of the same cubin.
    python tools/ncu_by_line.py rep.ncu-rep 'scan_mma_kernelILi4ELb0' [min_pct]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, pat = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "rabitq_b200", "librabitq_b200.so")], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
lines_of = []  # per instruction: source line
on, cur = False, 0
for l in sass:
    if l.startswith(".text."):
        on = pat in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith("kernels.cuh") else -1
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines_of.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
b = out.split('"Kernel Name"')[1]
rows = list(csv.reader(io.StringIO("\n".join(b.split("\n")[1:]))))
hdr = rows[0]
iS, iX = hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows[1:] if len(r) == len(hdr)]
assert len(data) == len(lines_of), (len(data), len(lines_of))
ex, sm = collections.Counter(), collections.Counter()
for r, ln in zip(data, lines_of):
    ex[ln] += int(r[iX]); sm[ln] += int(r[iS])
tx, ts = sum(ex.values()), sum(sm.values())
src = open(os.path.join(root, "rabitq_b200", "csrc", "kernels.cuh")).read().split("\n")
print(f"executed {tx}  samples {ts}")
for ln in sorted(ex):
    if 100 * ex[ln] / tx >= min_pct or 100 * sm[ln] / ts >= min_pct:
        print(f"{ln:5d} exec {100*ex[ln]/tx:5.1f}% samp {100*sm[ln]/ts:5.1f}%  {src[ln-1].strip()[:100] if ln > 0 else '(other file)'}")
