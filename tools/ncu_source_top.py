"""Top stall sites of a kernel from an ncu --import-source capture:  python tools/ncu_source_top.py rep.ncu-rep [n]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
blocks = out.split('"Kernel Name"')
for b in blocks[int(sys.argv[3]) if len(sys.argv) > 3 else 1:][:1]:
    lines = b.split("\n")
    print("Kernel", lines[0][:100])
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    iS, iSrc, iX = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[1:] if len(r) == len(hdr)]
    tot = sum(int(r[iS]) for r in data)
    print("total samples", tot, "instructions", len(data))
    for r in sorted(data, key=lambda r: -int(r[iS]))[:n]:
        st = sorted(((int(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
        print(f"{int(r[iS]):7d} {100*int(r[iS])/tot:5.1f}%  exec={r[iX]:>9s}  {r[iSrc].strip()[:70]:70s} {st}")
