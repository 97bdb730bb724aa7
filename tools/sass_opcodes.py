"""Opcode evidence from the shipped cubin (no GPU needed): per kernel, how often the SASS mnemonics that prove the tensor-core /
TMA / mbarrier paths occur.   python tools/sass_opcodes.py > profiles/sass_opcodes_rNN.txt"""
import collections, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "rabitq_b200", "librabitq_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["IMMA", "HMMA", "UTC", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "POPC", "LOP3", "FFMA2", "FADD2", "FFMA", "I2FP", "VOTE", "REDUX", "NANOSLEEP"]
per, cur = collections.OrderedDict(), None
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                per[cur][op if w in ("IMMA", "HMMA", "UTC", "UBLKCP", "SYNCS") else w] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, root)} ({os.path.getsize(so)} bytes): watched SASS mnemonics per kernel")
tot = collections.Counter()
for k, c in per.items():
    tot.update(c)
    items = ", ".join(f"{op} {n}" for op, n in sorted(c.items()) if op != "_total")
    print(f"{k:70s} instr {c['_total']:6d}  {items}")
print("TOTAL", dict(sorted(tot.items())))
