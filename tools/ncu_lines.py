"""Attribute ncu SASS-level samples / executed instructions to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <nvdisasm -g -c output> <mangled-kernel-substring> [file-filter]"""
import csv, re, subprocess, sys
from collections import defaultdict
rep, dis, kern = sys.argv[1:4]
ffilter = sys.argv[4] if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
sass = [(r[1].strip(), int(r[iS] or 0), int(r[iI] or 0)) for r in rows[h + 1:] if len(r) > iI and r[0].startswith("0x")]
# disassembly: instructions of the kernel in order with their source line
lines = open(dis).read().split("\n")
infn = False
cur = ("?", 0)
seq = []
for l in lines:
    if l.startswith(".text."):
        infn = kern in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        seq.append(cur)
print(f"sass rows {len(sass)}, disasm instrs {len(seq)}")
n = min(len(sass), len(seq))
S, I = defaultdict(int), defaultdict(int)
for (txt, s, i), key in zip(sass[:n], seq[:n]):
    S[key] += s; I[key] += i
totS, totI = sum(S.values()), sum(I.values())
print(f"total samples {totS}, warp instructions {totI}")
items = sorted(S.items(), key=lambda kv: -kv[1])
for (f, ln), s in items[:45]:
    if ffilter and ffilter not in f: continue
    print(f"{f}:{ln:4d}  samples {100*s/totS:5.1f}%  instr {100*I[(f,ln)]/totI:5.1f}%")
