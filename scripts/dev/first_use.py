"""Development: for every 128-bit global load of a kernel print how many instructions later its
destination registers are first read (checks that software-pipelined loads are not consumed early).
Usage: first_use.py <nvdisasm -g -c dump> <kernel name substring>"""
import re, sys
lines = open(sys.argv[1]).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and sys.argv[2] in l)
seq = []
for l in lines[start + 1:]:
    if l.startswith("//-----"): break
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((int(m.group(1), 16), m.group(2)))
for k, (off, txt) in enumerate(seq):
    m = re.search(r"LDG\.E\.128 (R\d+)", txt)
    if not m: continue
    r = int(m.group(1)[1:]); regs = {f"R{r + i}" for i in range(4)}
    first = None
    for j in range(k + 1, min(k + 600, len(seq))):
        body = seq[j][1]
        ops = re.findall(r"R\d+", body.split(",", 1)[1] if "," in body else "")
        if regs & set(ops): first = (j - k, body[:60]); break
    print(f"{off:6x} {txt[:64]} first use: {first}")
