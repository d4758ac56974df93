"""Development: attribute the warp-stall samples of an ncu report (--import-source on) to source lines.
Usage: ncu_lines.py <report.ncu-rep> <kernel mangled-name substring> [top]
Needs the same build of liburmvo_b200.so that was profiled (line table comes from nvdisasm -g)."""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "ur-mvo_b200/lib/liburmvo_b200.so")], cwd=tmp, capture_output=True)
unit = os.environ.get("UNIT", "ba_kernels")  # translation unit (cubin name prefix / source file) of the kernel
cubin = [f for f in os.listdir(tmp) if f.startswith(unit)][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
seq, cur = [], None
for l in dis[start + 1:]:
    if l.startswith("//-----") : break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = int(m.group(2)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((int(m.group(1), 16), cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + os.environ.get("KREGEX", kname)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
for k, r in enumerate(data):  # keep the first kernel section only
    if r and r[0] == "Kernel Name":
        data = data[:k]
        break
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stallcols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
assert len(data) == len(seq), (len(data), len(seq))
byline, exline, bystall, tot = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
for d, (off, ln, txt) in zip(data, seq):
    ln = ln or -1
    byline[ln] += int(d[isamp]); exline[ln] += int(d[iex])
    for i in stallcols:
        bystall[ln][hdr[i]] += int(d[i]); tot[hdr[i]] += int(d[i])
n = sum(byline.values())
src = open(os.path.join(ROOT, f"ur-mvo_b200/csrc/{unit}.cu")).read().split("\n")
print("samples", n, "warp-instructions %.3f G" % (sum(exline.values()) / 1e9))
print(" ".join(f"{k[6:]}={v / sum(tot.values()) * 100:.1f}%" for k, v in tot.most_common(8)))
for ln, c in byline.most_common(top):
    t3 = ", ".join(f"{k[6:]}:{v}" for k, v in bystall[ln].most_common(3))
    print(f"{ln:5d} {c / n * 100:5.1f}% ex={exline[ln]:9d} [{t3}] {src[ln - 1].strip()[:90] if 0 < ln <= len(src) else ''}")
if os.environ.get("BY_EXEC"):  # short kernels have too few samples: rank the lines by executed warp instructions
    ne = sum(exline.values())
    print("by executed warp instructions:")
    for ln, c in exline.most_common(top):
        print(f"{ln:5d} {c / ne * 100:5.1f}% ex={c:9d} {src[ln - 1].strip()[:100] if 0 < ln <= len(src) else ''}")
if len(sys.argv) > 4:  # dump the SASS of the given source lines with their samples
    want = {int(x) for x in sys.argv[4].split(",")}
    for k, (d, (off, ln, txt)) in enumerate(zip(data, seq)):
        if ln in want and int(d[isamp]) > 50:
            ctx = " | ".join(seq[j][2][:40] for j in range(max(0, k - 3), k))
            print(f"{ln:5d} {off:6x} samples={d[isamp]:>6s} ex={d[iex]:>9s} {txt[:70]}    <- {ctx}")
if os.environ.get("REGIONS"):  # REGIONS="name:lo-hi,name:lo-hi": executed warp instructions / samples per line range
    tot_ex = sum(exline.values())
    for r in os.environ["REGIONS"].split(","):
        name, rng = r.split(":"); lo, hi = (int(x) for x in rng.split("-"))
        ex = sum(v for l, v in exline.items() if lo <= l <= hi); sm = sum(v for l, v in byline.items() if lo <= l <= hi)
        print(f"{name:14s} lines {lo}-{hi}: instr {ex / tot_ex * 100:5.1f}%  samples {sm / n * 100:5.1f}%")
