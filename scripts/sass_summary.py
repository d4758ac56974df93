"""Opcode histogram per kernel of ur-mvo_b200/lib/liburmvo_b200.so (cuobjdump -sass), written to
profiles/sass_summary_rNN.txt: evidence of what the kernels are made of (DFMA / LDGSTS / RED / ATOMS /
no UTCMMA / no UTMALDG — the path has no dense contraction, BASELINE.json north_star)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "ur-mvo_b200", "lib", "liburmvo_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
hist, cur = collections.OrderedDict(), None
for line in out.split("\n"):
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").split("(")[0]
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
keys = ["DFMA", "DMUL", "DADD", "FFMA", "MUFU", "LDG", "STG", "LDS", "STS", "LDGSTS", "REDG", "ATOM", "ATOMS", "SHFL", "BAR", "UTCMMA", "UTMALDG", "HMMA"]
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary_r02.txt")
with open(dst, "w") as f:
    f.write("kernel".ljust(64) + "total " + " ".join(k.rjust(7) for k in keys) + "\n")
    for k, c in hist.items():
        tot = sum(c.values())
        row = [sum(v for op, v in c.items() if op == key or op.startswith(key + "_")) if key not in ("LDG", "STG", "LDS", "STS", "ATOM") else
               sum(v for op, v in c.items() if op == key) for key in keys]
        f.write(k[-63:].ljust(64) + f"{tot:6d} " + " ".join(f"{v:7d}" for v in row) + "\n")
print("wrote", dst, "kernels", len(hist))
