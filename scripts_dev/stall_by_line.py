"""Aggregate ncu warp-stall samples by CUDA source line.
usage: stall_by_line.py <rep.ncu-rep> <object.o> <kernel-substring> [top]
Joins `ncu --page source --csv` (SASS rows, in address order) with `nvdisasm --print-line-info` of the same cubin."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# instructions of the kernel's text section, in order, with the last seen line marker
lines, cur, inside = [], ("?", 0), False
for l in dis.splitlines():
    if l.startswith("//--------------------- .text."):
        inside = kname in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
h = rows[hi]; idx = {n: i for i, n in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) >= len(h) - 2 and r[0].startswith("0x")]
base = int(data[0][0], 16)
off2line = {o: (fl, ins) for o, fl, ins in lines}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.defaultdict(lambda: collections.Counter())
tot = 0
for r in data:
    o = int(r[0], 16) - base
    fl, _ = off2line.get(o, (("?", 0), ""))
    s = float(r[idx["# Samples"]] or 0)
    tot += s
    a = agg[fl]
    a["samples"] += s
    a["inst"] += float(r[idx["Instructions Executed"]] or 0)
    for n in stalls:
        a[n] += float(r[idx[n]] or 0)
print(f"total samples {tot:.0f}; instructions {sum(a['inst'] for a in agg.values()):.0f}; mapped lines {len(agg)}")
src_cache = {}
def src(fl):
    f, n = fl
    p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", f)
    if f not in src_cache:
        try: src_cache[f] = open(p).read().splitlines()
        except Exception: src_cache[f] = []
    L = src_cache[f]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
for fl, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    t3 = sorted(((a[n], n[6:]) for n in stalls), reverse=True)[:3]
    print(f"{100*a['samples']/tot:5.1f}%  inst {100*a['inst']/max(1,sum(x['inst'] for x in agg.values())):4.1f}%  {fl[0]}:{fl[1]:<4d} " + " ".join(f"{n}={100*v/max(a['samples'],1):.0f}%" for v, n in t3) + "  | " + src(fl))
