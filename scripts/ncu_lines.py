"""Aggregate ncu stall samples per source line: merges `ncu --page source --csv` (SASS view) with nvdisasm -g line info.
usage: ncu_lines.py REP.ncu-rep KERNEL_MANGLED_SUBSTRING [top]"""
import csv, io, os, re, subprocess, sys, tempfile, collections, glob

rep, ksub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
inst = [dict(zip(hdr, r)) for r in rows[hi + 1:] if len(r) == len(hdr)]
base = int(inst[0]["Address"], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "nonlin_b200", "libnonlin_b200.so")], cwd=tmp, capture_output=True)
lines = {}
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if ksub not in txt:
        continue
    sec = txt.split(".text." + [l for l in re.findall(r"\.text\.(\S+):", txt) if ksub in l][0] + ":")[1]
    cur = None
    for l in sec.splitlines():
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            lines[int(m.group(1), 16)] = (cur, m.group(2))
        if l.startswith("//-----") or ".section" in l:
            break
    break
agg = collections.Counter(); stall = collections.defaultdict(collections.Counter); ninst = collections.Counter()
tot = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for d in inst:
    off = int(d["Address"], 16) - base
    key = lines.get(off, (("?", 0), ""))[0]
    s = int(d["# Samples"] or 0)
    agg[key] += s; tot += s
    ninst[key] += int(d["Instructions Executed"] or 0)
    for c in stall_cols:
        v = int(d[c] or 0)
        if v: stall[key][c] += v
print("total samples", tot)
alls = collections.Counter()
for k in stall: alls.update(stall[k])
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in alls.most_common(8)))
srcs = {}
for (f, ln), s in agg.most_common(top):
    p = os.path.join(root, "nonlin_b200", "csrc", f)
    if f not in srcs and os.path.exists(p): srcs[f] = open(p).read().splitlines()
    text = srcs.get(f, [""] * 100000)[ln - 1].strip() if f in srcs and ln > 0 else ""
    st = ", ".join("%s %d%%" % (k[6:], 100 * v / max(s, 1)) for k, v in stall[(f, ln)].most_common(3))
    print("%5.1f%% %9d inst  %-24s:%-4d %-70s | %s" % (100 * s / tot, ninst[(f, ln)], f, ln, text[:70], st))
