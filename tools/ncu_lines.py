"""Attribute an .ncu-rep's per-SASS samples to CUDA source lines by aligning with nvdisasm line info of the built object.
Usage: ncu_lines.py rep object.o kernel_substr [topn]"""
import csv, collections, os, re, subprocess, sys, tempfile
rep, obj, ksub = sys.argv[1:4]; topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
skip = sys.argv[5] if len(sys.argv) > 5 else "0"   # which launch inside the report
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of = []; cur = None; infn = False
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", l)
    if m:
        infn = ksub in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        line_of.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# one section per profiled launch: ["Kernel Name", name], header row, instruction rows
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sel = int(skip)
print("launches in report:", [rows[i][1][:60] for i in starts], "-> using", sel)
rows = rows[starts[sel]:(starts[sel + 1] if sel + 1 < len(starts) else len(rows))]
h = rows[1]; iN = h.index("# Samples"); iI = h.index("Instructions Executed"); iS = h.index("Source")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
ins = [r for r in rows[2:] if len(r) > iI]
print(f"sass rows ncu={len(ins)} nvdisasm={len(line_of)}")
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for k, r in enumerate(ins):
    key = line_of[k] if k < len(line_of) else None
    a = agg[key]
    a[0] += int(r[iN] or 0); a[1] += int(r[iI] or 0)
    for i in stall_cols:
        if r[i] not in ("", "0"): a[2][h[i][6:]] += int(r[i])
tot = sum(a[0] for a in agg.values())
files = {}
def text(key):
    if not key: return ""
    f, n = key
    if f not in files:
        for root in ("ai-interior-image-classifier_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p): files[f] = open(p).read().splitlines(); break
        else: files[f] = []
    return files[f][n - 1].strip()[:90] if 0 < n <= len(files[f]) else ""
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    st = ", ".join(f"{k}:{v}" for k, v in a[2].most_common(3))
    print(f"{a[0]:6d} {100*a[0]/max(tot,1):5.1f}%  inst {a[1]:10d}  {str(key):32s} {text(key)}   [{st}]")
