"""Blackwell evidence from the built library: per kernel, how many tcgen05 / TMEM / TMA instructions its SASS holds.

    python tools/sass_summary.py [lib.so] > profiles/sass_summary.txt

Mnemonics (B200_PROFILING.md): UTCHMMA / UTCQMMA = tcgen05.mma (kind::f16 / f8f6f4), LDTM / STTM = tcgen05.ld / st (TMEM),
UTMALDG / UTMASTG = TMA (cp.async.bulk.tensor) load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy
mma.sync, LDGSTS = cp.async, LDSM = ldmatrix.  Kernel templates are grouped by name (instantiations summed; count in brackets).
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ai-interior-image-classifier_b200", "_lib", "libiic_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "LDSM", "MUFU", "LDL", "STL"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
per = collections.OrderedDict()
cur, k = None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        full = names[k] if k < len(names) else m.group(1)
        k += 1
        base = re.sub(r"^void ", "", full).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("iic::", "")
        base = re.sub(r"<.*", "", base).split("(")[0]
        cur = per.setdefault(base, {"n": 0, "ops": collections.Counter(), "instr": 0})
        cur["n"] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur["instr"] += 1
        op = m.group(1)
        for o in OPS:
            if op.startswith(o):
                cur["ops"][o] += 1
                break
arch = re.search(r"arch = (\S+)", sass)
print(f"# {os.path.relpath(lib, ROOT)}  (cuobjdump -sass; {arch.group(1) if arch else '?'}; {len(per)} kernel templates)")
print(f"{'kernel [instantiations]':44s} {'SASS':>8s} " + " ".join(f"{o:>8s}" for o in OPS) + "   tensor path")
for name, d in sorted(per.items(), key=lambda kv: -kv[1]["instr"]):
    o = d["ops"]
    kind = "tcgen05 + TMEM" + (" + TMA" if o["UTMALDG"] or o["UTMASTG"] else "") if o["UTCHMMA"] else ("mma.sync (legacy)" if o["HMMA"] else "-")
    print(f"{(name + ' [' + str(d['n']) + ']'):44s} {d['instr']:8d} " + " ".join(f"{o[x]:8d}" for x in OPS) + f"   {kind}")
