"""Turn the ncu captures of tools/profile_round.sh into the tracked summaries under profiles/.
    python tools/summarize_profiles.py <tag> <name>     e.g.  r01v6 r01_v6
writes profiles/<name>_launches.csv (copy), <name>_launch_summary.txt, <name>_gemm_ncu.json, <name>_attention_ncu.json and
profiles/roofline_traffic.json (what bench.py reports as roofline.traffic)."""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, name = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list ----
src = os.path.join(G, f"{tag}_launches.csv")
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
iK, iM, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iU = hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    if r is hdr or r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iU], 1e-6)
    k = re.sub(r"\(.*", "", r[iK])[:110]
    tot[k] += v; cnt[k] += 1
shutil.copy(src, os.path.join(P, f"{name}_launches.csv"))
allms = sum(tot.values())
with open(os.path.join(P, f"{name}_launch_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 1200  python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
    f.write("# Per-launch times are cold-cache and serialised: compare SHARES with bench.py's roofline.share_of_step\n")
    for k, v in tot.most_common():
        f.write(f"{v:9.2f} ms {cnt[k]:5d} launches  avg {v / cnt[k]:7.3f} ms  {100 * v / allms:5.1f}%  {k}\n")
print(open(os.path.join(P, f"{name}_launch_summary.txt")).read())

# ---- full captures ----
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__warps_active.avg.per_cycle_active"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def read_rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u = rr[0], rr[1]
    out = []
    for vals in rr[2:]:
        d = {"kernel": vals[h.index("Kernel Name")][:120]}
        for k in KEEP:
            if k in h:
                i = h.index(k)
                try:
                    x = float(vals[i].replace(",", ""))
                except ValueError:
                    continue
                if u[i] in UNIT:
                    x *= UNIT[u[i]]
                elif k == "gpu__time_duration.sum":
                    x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u[i], 1.0)   # -> microseconds
                d[k] = x
        out.append(d)
    return out


gem = read_rep(os.path.join(G, f"{tag}_gemm_full.ncu-rep"))
def gemm_shape(d):
    # identify the launch by its epilogue template argument and its DRAM read volume (c_proj reads the [M, 4d] operand)
    m = re.search(r"gemm_bf16_tn_kernel<\(?(?:int\))?(\d+), \(?(?:int\))?(\d+), \(?(?:int\))?(\d+)", d["kernel"])
    epi = int(m.group(3)) if m else -1
    if epi == 0: return "qkv (N=2304,K=768, bias)"
    if epi == 1: return "c_fc (N=3072,K=768, QuickGELU + LoRA + fused down-proj)"
    if epi == 5: return "c_proj (N=768,K=3072, +residual f32 + LoRA)"
    if epi == 2: return "c_proj (N=768,K=3072, +residual f32 + LoRA)" if d.get("dram__bytes_read.sum", 0) > 1.5e9 else "out_proj (N=768,K=768, +residual f32)"
    return "patch embedding" if epi == 3 else "?"


for d in gem:
    d["shape"] = gemm_shape(d)
    d["dram_bytes"] = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
json.dump(gem, open(os.path.join(P, f"{name}_gemm_ncu.json"), "w"), indent=1)
att = read_rep(os.path.join(G, f"{tag}_attn_full.ncu-rep"))
for d in att:
    d["dram_bytes"] = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
json.dump(att, open(os.path.join(P, f"{name}_attention_ncu.json"), "w"), indent=1)
per_launch = sum(d["dram_bytes"] for d in gem) / max(len(gem), 1)
json.dump({"gemm_dram_bytes_per_launch": per_launch,
           "how": "mean of dram__bytes_read.sum + dram__bytes_write.sum over the four GEMM launches of one transformer block "
                  f"(qkv, out_proj, c_fc, c_proj; batch 1024) from one `ncu --set full` capture of bench.py ({name}_gemm_ncu.json)",
           "per_shape": {d["shape"]: d["dram_bytes"] for d in gem}},
          open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
for d in gem + att:
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()})
