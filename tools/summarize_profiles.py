"""Turn the ncu outputs of tools/profile_round.sh (gpurun_out/) into the committed summaries under
profiles/: a per-kernel launch table with durations and DRAM traffic, and the key metrics of the
--set full captures.  usage: python tools/summarize_profiles.py r01"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
TAGS = sys.argv[2:] or ["gemm_big", "gemm_small", "attn", "ln"]
TRAIN = "train" in R
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list
lines = [l for l in open(os.path.join(GO, R + "_launches.csv")) if not l.startswith("==")]
per = collections.OrderedDict()
for x in csv.DictReader(lines):
    k = (x["ID"], x["Kernel Name"], x["Grid Size"])
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    d = per.setdefault(k, {})
    if x["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u.startswith("ms") else v)
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d[x["Metric Name"]] = v * mult
agg = collections.OrderedDict()
for (_, name, grid), d in per.items():
    short = name.split("(")[0].replace("void ", "").replace("mtn::", "")
    a = agg.setdefault(short, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"] += 1; a["us"] += d.get("us", 0); a["rd"] += d.get("dram__bytes_read.sum", 0); a["wr"] += d.get("dram__bytes_write.sum", 0)
tot = sum(a["us"] for a in agg.values())
out = ["# %s: ncu launch list of one %s (tools/%s: B=32, T=256; cold cache, serialised)"
       % (R, "eager TRAINING step (forward + loss + backward + Adam, dropout 0.1)" if TRAIN else "bench step",
          "profile_train_step.py" if TRAIN else "profile_step.py"), "",
       "| kernel | launches | total us | share | avg us | DRAM read MB | DRAM write MB |", "|---|---:|---:|---:|---:|---:|---:|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    out.append("| `%s` | %d | %.1f | %.1f%% | %.2f | %.1f | %.1f |" % (k[:70], a["n"], a["us"], 100 * a["us"] / tot, a["us"] / a["n"],
                                                                       a["rd"] / 1e6, a["wr"] / 1e6))
out.append("| **total** | %d | %.1f | | | | |" % (sum(a["n"] for a in agg.values()), tot))
ours = {k: a for k, a in agg.items() if not (k.startswith("at::") or "at::native" in k or k.startswith("nccl"))}
out += ["", "Kernels of this repo: %d launches, %.1f us (%.1f%% of the step's kernel time); the rest are PyTorch glue "
        "(mask construction in Batch, clone of the residual stream)." % (sum(a["n"] for a in ours.values()),
                                                                         sum(a["us"] for a in ours.values()),
                                                                         100 * sum(a["us"] for a in ours.values()) / tot)]
open(os.path.join(PR, R + "_launches.md"), "w").write("\n".join(out) + "\n")
lin = [a for k, a in agg.items() if "gemm_f16_tc" in k]
traffic = {"kernel": "gemm_f16_tc_kernel (all linear launches of one step)", "launches": sum(a["n"] for a in lin),
           "dram_bytes_read": sum(a["rd"] for a in lin), "dram_bytes_write": sum(a["wr"] for a in lin),
           "source": "profiles/%s_launches.md (ncu dram__bytes_read.sum + dram__bytes_write.sum, cold cache)" % R}
json.dump(traffic, open(os.path.join(PR, R + "_gemm_traffic.json"), "w"), indent=1)

# ---- full captures
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
out = ["# %s: ncu --set full captures (key metrics; .ncu-rep files are scratch under gpurun_out/)" % R, ""]
for tag in TAGS:
    rep = os.path.join(GO, "%s_%s.ncu-rep" % (R, tag))
    if not os.path.isfile(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        out.append("## %s: `%s` grid %s" % (tag, r[H.index("Kernel Name")][:80], r[H.index("Grid Size")]))
        for k in KEYS:
            if k in H:
                out.append("- %s = %s %s" % (k, r[H.index(k)], U[H.index(k)]))
        out.append("")
open(os.path.join(PR, R + "_ncu_full.md"), "w").write("\n".join(out) + "\n")
print(open(os.path.join(PR, R + "_launches.md")).read())
