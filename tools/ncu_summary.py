"""Summarise an .ncu-rep (read on the CPU box): key throughput metrics + top stall sites.
usage: python tools/ncu_summary.py file.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "lts__t_sector_hit_rate.pct"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")][:70], "grid", r[hdr.index("Grid Size")])
    for k in KEYS:
        if k in hdr:
            print("  %-72s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
for i, r in enumerate(rows):
    if "Source" in r and "# Samples" in r:
        h = i
        break
if h is not None:
    H = rows[h]
    si, ni = H.index("Source"), H.index("# Samples")
    stall_cols = [i for i, c in enumerate(H) if c.startswith("stall_") or "Stall" in c]
    data = []
    seen = set()
    for r in rows[h + 1:]:
        try:
            key = (r[0], r[si])
            if key in seen:
                continue
            seen.add(key)
            data.append((int(r[ni]), r[si][:100], r))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    print("== top stall sites (%d samples)" % tot)
    for n, s, r in sorted(data, key=lambda d: -d[0])[:ntop]:
        print("  %6d %5.1f%%  %s" % (n, 100.0 * n / tot, s))
