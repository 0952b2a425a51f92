"""Turn gpurun_out ncu artefacts into the small text/JSON summaries committed under profiles/.
usage: summarise_profile.py <round tag> <launch csv> <ncu-rep> [kernel filter]"""
import collections, csv, json, subprocess, sys
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
d = collections.defaultdict(list)
for r in rows:
    d[r[4]].append(float(r[-1]))
total = sum(sum(v) for v in d.values())
lines = [f"# ncu launch list ({launches}); durations are cold-cache and serialised: compare SHARES", ""]
lines.append(f"{'kernel':90s} {'launches':>8s} {'median us':>10s} {'share %':>8s}")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    v2 = sorted(v)
    lines.append(f"{k[:90]:90s} {len(v):8d} {v2[len(v2)//2]/1e3:10.2f} {100*sum(v)/total:8.1f}")
open(f"profiles/{tag}_launches.txt", "w").write("\n".join(lines) + "\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
out = [f"# ncu --set full --clock-control none ({rep})", ""]
traffic = None
for r in rr[2:]:
    dd = dict(zip(hdr, r))
    out.append(dd["Kernel Name"])
    for k in keys:
        if k in dd:
            out.append(f"  {k:90s} {dd[k]:>16s} {units[hdr.index(k)]}")
    def tobytes(k):
        v = float(dd[k].replace(",", "")); u = units[hdr.index(k)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    traffic = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
    out.append(f"  dram traffic per launch (read+write): {traffic/1e6:.2f} MB")
    out.append("")
open(f"profiles/{tag}_fused_kernel.txt", "w").write("\n".join(out) + "\n")
json.dump({"dram_bytes_per_launch": traffic, "source": rep, "note": "dram__bytes_read.sum + dram__bytes_write.sum of one "
           "isolated launch; the 42 MB of results are still dirty in the 126 MB L2 when the kernel ends, so their "
           "write-back is not inside this window"}, open("profiles/traffic.json", "w"), indent=1)
print(open(f"profiles/{tag}_launches.txt").read()); print(open(f"profiles/{tag}_fused_kernel.txt").read())
