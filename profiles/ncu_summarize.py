"""`ncu --set full` report -> per-kernel JSON + text summary (run HERE, no GPU needed: reads the .ncu-rep).

    python profiles/ncu_summarize.py gpurun_out/geom.ncu-rep profiles/r02/ncu_geometry   # writes .json and .txt

Per profiled launch: duration, DRAM bytes / % of peak, L2 hit rate, issue-slot utilisation (the roofline fraction of an
issue-bound kernel), achieved occupancy, FMA-pipe utilisation, tensor-pipe %, registers / grid / block, and the
warp-stall mix (cycles a warp spends stalled per issued instruction, by reason: WarpStateStats section)."""
import csv
import io
import json
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(head)}
WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed.avg.per_cycle_active": "ipc_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_inst_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_size": "cluster",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "smsp__cycles_active.avg": "sm_active_cycles",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3,
         "nsecond": 1e-6, "msecond": 1.0, "second": 1e3}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


kernels = []
for r in rows[2:]:
    if len(r) < len(head):
        continue
    k = {"id": int(r[col["ID"]]), "kernel": re.sub(r"\(.*", "", r[col["Kernel Name"]]).strip()}
    for metric, short in WANT.items():
        if metric in col:
            v = num(r[col[metric]])
            if v is None:
                continue
            u = units[col[metric]]
            if short == "duration":
                k["ms"] = round(v * SCALE.get(u, 1e-6), 4)
            elif short.startswith("dram_r") or short.startswith("dram_w"):
                k[short + "_MB"] = round(v * SCALE.get(u, 1.0) / 1e6, 2)
            else:
                k[short] = round(v, 2)
    stalls = {}
    for name, i in col.items():
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", name)
        if m and "not_issued" not in name:
            v = num(r[i])
            if v:
                stalls[m.group(1)] = v
    tot = sum(stalls.values())
    if tot > 0:
        k["stall_mix_pct"] = {n: round(100 * v / tot, 1) for n, v in sorted(stalls.items(), key=lambda x: -x[1])[:6]}
    kernels.append(k)

json.dump({"report": rep, "note": "ncu --set full --clock-control none; per-launch, cold-cache, serialised", "kernels": kernels},
          open(out + ".json", "w"), indent=1)
with open(out + ".txt", "w") as f:
    for k in kernels:
        f.write("%s\n" % k["kernel"])
        for key in ("ms", "grid", "block", "cluster", "regs", "dram_read_MB", "dram_write_MB", "dram_pct_of_peak", "l2_hit_pct",
                    "l1_hit_pct", "issue_active_pct", "ipc_active", "achieved_occupancy_pct", "fma_pipe_pct", "fma_inst_pct",
                    "tensor_pipe_pct", "sm_throughput_pct"):
            if key in k:
                f.write("    %-26s %s\n" % (key, k[key]))
        if "stall_mix_pct" in k:
            f.write("    stall mix (%% of stalled warp-cycles): %s\n" % ", ".join("%s %.1f" % kv for kv in k["stall_mix_pct"].items()))
        f.write("\n")
print(open(out + ".txt").read())
