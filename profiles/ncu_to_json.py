"""ncu launch-list CSV -> JSON summary (the profiles/rNN/ncu_launches*.json files).

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,\
lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__grid_size,launch__block_size --clock-control none --csv --log-file gpurun_out/launches.csv python profiles/one_forward.py
    python profiles/ncu_to_json.py gpurun_out/launches.csv profiles/r01/ncu_launches_vN.json [launches per forward = 47]

Keeps the launches of the LAST forward (the first one pays lazy module loading)."""
import csv
import json
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
head = rows[0]
ix = {n: head.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
launches = {}
for r in rows[1:]:
    k = int(r[ix["ID"]])
    e = launches.setdefault(k, {"kernel": re.sub(r"\(.*", "", r[ix["Kernel Name"]]).strip()})
    name, unit = r[ix["Metric Name"]], r[ix["Metric Unit"]]
    val = float(r[ix["Metric Value"]].replace(",", "")) if r[ix["Metric Value"]] not in ("", "n/a") else 0.0
    if name == "gpu__time_duration.sum":
        e["ms"] = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    elif name.startswith("dram__bytes_"):
        e["dram_%s_MB" % ("read" if "read" in name else "write")] = round(
            val * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6), 2)
    elif name.startswith("sm__pipe_tensor"):
        e["tensor_active_pct"] = round(val, 2)
    elif name.startswith("smsp__issue_active"):
        e["issue_active_pct"] = round(val, 2)
    elif name.startswith("lts__t_sector_hit_rate"):
        e["l2_hit_pct"] = round(val, 2)
    elif name == "launch__registers_per_thread":
        e["regs"] = int(val)
    elif name == "launch__grid_size":
        e["grid"] = int(val)
    elif name == "launch__block_size":
        e["block"] = int(val)
ids = sorted(launches)
names = [launches[i]["kernel"] for i in ids]
per = int(sys.argv[3]) if len(sys.argv) > 3 else 47  # launches of one forward (s4g_launch_count() per step)
last = [launches[i] for i in ids[-per:]]
total = sum(e["ms"] for e in last)
for e in last:
    e["ms"] = round(e["ms"], 4)
    e["share"] = round(e["ms"] / total, 4)
out = {"note": "ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum,... --clock-control none; one fused forward of 64 scenes "
               "x 25600 points (profiles/one_forward.py, last iteration); per-launch times are cold-cache and serialised: "
               "compare SHARES with bench.py, not absolutes", "total_ms": round(total, 3), "n_launches": len(last),
       "launches": last}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print("launches", len(last), "total ms", round(total, 3))
for e in sorted(last, key=lambda e: -e["ms"])[:12]:
    print("%8.3f ms %5.1f%%  tensor %5.1f%%  dram %8.1f MB  %s" % (e["ms"], 100 * e["share"], e.get("tensor_active_pct", 0),
          e.get("dram_read_MB", 0) + e.get("dram_write_MB", 0), e["kernel"][:70]))
