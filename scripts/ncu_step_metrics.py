"""Turns an `ncu --csv --metrics ...` log of `python bench.py --profile` (or scripts/profile_forward.py) into a per-kernel
table: launches, total time, share, DRAM bytes, achieved HBM GB/s, tensor-pipe active %.

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed \
      --clock-control none --csv --log-file gpurun_out/step_metrics.csv python bench.py --profile
  python scripts/ncu_step_metrics.py gpurun_out/step_metrics.csv [second_half_only=1] > profiles/<name>.md
"""
import collections
import csv
import json
import sys

path = sys.argv[1]
second_half = len(sys.argv) < 3 or sys.argv[2] != "0"
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
by_id = collections.OrderedDict()
for r in rows:
    d = by_id.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
    v = r["Metric Value"].replace(",", "")
    try:
        v = float(v)
    except ValueError:
        continue
    unit = r["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[r["Metric Name"]] = v * scale
launches = list(by_id.values())
if second_half:
    launches = launches[len(launches) // 2:]          # bench.py --profile = warm-up step + measured step


def short(n):
    n = n.split("(")[0].replace("void ", "").replace("p2p::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return n.strip()


agg = collections.OrderedDict()
tot = 0.0
for l in launches:
    k = short(l["name"])
    a = agg.setdefault(k, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tp": 0.0, "sm": 0.0})
    t = l.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1
    a["us"] += t
    a["rd"] += l.get("dram__bytes_read.sum", 0.0)
    a["wr"] += l.get("dram__bytes_write.sum", 0.0)
    a["tp"] += t * l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    a["sm"] += t * l.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
    tot += t
print("| kernel | launches | total us | share | DRAM read MB | DRAM write MB | HBM GB/s | tensor-pipe active % (time-weighted) | SM throughput % |")
print("|---|---|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
    print("| %s | %d | %.1f | %.1f %% | %.1f | %.1f | %.0f | %.1f | %.1f |" % (k, a["n"], a["us"], 100 * a["us"] / tot, a["rd"] / 1e6, a["wr"] / 1e6, gbs,
                                                                     a["tp"] / a["us"] if a["us"] else 0, a["sm"] / a["us"] if a["us"] else 0))
print("\ntotal %.1f us over %d launches" % (tot, len(launches)))
conv = [a for k, a in agg.items() if k.startswith("conv_tc")]
if conv:
    us = sum(a["us"] for a in conv)
    by = sum(a["rd"] + a["wr"] for a in conv)
    tp = sum(a["tp"] for a in conv) / us
    print("\nconv_tc kernels together: %.1f us, DRAM %.1f MB, tensor-pipe active %.1f %% (time-weighted)" % (us, by / 1e6, tp))
    print("JSON " + json.dumps({"conv_us": us, "conv_dram_bytes": by, "conv_tensor_pipe_pct": tp, "launches": sum(a["n"] for a in conv)}))
