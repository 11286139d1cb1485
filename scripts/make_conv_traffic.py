"""profiles/conv_traffic.json from an `ncu --set full` capture of ONE forward (scripts/profile_forward.py <capacity>), exported on
the GPU box with `ncu -i <rep> --page raw --csv`: DRAM bytes (read + write) summed over the conv_tc launches, stamped with the
fingerprint of the kernel sources so that bench.py refuses the number once the kernels change.
  python scripts/make_conv_traffic.py profiles/r02_fwd256_ncu_raw.csv 256
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

path, cap = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
iname, ird, iwr, it, itp = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
                                                    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
tot = us = tp = n = 0
for r in rows[2:]:
    if "conv_tc" not in r[iname]:
        continue
    tot += float(r[ird]) * scale[units[ird]] + float(r[iwr]) * scale[units[iwr]]
    t = float(r[it]) * {"ms": 1e3, "us": 1.0, "ns": 1e-3}.get(units[it], 1e3)
    us += t
    tp += t * float(r[itp])
    n += 1
out = {"capacity": cap, "dram_bytes_per_forward": tot, "conv_launches": n, "conv_us_under_ncu": us,
       "conv_tensor_pipe_active_pct_time_weighted": tp / us, "csrc_sha": bench.conv_sources_sha(),
       "source": os.path.relpath(path, ROOT) + ": sum of dram__bytes_read.sum + dram__bytes_write.sum over the %d conv_tc launches of one %d-crop forward" % (n, cap)}
json.dump(out, open(os.path.join(ROOT, "profiles", "conv_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
