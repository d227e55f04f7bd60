"""Per-kernel DRAM traffic of one training step from an ncu launch list
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`), trimmed to the timed
step [first ingest_kernel .. the repack_batch_kernel after the fused Adam]; bench.py reads the result for
`roofline.traffic`.   python scripts/ncu_traffic_json.py launches.csv profiles/r02_traffic.json"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iM, iV, iU, iID = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
ids = {}
for r in rows[1:]:
    ids.setdefault(int(r[iID]), r[iK])
starts = [i for i in sorted(ids) if "ingest_kernel" in ids[i]]
lo = starts[0]
ends = [i for i in sorted(ids) if "repack_batch_kernel" in ids[i] and i > lo]
hi = ends[0] if ends else max(ids)
per = collections.defaultdict(lambda: {"ids": set(), "ms": 0.0, "bytes": 0.0})
for r in rows[1:]:
    if not (lo <= int(r[iID]) <= hi):
        continue
    k = re.sub(r"\(.*$", "", r[iK]).replace("void ", "").strip()
    v, u = float(r[iV].replace(",", "")), r[iU]
    p = per[k]
    p["ids"].add(r[iID])
    if r[iM] == "gpu__time_duration.sum":
        p["ms"] += v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    else:
        p["bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def entry(keys):
    n = sum(len(per[k]["ids"]) for k in keys)
    return {"launches": n, "ms_per_launch": sum(per[k]["ms"] for k in keys) / max(n, 1),
            "dram_bytes_per_launch": sum(per[k]["bytes"] for k in keys) / max(n, 1)}


out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on "
                 "`python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-infer` (B=64, 1xB200), trimmed to the timed "
                 "step; per-launch means (scripts/ncu_traffic_json.py)",
       "step_launches": sum(len(p["ids"]) for p in per.values()), "step_ms_under_ncu": sum(p["ms"] for p in per.values()),
       "kernels": {k: entry([k]) for k in per}}
for fam in ("conv3x3_halo_kernel", "conv3x3_wgrad_kernel"):
    out[fam + "_all"] = entry([k for k in per if fam in k])
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(f"{sys.argv[2]}: {out['step_launches']} launches, {out['step_ms_under_ncu']:.2f} ms")
