"""profiles/traffic_r2_{wave,cfg3,cfg5}.json (what bench.py reports as roofline.traffic, scaled by loci) from the launch
lists of scripts/gpu_r2_final.sh: DRAM bytes and per-kernel ms of the SECOND device-entry call of scripts/prof_wave.py.
python scripts/make_traffic_json.py TAG   (TAG = r2f: reads profiles/traffic_TAG_cfg{2,3,5}.csv)"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2f"
for cfg, name, loci in ((2, "wave", 65536), (3, "cfg3", 65536), (5, "cfg5", 16384)):
    src = "profiles/traffic_%s_cfg%d.csv" % (tag, cfg)
    if not os.path.exists(os.path.join(ROOT, src)):
        continue
    rows = [r for r in csv.reader(open(os.path.join(ROOT, src))) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, idi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault(r[idi], {"k": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
    ids = list(per)
    rd = wr = 0.0
    ms = collections.OrderedDict()
    for i in ids[len(ids) // 2:]:
        d = per[i]
        k = d["k"].split("(")[0].split("::")[-1]
        rd += d.get("dram__bytes_read.sum", 0.0)
        wr += d.get("dram__bytes_write.sum", 0.0)
        ms[k] = ms.get(k, 0.0) + d["gpu__time_duration.sum"] / 1e6
    path = os.path.join(ROOT, "profiles", "traffic_r2_%s.json" % name)
    old = json.load(open(path))
    old.update({"source": src, "loci": loci, "dram_bytes_read": rd, "dram_bytes_write": wr,
                "dram_bytes_per_locus": (rd + wr) / loci, "per_kernel_ms": [[k, round(v, 4)] for k, v in ms.items()]})
    json.dump(old, open(path, "w"), indent=1)
    print(name, "%.1f KB per locus" % ((rd + wr) / loci / 1e3), "%.3f ms" % sum(ms.values()))
