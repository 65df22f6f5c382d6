"""Summaries of an `ncu --set full --import-source on` capture for profiles/: key metrics (CSV) and where the warp
instructions and the stall samples go per source function.   python scripts/ncu_summary.py X.ncu-rep OUT_PREFIX"""
import collections, csv, io, os, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
KEYS = r"^(gpu__time_duration.sum|dram__bytes_(read|write).sum|launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_(registers|shared_mem|warps))|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct_of_peak_sustained_active|sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_(fp64|alu|fma|lsu|xu).avg.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed.ratio|smsp__inst_executed.sum|smsp__warps_eligible.avg.per_cycle_active|l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|sass__inst_executed_local_(loads|stores)|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|smsp__average_warps_issue_stalled_\w+_per_issue_active.ratio)$"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
with open(out + "_summary.csv", "w") as f:
    f.write("metric,unit,value\n")
    f.write('Kernel Name,,"%s"\n' % r[hdr.index("Kernel Name")])
    for i, k in enumerate(hdr):
        if re.match(KEYS, k) and r[i] not in ("", "0"):
            f.write("%s,%s,%s\n" % (k, units[i], r[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, h, data = None, None, []
for row in csv.reader(io.StringIO(src)):
    if len(row) == 2 and row[0] == "File Path":
        cur = os.path.basename(row[1]); continue
    if row and row[0] == "Line No":
        h = row; continue
    if h and len(row) == len(h):
        try:
            data.append((cur, int(row[0]), int(row[7]), int(row[8]), int(row[6])))
        except ValueError:
            pass
funcs = {}
csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "varlociraptor_b200", "csrc")
for fn in set(d[0] for d in data):
    p = os.path.join(csrc, fn or "")
    if os.path.exists(p):
        funcs[fn] = [(i, m.group(1)) for i, l in enumerate(open(p), 1)
                     for m in [re.match(r"^(?:template.*>\s*)?(?:VLR_\w+|static|__global__|inline|__device__)[\w\s\*&:<>,]*?\b(\w+)\s*\(", l)] if m]
def fname(f, ln):
    name = "(other)"
    for i, n in funcs.get(f, []):
        if i <= ln: name = n
        else: break
    return name
agg = collections.defaultdict(lambda: [0, 0, 0])
for f, ln, ie, te, sm in data:
    k = (f, fname(f, ln)); agg[k][0] += ie; agg[k][1] += te; agg[k][2] += sm
ti, ts = sum(v[0] for v in agg.values()), sum(v[2] for v in agg.values())
with open(out + "_functions.csv", "w") as f:
    f.write("file,function,warp_instructions_pct,active_lanes_per_instruction,stall_samples_pct\n")
    f.write("(all),(all),100,%.2f,100\n" % (sum(v[1] for v in agg.values()) / max(1, ti)))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        if v[0] and (100.0 * v[0] / ti >= 0.2 or 100.0 * v[2] / ts >= 0.2):
            f.write("%s,%s,%.2f,%.1f,%.2f\n" % (k[0], k[1], 100.0 * v[0] / ti, v[1] / v[0], 100.0 * v[2] / ts))
print(open(out + "_functions.csv").read())
