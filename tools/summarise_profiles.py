"""Turns the ncu outputs of tools/profile_bench.sh (gpurun_out/) into the committed summaries under profiles/.
usage: python tools/summarise_profiles.py <round tag, e.g. r1> <workload, e.g. cfg2>"""
import csv, json, os, subprocess, sys
from collections import defaultdict

tag, w = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio"]

# launch list -> shares
rows = list(csv.reader(open(os.path.join(root, "gpurun_out", f"launches_{w}.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
with open(os.path.join(out, f"{tag}_{w}_launches_ncu.csv"), "w") as f:
    csv.writer(f).writerows(rows[hi:])
t, n = defaultdict(float), defaultdict(int)
for r in rows[hi + 1:]:
    d = dict(zip(h, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = d["Kernel Name"].split("(")[0].split("::")[-1]
    v = float(d["Metric Value"].replace(",", ""))
    t[name] += v / (1e3 if d["Metric Unit"] == "ns" else 1 if d["Metric Unit"] == "us" else 1e-3)
    n[name] += 1
tot = sum(t.values())
with open(os.path.join(out, f"{tag}_{w}_launch_shares.txt"), "w") as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --workload {w} --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config\n")
    f.write("kernel, launches, total_us, share_of_all_kernel_time (cold-cache serialised: compare shares only; torch kernels = synthetic-volume generator)\n")
    for k in sorted(t, key=lambda k: -t[k]):
        f.write(f"{k}, {n[k]}, {t[k]:.1f}, {100 * t[k] / tot:.1f}%\n")

# full captures -> metric tables
traffic = {}
for short in ("enc", "dec", "swap"):
    rep = os.path.join(root, "gpurun_out", f"full_{short}_{w}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines()))
    hh, units, vals = rr[0], rr[1], rr[2]
    d = dict(zip(hh, zip(units, vals)))
    with open(os.path.join(out, f"{tag}_{w}_full_{short}_{w}_metrics.csv"), "w") as f:
        wr = csv.writer(f)
        wr.writerow(["metric", "unit", "value"])
        wr.writerow(["Kernel Name", "", d["Kernel Name"][1]])
        for k in hh:
            if k in KEEP or "issue_stalled" in k and k.endswith("_per_warp_active.pct"):
                wr.writerow([k, d[k][0], d[k][1]])
    def to_bytes(k):
        u, v = d[k]
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    traffic[short] = int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
if "enc" in traffic:
    p = os.path.join(out, "traffic.json")
    j = json.load(open(p)) if os.path.exists(p) else {}
    j[w] = {"lz4_encode_kernel_dram_bytes_per_launch": traffic["enc"], "lz4_decode_kernel_dram_bytes_per_launch": traffic.get("dec"),
            "source": f"profiles/{tag}_{w}_full_enc_{w}_metrics.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
    json.dump(j, open(p, "w"), indent=1)
print("ok", traffic)
