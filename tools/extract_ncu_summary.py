"""ncu --set full report (.ncu-rep) -> per-kernel summary CSV (averaged over the captured launches) and the DRAM
traffic JSON that bench.py attaches to its roofline objects.
usage: extract_ncu_summary.py report.ncu-rep|raw_page.csv out_summary.csv out_traffic.json workload"""
import csv, json, subprocess, sys, collections, io
rep, out_csv, out_json, workload = sys.argv[1:5]
# a .csv argument is the raw page already exported on the GPU box (ncu -i report --page raw --csv)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
cols = {"duration_us": "gpu__time_duration.sum", "dram_read_B": "dram__bytes_read.sum", "dram_write_B": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "registers": "launch__registers_per_thread",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
        "grid": "launch__grid_size", "block": "launch__block_size"}
units = rows[1]
ki = hdr.index("Kernel Name")
agg = collections.OrderedDict()
def scale(v, u):
    v = float(v.replace(",", ""))
    u = u.lower()
    if u in ("msecond", "ms"): return v * 1e3
    if u in ("nsecond", "ns"): return v / 1e3
    if u in ("second", "s"): return v * 1e6
    if u == "kbyte": return v * 1e3
    if u == "mbyte": return v * 1e6
    if u == "gbyte": return v * 1e9
    return v
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, collections.defaultdict(list))
    for k, m in cols.items():
        if m in hdr:
            i = hdr.index(m)
            try: a[k].append(scale(r[i], units[i]))
            except Exception: pass
with open(out_csv, "w") as f:
    f.write("kernel,launches," + ",".join(cols) + "\n")
    for name, a in agg.items():
        n = len(a["duration_us"])
        f.write('"' + name + '",' + str(n) + "," + ",".join("%.4g" % (sum(a[k]) / len(a[k])) if a[k] else "" for k in cols) + "\n")
# bench.py kernel names -> CUDA kernels
mapping = {"pair_real_space": ["k_pair_verlet", "k_pair_tiles"], "pme_spread": ["k_spread"], "pme_fft": ["k_fft16_fwd_xy", "k_fft16_inv_xy", "k_fft_fwd_xy", "k_fft_inv_xy"],
           "pme_convolve": ["k_fft16_z_conv", "k_fft_z_conv", "k_conv_energy"], "evb_grid_broadcast": ["k_evb_broadcast_grid"],
           "evb_theta_mix": ["k_evb_theta_mix"], "evb_mix_forces": ["k_evb_mix_forces"], "evb_gather_mix": ["k_evb_gather_mix", "k_evb_gather_range"], "pme_gather": ["k_gather"]}
traffic = {}
for bname, kn in mapping.items():
    vals = []
    for name, a in agg.items():
        if any(name.startswith(k) for k in kn) and a["dram_read_B"]:
            vals += [x + y for x, y in zip(a["dram_read_B"], a["dram_write_B"])]
    if vals: traffic[bname] = sum(vals) / len(vals)
try: allj = json.load(open(out_json))
except Exception: allj = {}
allj[workload] = traffic
allj["_note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full (cold caches: ncu flushes L2 before every replay)"
json.dump(allj, open(out_json, "w"), indent=1)
print(open(out_csv).read())
