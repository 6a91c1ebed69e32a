"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total/avg us, share."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", ""))
    if r[ui] == "ns": v /= 1e3
    elif r[ui] == "ms": v *= 1e3
    elif r[ui] == "s": v *= 1e6   # msecond / second spellings
    name = r[ki].split("(")[0]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
# bench.py's own instrumentation (fp64 peak micro-benchmark, L2 flush fills) is listed but kept out of the shares
aux = lambda k: k.startswith("k_fp64_peak") or "at::" in k
tot = sum(a[1] for k, a in agg.items() if not aux(k))
print("kernel,launches,total_us,avg_us,share_of_step_kernels")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.2f,%s" % (k, n, t, t / n, "bench-instrumentation" if aux(k) else "%.4f" % (t / tot)))
