#!/usr/bin/env python
"""Per-kernel CUDA-event times of one workload on a serial-stream context:  python tools/time_kernels.py c3 [n_steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["RPB_SERIAL_STREAMS"] = "1"
import torch
import bench
from reactive_pb_nn_md_b200 import engine
from reactive_pb_nn_md_b200._binding import load_cuda
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
s = bench.build_system(wl)
evb = bench.WORKLOADS[wl]["ms_evb"]
sim = engine.Simulation(s, bench.params_for(wl), library=load_cuda())
(sim.ms_evb_calculate_total_force_energy if evb else sim.calculate_total_force_energy)()
sim.md_integrate_atomic(5, ms_evb=evb)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
sim.timers_enable(True); sim.timers(reset=True)
for k in range(n):
    if not os.environ.get("NO_FLUSH"):
        flush.fill_(float(k)); torch.cuda.synchronize()
    sim.md_integrate_atomic(1, ms_evb=evb)
tm = sim.timers()
tot = tm["step_total"][0] / n
print("%s variant=%s step %.1f us  S=%s" % (wl, os.environ.get("RPB_PAIR_VARIANT", "-"), 1e3 * tot, sim.evb()["n_states"] if evb else 1))
for k, (ms, calls) in sorted(tm.items(), key=lambda kv: -kv[1][0]):
    if calls and k != "step_total":
        print("  %-20s %7.1f us/step  %6.1f us/launch  x%.2f" % (k, 1e3 * ms / n, 1e3 * ms / calls, calls / n))
