#!/usr/bin/env python
"""A/B timing of two builds of librpbmd.so in ONE process on the same device, interleaved:
    python tools/ab_bench.py libA.so libB.so [workload] [steps] [rounds]
Same timing as bench.py's headline loop (CUDA events per step on the library stream, L2 flushed between steps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from reactive_pb_nn_md_b200 import engine
from reactive_pb_nn_md_b200._binding import Library

# an argument "NAME=VALUE:path" sets that environment variable while the context of that build is created
specs = []
for a in sys.argv[1:3]:
    env, path = (a.split(":", 1) if "=" in a.split(":", 1)[0] else ("", a))
    specs.append((env, path))
libs = [Library(os.path.abspath(p)) for _, p in specs]
wl = sys.argv[3] if len(sys.argv) > 3 else "c3"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 3
evb = bench.WORKLOADS[wl]["ms_evb"]
s = bench.build_system(wl)
sims = []
for (env, _), l in zip(specs, libs):
    if env:
        k, v = env.split("=", 1); os.environ[k] = v
    sims.append(engine.Simulation(s, bench.params_for(wl), library=l))
    if env:
        del os.environ[k]
for sim in sims:
    (sim.ms_evb_calculate_total_force_energy if evb else sim.calculate_total_force_energy)()
    sim.md_integrate_atomic(20, ms_evb=evb)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
res = [[], []]
for r in range(rounds):
    for k, sim in enumerate(sims):
        stream = torch.cuda.ExternalStream(sim.dll.rpb_get_stream(sim.ctx))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            flush.fill_(float(i)); torch.cuda.synchronize()
            evs[i][0].record(stream); sim.md_integrate_atomic(1, ms_evb=evb); evs[i][1].record(stream)
        torch.cuda.synchronize()
        t = np.array([a.elapsed_time(b) for a, b in evs])
        res[k].append((float(np.median(t)), float(t.mean())))
        if os.environ.get("AB_SHOW_OUTLIERS"):
            big = np.argsort(-t)[:8]
            print("AB"[k], "round", r, "slowest steps:", ["%d:%.3f" % (i, t[i]) for i in sorted(big)], " p10 %.4f p90 %.4f" % (np.percentile(t, 10), np.percentile(t, 90)))
for k in range(2):
    print("AB"[k], sys.argv[1 + k], " median ms/step per round:", ["%.4f" % m for m, _ in res[k]], " mean:", ["%.4f" % a for _, a in res[k]],
          " S=%s" % (sims[k].evb()["n_states"] if evb else 1))
