#!/usr/bin/env python
"""Diabatic-state sharding over NCCL on real GPUs vs the single-process oracle:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_sharded_nccl.py [n_steps]

Every rank drives the CUDA library through engine.Simulation (rank r of N); rank 0 also runs the CPU oracle on the same
box and compares neighbour-independent results (S, proton log, hydronium molecule, energies 1e-10, forces 1e-8 RMS,
positions after n_steps)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from reactive_pb_nn_md_b200 import engine
    from reactive_pb_nn_md_b200._binding import Library, load_cuda
    from tests.util import E_RTOL, F_RTOL, rel_rms, small_params, water_system
    n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s = water_system(10, hydronium=True)
    p = small_params()
    sim = engine.Simulation(s, p, library=load_cuda(), device=local, rank=rank, world_size=world, process_group=dist.group.WORLD)
    sim.ms_evb_calculate_total_force_energy()
    f0, e0, ev0 = sim.forces(), sim.energies(), sim.evb()
    sim.md_integrate_atomic(n_steps, ms_evb=True)
    st = sim.download_state()
    ok = True
    if rank == 0:
        ref = engine.Simulation(s, p, library=Library(os.path.join(ROOT, "oracle", "librpbmd_oracle.so")))
        ref.ms_evb_calculate_total_force_energy()
        er, evr = ref.energies(), ref.evb()
        checks = {
            "n_states": ev0["n_states"] == evr["n_states"],
            "proton_log": np.array_equal(ev0["proton_log"], evr["proton_log"]),
            "adiabatic": abs(ev0["adiabatic_potential"] - evr["adiabatic_potential"]) <= E_RTOL * abs(evr["adiabatic_potential"]),
            "hamiltonian": np.abs(ev0["hamiltonian"] - evr["hamiltonian"]).max() <= E_RTOL * np.abs(np.diag(evr["hamiltonian"])).max(),
            "potential_energy": abs(e0["potential_energy"] - er["potential_energy"]) <= E_RTOL * abs(er["potential_energy"]),
            "forces": rel_rms(f0, ref.forces()) < F_RTOL,
        }
        ref.md_integrate_atomic(n_steps, ms_evb=True)
        xr = ref.download_state()
        checks["hydronium_mol"] = st["hydronium_mol"] == xr["hydronium_mol"]
        checks["xyz_after_%d_steps" % n_steps] = float(np.abs(st["xyz"] - xr["xyz"]).max()) < 1e-9
        checks["force_after_steps"] = rel_rms(st["force"], xr["force"]) < F_RTOL
        ok = all(checks.values())
        print("sharded NCCL x%d vs oracle: %s  %s" % (world, "OK" if ok else "FAIL", checks), flush=True)
    # all ranks must hold identical replicated state
    t = torch.from_numpy(st["xyz"]).cuda()
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool((lo == hi).all().item())
    if rank == 0:
        print("replicated state bit-identical across ranks:", same, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if (ok and same) else 1)


if __name__ == "__main__":
    main()
