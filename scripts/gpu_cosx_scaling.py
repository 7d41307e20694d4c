#!/usr/bin/env python
"""COSX on N GPUs of one box (torchrun, one rank per GPU): grid batches dealt round-robin (options.rank / world_size),
D2q replicated, the partial K' matrices summed by ONE NCCL all-reduce.  Prints one JSON line on rank 0.
usage: torchrun --nproc-per-node N scripts/gpu_cosx_scaling.py [mol basis pure angular reps check]"""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, cosx

a = sys.argv[1:]
mol, bas, pure, ang, reps, check = (a + ["valinomycin", "def2-tzvp", "0", "50", "3", "1"][len(a):])[:6]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Z, pos = molecules.MOLECULES[mol]()
b = aobasis.AOBasis.load(Z, pos, bas, pure=pure == "1")
nbf = b.nbf
D2q = torch.from_numpy(np.ascontiguousarray(molecules.synthetic_density(b, 1).T)).cuda()   # identity projector: D2q = 2 D
sgx = cosx.SemiNumericalExchange(b, cosx.sgx_grid_settings(int(ang)), projector=None, device=local, rank=rank, world_size=world)
K = torch.zeros((nbf, nbf), dtype=torch.float64, device="cuda")


def step():
    sgx.compute_K_for_density_device(D2q, K)
    if world > 1:
        dist.all_reduce(K)
    return K


step()
ts = []
for _ in range(int(reps)):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    ts.append(float(dt.item()))
out = {"row": "f-2 COSX K' for one density, N GPUs", "workload": f"{mol}/{bas} sgx grid {ang} angular", "n_gpus": world,
       "nbf": nbf, "grid_points": sgx.num_grid_points(), "batches_this_rank": sgx.num_batches(),
       "ms_per_call_max_over_ranks": 1e3 * min(ts), "grid_points_per_s": sgx.num_grid_points() / min(ts)}
if check == "1" and world > 1 and rank == 0:
    full = cosx.SemiNumericalExchange(b, atom_grids=(sgx.atom_offsets, sgx.points, sgx.weights), projector=None, device=local)
    Kf = torch.zeros_like(K)
    full.compute_K_for_density_device(D2q, Kf)
    out["max_abs_diff_vs_one_gpu"] = float((Kf - K).abs().max().item())
    out["K_absmax"] = float(Kf.abs().max().item())
    full.close()
if rank == 0:
    print(json.dumps(out), flush=True)
sgx.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
