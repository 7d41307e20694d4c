"""N>1 path on CPU: two gloo ranks each build the PARTIAL Vxc of their leaf shard (stand-in builder: the CPU oracle on
the leaves the library's own partition assigns to the rank), pack [K | exc, n_alpha, n_beta], all-reduce through
occ_b200.parallel.allreduce_partial, and must reproduce the unsharded build."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from occ_b200 import aobasis, dft, molecules, parallel
    from oracle import oracle as O
    Z, pos = molecules.water()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp")
    ob = O.Basis(b.as_arrays())
    pts, w, _ = O.molecular_grid(Z, pos, ob, nthreads=1)
    spts, sw, perm, leaves = O.hierarchy(pts, w)
    D2 = molecules.synthetic_density(b, 2)
    leaf_rank, _ = dft.partition_leaves(b, leaves, world)
    mine = leaves[leaf_rank == rank]
    part = O.xc_build(ob, spts, sw, mine, O.METHODS["pbe"], 2, D2, nthreads=1)
    nbf, nspin = b.nbf, 2
    buf = torch.zeros(parallel.pack_len(nspin, nbf), dtype=torch.float64)
    buf[:nspin * nbf * nbf] = torch.from_numpy(part["K"].ravel(order="F").copy())
    buf[nspin * nbf * nbf:nspin * nbf * nbf + 3] = torch.tensor([part["exc"], part["n_alpha"], part["n_beta"]], dtype=torch.float64)
    parallel.allreduce_partial(buf)
    K, sc = parallel.unpack(buf, nspin, nbf)
    if rank == 0:
        full = O.xc_build(ob, spts, sw, leaves, O.METHODS["pbe"], 2, D2, nthreads=1)
        q.put((float(np.abs(K.numpy() - full["K"]).max()), abs(float(sc[0]) - full["exc"]),
               abs(float(sc[1]) - full["n_alpha"]), abs(float(sc[2]) - full["n_beta"]), int((leaf_rank == 0).sum()),
               int((leaf_rank == 1).sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sum_equals_unsharded_build():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    dk, de, dna, dnb, n0, n1 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert n0 > 0 and n1 > 0
    assert dk < 1e-12 and de < 1e-12 and dna < 1e-12 and dnb < 1e-12
