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


def _shared_worker(rank, tag, q):
    sys.path.insert(0, ROOT)
    import time
    from occ_b200 import parallel
    n = 1000
    if rank == 0:
        d = parallel.SharedHostArray(tag + "_D", n, True)
        c = parallel.SharedHostArray(tag + "_c", 4, True, dtype=np.int64, register=False)
        c.array[:] = 0
        d.array[:] = np.arange(n)
        c.array[0] = 1                              # "D is ready"
        t0 = time.time()
        while c.array[3] < 1 and time.time() - t0 < 60:
            pass
        q.put(("r0", float(d.array[-1]), int(c.array[3])))
        d.close()
        c.close()
    else:
        t0 = time.time()
        d = c = None
        while time.time() - t0 < 60:                # rank 0 may not have created the segments yet
            try:
                c = parallel.SharedHostArray(tag + "_c", 4, False, dtype=np.int64, register=False)
                d = parallel.SharedHostArray(tag + "_D", n, False)
                break
            except FileNotFoundError:
                time.sleep(0.01)
        while c.array[0] < 1 and time.time() - t0 < 60:
            pass
        s = float(d.array.sum())
        d.array[-1] = -7.0                          # this rank's "slice of K"
        c.array[3] = 1
        q.put(("r1", s, 0))
        d.close()
        c.close()


def test_shared_host_array_rendezvous_between_two_processes():
    """the host side of DistributedDFT.build_shared: node-shared buffers + control words (no collective on the host path)"""
    from occ_b200 import parallel
    assert parallel.slice_len(1, 10, 8) * 8 >= parallel.pack_len(1, 10)
    assert parallel.slice_len(2, 3348, 8) == -(-(2 * 3348 * 3348 + 4) // 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    tag = f"occb200_test_{os.getpid()}"
    procs = [ctx.Process(target=_shared_worker, args=(r, tag, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((k, (a, b)) for k, a, b in (q.get(timeout=120), q.get(timeout=120)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got["r1"][0] == 999 * 1000 / 2            # rank 1 saw rank 0's D
    assert got["r0"] == (-7.0, 1)                    # rank 0 saw rank 1's slice and its completion word


def _cosx_worker(rank, world, port, q):
    """COSX N>1 path on CPU: each gloo rank evaluates K' on the batches occ_b200.cosx.batch_shard gives it (the rule
    the library applies for options.rank / world_size; stand-in builder: the CPU oracle, every batch passed as its own
    atom grid), all-reduces the partial K' and must reproduce the unsharded sweep."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from occ_b200 import aobasis, cosx, molecules, parallel
    from oracle import oracle as O
    Z, pos = molecules.water()
    b = aobasis.AOBasis.load(Z, pos, "def2-svp", pure=True)
    ob = O.Basis(b.as_arrays())
    gs = O.default_grid_settings(max_angular_points=26, min_angular_points=26, radial_precision=1e-4,
                                 reduced_first_row_element_grid=0)
    pts, w, info = O.molecular_grid(Z, pos, ob, gs, nthreads=1)
    off = O.cosx_atom_grids(pts, info)
    D2q = molecules.synthetic_density(b, 1)
    mine = cosx.batch_shard(off, rank, world)
    spts = np.concatenate([pts[s:s + c] for s, c in mine])
    sw = np.concatenate([w[s:s + c] for s, c in mine])
    soff = np.concatenate([[0], np.cumsum([c for _, c in mine])]).astype(np.uint64)
    part, st = O.cosx_K_for_density(ob, soff, spts, sw, D2q)
    buf = torch.from_numpy(part.ravel(order="F").copy())
    parallel.allreduce_partial(buf)
    if rank == 0:
        full, stf = O.cosx_K_for_density(ob, off, pts, w, D2q)
        nb_all = sum(len(cosx.batch_shard(off, r, world)) for r in range(world))
        q.put((float(np.abs(buf.numpy().reshape(b.nbf, b.nbf).T - full).max()), len(mine), nb_all, stf["nbatches"]))
    dist.barrier()
    dist.destroy_process_group()


def test_cosx_two_rank_gloo_sum_equals_unsharded_sweep():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_cosx_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    dk, n0, nall, nfull = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert n0 > 0 and nall == nfull
    assert dk < 1e-12
