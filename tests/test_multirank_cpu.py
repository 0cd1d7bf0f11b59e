"""Host-side logic of the N>1 paths on the CPU (gloo, world_size 2): LPT sharding of independent subproblems and
the chain-split + reduce of a single component.  The compute stand-in is the oracle (no GPU here); what is under
test is multibreak_sv_b200.sharding and the skeleton sharding used by bench.py."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import EXAMPLE, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_lpt_shard_properties(mbsv):
    from multibreak_sv_b200 import sharding, synth
    sk = synth.draw_skeleton("cfg4", n_frag=200000)
    w = sk.weights * 64
    for parts in (1, 2, 4, 8):
        part = sharding.lpt_shard(w, parts)
        assert part.min() >= 0 and part.max() < parts
        loads = np.bincount(part, weights=w, minlength=parts)
        # LPT guarantee: max load <= mean + max item
        assert loads.max() <= loads.mean() + w.max()
        assert loads.max() / loads.mean() < 1.05 or w.max() > 0.05 * loads.mean()
        assert np.array_equal(part, sharding.lpt_shard(w, parts))          # deterministic
    # the heaviest items land on distinct parts
    part = sharding.lpt_shard(np.array([10, 9, 8, 1, 1, 1], np.int64), 3)
    assert len(set(part[:3].tolist())) == 3
    # shards of a skeleton partition its fragments
    part = sharding.lpt_shard(w, 2)
    n0 = sk.select(np.nonzero(part == 0)[0]).sizes.sum()
    n1 = sk.select(np.nonzero(part == 1)[0]).sizes.sum()
    assert n0 + n1 == sk.sizes.sum()
    assert sharding.split_chains(64, 8, 3) == (24, 8)
    assert [sharding.split_chains(10, 4, r) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from multibreak_sv_b200 import sharding
    prob = orc.load_files(os.path.join(EXAMPLE, "gasv.clusters"), os.path.join(EXAMPLE, "assignments.txt"),
                          os.path.join(EXAMPLE, "experiments.txt"))
    chains, iters = 6, 3000
    first, n = sharding.split_chains(chains, world, rank)
    out = orc.run(prob, "incremental", iters, n_chains=n, seed=123456 + first)
    red = sharding.reduce_tallies(out, dst=0)
    if rank == 0:
        full = orc.run(prob, "incremental", iters, n_chains=chains)
        ok = (np.array_equal(red[0], full.aln_count.astype(np.int64)) and
              np.array_equal(red[1], full.frag_err_count.astype(np.int64)) and
              np.array_equal(red[2], full.cluster_hist.astype(np.int64)))
        q.put(bool(ok))
    dist.destroy_process_group()


def test_chain_split_and_reduce_gloo_world2(oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
