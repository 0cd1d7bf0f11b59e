"""Multi-GPU partitioning of independent subproblems.

The reference scales out only by running one JVM per `subset-N.clusters` file
(doc/MultiBreakSV-Manual.txt:411-473; generateIndependentSubproblems.pl).  Subproblems share nothing, so they
are spread over the GPUs of one box by LPT greedy on weight = chains x fragment x alignment count, with no
data-path collective.  A single giant component cannot be split spatially (every fragment is coupled through the
per-experiment totals), so its chains are split across GPUs instead and the integer tallies are summed with one
reduce (torch.distributed, NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._abi import I32P, I64P, check, load_library


def lpt_shard(weights: np.ndarray, n_parts: int) -> np.ndarray:
    """part[i] of every subproblem: descending weight, each to the least loaded part (mbsv_lpt_shard)."""
    w = np.ascontiguousarray(weights, np.int64)
    out = np.zeros(len(w), np.int32)
    check(load_library().mbsv_lpt_shard(w.ctypes.data_as(I64P), len(w), n_parts, out.ctypes.data_as(I32P)))
    return out


def split_chains(n_chains: int, n_parts: int, rank: int) -> tuple[int, int]:
    """(first chain, number of chains) of `rank` when one subproblem's chains are split over n_parts GPUs.
    Chain c keeps the seed 123456 + c wherever it runs, so chain 0 stays the reference's chain."""
    first, count = C.c_int32(0), C.c_int32(0)
    check(load_library().mbsv_split_chains(n_chains, n_parts, rank, C.byref(first), C.byref(count)))
    return first.value, count.value


def reduce_tallies(out, dst: int = 0, device=None):
    """Sum aln_count / frag_err_count / cluster_hist over ranks onto `dst` (one reduce of the concatenated
    integer tallies).  `out` is a RunOutput; returns the reduced arrays on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    flat = np.concatenate([out.aln_count.astype(np.int64), out.frag_err_count.astype(np.int64),
                           out.cluster_hist.astype(np.int64)])
    t = torch.from_numpy(flat)
    if device is not None:
        t = t.to(device)
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    if dist.get_rank() != dst:
        return None
    r = t.cpu().numpy()
    a, f = len(out.aln_count), len(out.frag_err_count)
    return r[:a], r[a:a + f], r[a + f:]
