"""Gene sharding over one process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Genes are independent through the whole device path; only the Benjamini-Hochberg step of cZ needs every gene's Z
(R/functions.R:5051).  Each rank runs a contiguous gene range -- the same split the reference's `chunk()` makes for
`n.cores` (R/functions.R:606) -- with the n.cores = 1 draw semantics (one draw set for all genes, so the result does not
depend on the number of ranks), and the per-shard grid indices and Z are all-gathered once at the end.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np


def shard_range(n_genes: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced gene range of `rank` (first n_genes % world ranks get one more gene)."""
    base, rem = divmod(n_genes, world)
    g0 = rank * base + min(rank, rem)
    return g0, g0 + base + (1 if rank < rem else 0)


def gather_rows(local: np.ndarray, n_genes: int, dist_mod=None, device=None) -> np.ndarray:
    """All-gather per-gene rows (first axis = this rank's genes) into the full array, in gene order."""
    import torch
    import torch.distributed as dist

    dist_mod = dist_mod or dist
    if not (dist_mod.is_available() and dist_mod.is_initialized()) or dist_mod.get_world_size() == 1:
        return local
    world = dist_mod.get_world_size()
    sizes = [shard_range(n_genes, r, world) for r in range(world)]
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    longest = max(b - a for a, b in sizes)
    t = torch.zeros((longest, width), dtype=torch.from_numpy(np.zeros(1, local.dtype)).dtype, device=device or "cpu")
    t[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local).reshape(local.shape[0], width)).to(t.device)
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist_mod.all_gather(bufs, t)
    parts = [bufs[r][: b - a].cpu().numpy() for r, (a, b) in enumerate(sizes)]
    out = np.concatenate(parts, axis=0)
    return out.reshape((n_genes,) + local.shape[1:])


def expression_difference_sharded(run_shard: Callable[[int, int], dict], n_genes: int, finish: Callable[[dict], object],
                                  device=None, keys: Sequence[str] = ("idx", "z")):
    """run_shard(g0, g1) -> dict of per-gene arrays for genes [g0, g1); gathers `keys` over the ranks and calls
    finish(full_dict) (BH correction, labelling) on every rank."""
    import torch.distributed as dist

    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    g0, g1 = shard_range(n_genes, rank, world)
    local = run_shard(g0, g1)
    full = {k: gather_rows(np.asarray(local[k]), n_genes, device=device) for k in keys if k in local}
    return finish(full)


def scde_expression_difference_sharded(models, counts, prior, groups, n_randomizations: int = 100, expectation=0,
                                       seed: int = 1, context=None, device=None):
    """scde.expression.difference with the genes sharded over the ranks of the default process group."""
    import pandas as pd

    from . import _lib, api

    cm, genes = api._counts_for_models(models, counts)
    gcodes, glev = api._named_factor(groups, models.index)
    if len(glev) != 2:
        api._stop("wrong number of levels in the grouping factor (" + " ".join(map(str, glev)) + "), but must be two.")
    mm, lt, sq = api.pack_models(models)
    x = np.asarray(prior["x"], dtype=np.float64)
    diffv = api.fold_change_grid(x)
    zi = api._zero_index(diffv, expectation)
    ctx = context or _lib.default_context()

    def run_shard(g0, g1):
        job = api.DifferenceJob(ctx, cm, mm, x, np.asarray(prior["y"], dtype=np.float64), gcodes, n_randomizations, seed,
                                zero_index=zi, local_theta=lt, sqlogit=sq, gene_range=(g0, g1))
        try:
            job.run()
            return job.download()
        finally:
            job.close()

    def finish(full):
        return api._summary_frame(full["idx"], full["z"], diffv, genes)

    return expression_difference_sharded(run_shard, cm.shape[0], finish, device=device)
