"""world_size-2 gloo test of the sharding host logic (no GPU): shard ranges, the single end-of-path all-gather and the
global BH step, with the per-shard compute supplied by the CPU oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_and_balance():
    from scde_b200.distributed import shard_range

    for n, w in [(30000, 8), (13788, 3), (5, 8), (1, 1), (0, 2)]:
        rs = [shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        sizes = [b - a for a, b in rs]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from scde_b200 import _lib, api, synth
    from scde_b200.distributed import expression_difference_sharded

    w = synth.make_workload(3, n_genes=37, n_cells=10, seed=3)
    codes = np.asarray(w.groups.codes)
    gi = (np.nonzero(codes == 0)[0], np.nonzero(codes == 1)[0])
    diffv = api.fold_change_grid(w.prior["x"].to_numpy())

    def run_shard(g0, g1):  # CPU oracle stands in for the device shard
        r = O.expression_difference(w.models, w.counts[g0:g1], w.prior["x"].to_numpy(), w.prior["y"].to_numpy(), gi,
                                    nboot=20, seed=1)
        return {"idx": r["idx"], "z": r["results"][:, 4]}

    def finish(full):
        cz = np.empty_like(full["z"])
        _lib.check(_lib.lib().scde_b200_bh_cz(_lib.p_f64(np.ascontiguousarray(full["z"])), len(full["z"]), _lib.p_f64(cz)))
        return full["idx"], full["z"], cz

    idx, z, cz = expression_difference_sharded(run_shard, 37, finish)
    np.savez(os.path.join(outdir, f"r{rank}.npz"), idx=idx, z=z, cz=cz)
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from scde_b200 import synth

    w = synth.make_workload(3, n_genes=37, n_cells=10, seed=3)
    codes = np.asarray(w.groups.codes)
    ref = O.expression_difference(w.models, w.counts, w.prior["x"].to_numpy(), w.prior["y"].to_numpy(),
                                  (np.nonzero(codes == 0)[0], np.nonzero(codes == 1)[0]), nboot=20, seed=1)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        assert np.array_equal(d["idx"], ref["idx"])
        np.testing.assert_allclose(d["z"], ref["results"][:, 4], rtol=0, atol=0)
        np.testing.assert_allclose(d["cz"], ref["results"][:, 5], rtol=1e-12, atol=1e-14)  # BH over ALL genes
