#!/usr/bin/env python
"""One process, N GPUs: scde_b200_expression_difference on a multi-device context (scde_b200_create_multi) -- the form an R
session would use (n.cores -> devices).  Config 4 by default; prints one JSON line per device count.

    python tools/bench_multi_inprocess.py --devices 1 2 4 8 --steps 5
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, nargs="+", default=[1, 2])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--config", type=int, default=4)
    args = ap.parse_args()
    import torch

    import bench
    from scde_b200 import _lib, api

    G, C = bench.config_size(args.config)
    models, counts, prior, group, batch = bench.workload_host(args.config, G, C, torch.device("cuda", 0))
    mm, lt, sq = api.pack_models(models)
    x, y = prior["x"].to_numpy(), prior["y"].to_numpy()
    zi = api._zero_index(api.fold_change_grid(x), 0.0)
    ref = None
    for n in args.devices:
        if n > _lib.lib().scde_b200_device_count():
            continue
        ctx = _lib.Context(devices=list(range(n)))

        def call():
            return api.expression_difference_call(ctx, counts, mm, x, y, group, bench.N_BOOT, 1, batch_codes=batch,
                                                  n_batch_levels=2 if batch is not None else 0, zero_index=zi,
                                                  local_theta=lt, sqlogit=sq)

        for _ in range(2):
            res = call()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = call()
        ms = (time.perf_counter() - t0) / args.steps * 1e3
        if ref is None:
            ref = res
        same = bool(np.array_equal(res["idx"], ref["idx"]) and np.array_equal(res["z"], ref["z"]))
        print(json.dumps({"in_process_devices": n, "ms_per_call": ms, "genes_per_s": G / (ms * 1e-3),
                          "equals_first_device_count": same, "stage_ms_slowest_shard": res["stats"]["ms"],
                          "workload": f"cfg{args.config}: {G} x {C}, host (pinned) buffers, wall clock of the call"}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
