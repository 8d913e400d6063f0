"""One worker of the CPU arm: the reference's logBootPosterior / logBootBatchPosterior on one contiguous gene chunk, as
one forked R worker of scde.posteriors(n.cores > 1) would run it (R/functions.R:606-617) -- started as a separate process
(``python -m oracle.ref_worker in.npz out.npz``) because the reference draws from libc's global rand() state and because
a process that holds a CUDA context must not fork.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs).

impl "reference": oracle/_ref/libscde_ref.so (the reference's own C++ compiled against the header shim);
impl "port":      oracle/scde_oracle.c (used only where no prebuilt _ref exists).
Every joint is run twice, with n.randomizations = nboot and with 0: the second call builds the same per-cell table and
does (almost) nothing else, so the difference is the bootstrap loop alone -- the table of a real 1875-gene chunk is
amortised over 1875 genes, here over a handful, and must not be charged to the per-gene rate.
Seed = 1 for every chunk: the draws of n.cores = 1, which is what the GPU path reproduces.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(inp: str, outp: str) -> None:
    from oracle import oracle as O

    d = np.load(inp)
    impl = str(d["impl"])
    if impl == "reference":
        from oracle import ref as X
    else:
        X = O
    mm, counts, mag = d["models"], np.asfortranarray(d["counts"]), d["mag"]
    nboot, seed, lt, sq = int(d["nboot"]), int(d["seed"]), int(d["localtheta"]), int(d["sqlogit"])
    group = d["group"]
    out = {}
    t_full = t_table = 0.0

    def timed(f):
        t0 = time.perf_counter()
        r = f()
        return r, time.perf_counter() - t0

    for lev in (0, 1):
        ii = np.nonzero(group == lev)[0]
        sub = np.asfortranarray(counts[:, ii])
        flat, off, uci = O.unique_counts(sub)  # unique() / match() of the chunk, R/functions.R:609-610
        m = np.asfortranarray(mm[ii])
        r, t1 = timed(lambda: X.log_boot_posterior(m, flat, off, uci, mag, nboot, seed=seed, localtheta=lt, sqlogit=sq))
        _, t0 = timed(lambda: X.log_boot_posterior(m, flat, off, uci, mag, 0, seed=seed, localtheta=lt, sqlogit=sq))
        out[f"jp{lev}"] = r["jp"]
        t_full += t1
        t_table += t0
    if "batch" in d.files:
        batch = d["batch"]
        L = int(batch.max()) + 1
        pools = [np.nonzero(batch == l)[0].astype(np.int32) for l in range(L)]
        flat, off, uci = O.unique_counts(counts)
        for lev in (0, 1):
            ii = np.nonzero(group == lev)[0]
            comp = np.bincount(batch[ii], minlength=L).astype(np.int32)
            r, t1 = timed(lambda: X.log_boot_batch_posterior(mm, flat, off, uci, mag, pools, comp, nboot, seed=seed,
                                                             localtheta=lt, sqlogit=sq))
            _, t0 = timed(lambda: X.log_boot_batch_posterior(mm, flat, off, uci, mag, pools, comp, 0, seed=seed,
                                                             localtheta=lt, sqlogit=sq))
            out[f"bjp{lev}"] = r["jp"]
            t_full += t1
            t_table += t0
    out["t_full"], out["t_table"] = np.float64(t_full), np.float64(t_table)
    np.savez(outp, **out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
