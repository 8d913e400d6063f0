"""Synthetic workloads for configs 3-5 of BASELINE.json.

`numpy.random.Generator(PCG64(20160826 + config))`; per-cell 6-coefficient error models drawn uniformly from the
ranges of the reference's bundled ``o.ifm`` fit; counts from the model's own generative story (drop-out -> Poisson(0.1),
otherwise negative binomial around exp(corr.a * m + corr.b)); a flat prior on a 401-point grid.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import pandas as pd

SEED_BASE = 20160826
OIFM_RANGES = {  # min / max over data/o.ifm.rda
    "conc.b": (-6.22, -0.87), "conc.a": (0.56, 1.07), "corr.b": (0.23, 2.37), "corr.a": (0.29, 0.70),
    "corr.theta": (0.69, 1.12),
}
FAIL_R = -2.302585  # log(0.1)

CONFIGS = {
    3: dict(n_genes=20000, n_cells=2000, batch=False),
    4: dict(n_genes=30000, n_cells=10000, batch=False),
    5: dict(n_genes=30000, n_cells=10000, batch=True),
}


@dataclass
class Workload:
    counts: np.ndarray          # genes x cells int32, Fortran order
    models: pd.DataFrame        # cells x 6
    prior: pd.DataFrame         # x, y
    groups: pd.Categorical      # two levels, first half / second half
    batch: Optional[pd.Categorical]
    name: str


def make_prior(n_genes: int, length_out: int = 400, max_value: float = 4.8) -> pd.DataFrame:
    x = np.linspace(0.0, max_value, length_out + 1)
    y = np.full(length_out + 1, 1.0) + 1.0 / n_genes
    return pd.DataFrame({"x": x, "y": y / y.sum()})


def make_models(rng: np.random.Generator, n_cells: int) -> pd.DataFrame:
    cols = {}
    for name in ("conc.b", "conc.a"):
        lo, hi = OIFM_RANGES[name]
        cols[name] = rng.uniform(lo, hi, n_cells)
    cols["fail.r"] = np.full(n_cells, FAIL_R)
    for name in ("corr.b", "corr.a", "corr.theta"):
        lo, hi = OIFM_RANGES[name]
        cols[name] = rng.uniform(lo, hi, n_cells)
    return pd.DataFrame(cols, index=[f"cell{i}" for i in range(n_cells)])


def _gene_magnitudes(rng: np.random.Generator, n_genes: int):
    """True magnitudes (natural-log FPM) per group: log10(e^m + 1) ~ U(0, 4); 10 % of genes shifted in group 2."""
    u = rng.uniform(0.0, 4.0, n_genes)
    with np.errstate(divide="ignore"):
        m1 = np.log(np.power(10.0, u) - 1.0)
    shifted = rng.uniform(size=n_genes) < 0.10
    shift = rng.uniform(1.0, 3.0, n_genes) * np.log(2.0) * np.where(rng.uniform(size=n_genes) < 0.5, -1.0, 1.0)
    return m1, np.where(shifted, m1 + shift, m1)


def make_counts(rng: np.random.Generator, models: pd.DataFrame, n_genes: int, cell_chunk: int = 256) -> np.ndarray:
    n_cells = len(models)
    half = n_cells // 2
    m1, m2 = _gene_magnitudes(rng, n_genes)
    counts = np.empty((n_genes, n_cells), dtype=np.int32, order="F")
    ca, cb = models["conc.a"].to_numpy(), models["conc.b"].to_numpy()
    ra, rb, th = models["corr.a"].to_numpy(), models["corr.b"].to_numpy(), models["corr.theta"].to_numpy()
    for c0 in range(0, n_cells, cell_chunk):
        c1 = min(n_cells, c0 + cell_chunk)
        m = np.where((np.arange(c0, c1) < half)[None, :], m1[:, None], m2[:, None])  # genes x chunk
        with np.errstate(over="ignore", invalid="ignore"):
            pfail = 1.0 / (np.exp(ca[None, c0:c1] * m + cb[None, c0:c1]) + 1.0)
            mu = np.exp(ra[None, c0:c1] * m + rb[None, c0:c1])
        pfail = np.where(np.isnan(pfail), 1.0, pfail)
        theta = np.broadcast_to(th[None, c0:c1], m.shape)
        p = theta / (theta + mu)
        nb = rng.negative_binomial(theta, np.clip(p, 1e-300, 1.0))
        po = rng.poisson(0.1, size=m.shape)
        fail = rng.uniform(size=m.shape) < pfail
        counts[:, c0:c1] = np.where(fail, po, np.minimum(nb, 2**31 - 1)).astype(np.int32)
    return counts


def make_counts_torch(models: pd.DataFrame, n_genes: int, seed: int, device, pinned: bool = True,
                      cell_chunk: int = 500):
    """Same generative story sampled on the GPU with torch (negative binomial as a Gamma-Poisson mixture) for the
    bench-sized workloads, where the numpy generator takes minutes.  Returns a (n_genes x n_cells) int32
    Fortran-ordered numpy array backed by pinned host memory.  Not bit-identical to make_counts."""
    import torch

    n_cells = len(models)
    half = n_cells // 2
    rng = np.random.Generator(np.random.PCG64(seed))
    m1, m2 = _gene_magnitudes(rng, n_genes)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    host = torch.empty((n_cells, n_genes), dtype=torch.int32, pin_memory=pinned)  # [cell][gene] == Fortran genes x cells
    tm1 = torch.as_tensor(m1, device=device)
    tm2 = torch.as_tensor(m2, device=device)
    col = {k: torch.as_tensor(models[k].to_numpy().copy(), device=device) for k in ("conc.a", "conc.b", "corr.a", "corr.b", "corr.theta")}
    for c0 in range(0, n_cells, cell_chunk):
        c1 = min(n_cells, c0 + cell_chunk)
        first = (torch.arange(c0, c1, device=device) < half)[:, None]
        m = torch.where(first, tm1[None, :], tm2[None, :])  # chunk x genes
        pfail = 1.0 / (torch.exp(col["conc.a"][c0:c1, None] * m + col["conc.b"][c0:c1, None]) + 1.0)
        pfail = torch.nan_to_num(pfail, nan=1.0)
        mu = torch.exp(col["corr.a"][c0:c1, None] * m + col["corr.b"][c0:c1, None])
        theta = col["corr.theta"][c0:c1, None].expand_as(m).contiguous()
        lam = torch._standard_gamma(theta, generator=g) * (mu / theta)
        nb = torch.poisson(lam.clamp(max=2.0e9), generator=g)
        po = torch.poisson(torch.full_like(m, 0.1), generator=g)
        fail = torch.rand(m.shape, device=device, dtype=torch.float64, generator=g) < pfail
        host[c0:c1].copy_(torch.where(fail, po, nb).clamp(max=2**31 - 1).to(torch.int32))
    torch.cuda.synchronize(device) if str(device).startswith("cuda") else None
    return host.numpy().T  # (n_genes, n_cells), Fortran-contiguous view of the pinned buffer


def make_workload(config: int = 3, n_genes: Optional[int] = None, n_cells: Optional[int] = None,
                  batch: Optional[bool] = None, seed: Optional[int] = None) -> Workload:
    """Config 3 / 4 / 5 of BASELINE.json, optionally down-sized (same generator, same seed)."""
    spec = dict(CONFIGS[config])
    if n_genes is not None:
        spec["n_genes"] = n_genes
    if n_cells is not None:
        spec["n_cells"] = n_cells
    if batch is not None:
        spec["batch"] = batch
    rng = np.random.Generator(np.random.PCG64(SEED_BASE + config if seed is None else seed))
    models = make_models(rng, spec["n_cells"])
    counts = make_counts(rng, models, spec["n_genes"])
    half = spec["n_cells"] // 2
    groups = pd.Categorical(np.where(np.arange(spec["n_cells"]) < half, "g1", "g2"), categories=["g1", "g2"])
    b = None
    if spec["batch"]:
        b = pd.Categorical(np.where(rng.uniform(size=spec["n_cells"]) < 0.5, "batch1", "batch2"),
                           categories=["batch1", "batch2"])
    return Workload(counts=counts, models=models, prior=make_prior(spec["n_genes"]), groups=groups, batch=b,
                    name=f"cfg{config}:{spec['n_genes']}x{spec['n_cells']}")
