"""Input preparation that sits immediately before the hot path:

    scde_failure_probability  <-  scde.failure.probability  R/functions.R:725-750
    scde_expression_prior     <-  scde.expression.prior     R/functions.R:225-254
    clean_counts              <-  clean.counts (the vignette's filter; min.lib.size / min.reads / min.detected)

These produce the prior grid the differential-expression path consumes.  ``scde_expression_prior`` and
``scde_failure_probability`` go through the C ABI (``scde_b200_expression_prior`` / ``scde_b200_failure_probability``,
csrc/prior.cu): the O(genes x cells) work -- magnitudes, drop-out weights, the max.quantile order statistics, the linear
binning of ``density.default`` -- runs on the device, the 2048-point kernel smoothing on the host.  No fallback: without
a GPU they raise.  (A plain numpy restatement of the same published algorithm -- linear binning onto 2n points, FFT
convolution with the Gaussian kernel, linear interpolation -- lives in tests/prior_host.py: test infrastructure for the
CPU-only fixtures, not part of this package.)
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def clean_counts(counts: pd.DataFrame, min_lib_size: float = 1.8e3, min_reads: int = 10, min_detected: int = 5):
    """Filter counts matrix: cells by library complexity, genes by total reads and number of detecting cells."""
    v = counts.to_numpy()
    keep_cells = (v > 0).sum(axis=0) > min_lib_size
    v = v[:, keep_cells]
    keep_genes = v.sum(axis=1) > min_reads
    v2 = v[keep_genes]
    keep_genes2 = (v2 > 0).sum(axis=1) > min_detected
    out = counts.iloc[:, np.nonzero(keep_cells)[0]]
    out = out.iloc[np.nonzero(keep_genes)[0]]
    return out.iloc[np.nonzero(keep_genes2)[0]]


def scde_failure_probability(models: pd.DataFrame, magnitudes=None, counts=None, context=None) -> np.ndarray:
    """Drop-out probability per gene (rows) and cell (columns) (R/functions.R:725-750), on the device."""
    from . import _lib, api

    if magnitudes is None and counts is None:
        raise ValueError("ERROR: either magnitudes or counts should be provided")
    mm, _lt, sq = api.pack_models(models, clamp=False)
    ctx = context or _lib.default_context()
    if magnitudes is None:
        cm = counts.loc[:, list(models.index)].to_numpy() if isinstance(counts, pd.DataFrame) else np.asarray(counts)
        cm = _lib.i32(cm)
        G, Cn = cm.shape
        out = np.empty((G, Cn), dtype=np.float64, order="F")
        _lib.check(_lib.lib().scde_b200_failure_probability(ctx.handle, _lib.p_f64(mm), Cn, _lib.p_i32(cm), None, G, sq,
                                                            _lib.p_f64(out)))
        return out
    m = np.asarray(magnitudes.loc[:, list(models.index)].to_numpy() if isinstance(magnitudes, pd.DataFrame) else magnitudes,
                   dtype=np.float64)
    if m.ndim == 1:  # a common vector of magnitudes for all cells (:742-747)
        m = np.repeat(m[:, None], len(models), axis=1)
    m = _lib.f64(m)
    G, Cn = m.shape
    if Cn != len(models):
        raise ValueError("ERROR: provided magnitude data does not cover all of the cells specified in the model matrix")
    out = np.empty((G, Cn), dtype=np.float64, order="F")
    _lib.check(_lib.lib().scde_b200_failure_probability(ctx.handle, _lib.p_f64(mm), Cn, None, _lib.p_f64(m), G, sq,
                                                        _lib.p_f64(out)))
    return out


def scde_expression_prior(models: pd.DataFrame, counts, length_out: int = 400, show_plot: bool = False,
                          pseudo_count: float = 1, bw: float = 0.1, max_quantile: float = 1, max_value=None, context=None):
    """Expression-magnitude grid (``x``, log10 scale) and prior (``y``) (R/functions.R:225-254), through the C ABI
    (scde_b200_expression_prior): the genes x cells work on the device."""
    from . import _lib, api

    cm = counts.loc[:, list(models.index)].to_numpy() if isinstance(counts, pd.DataFrame) else np.asarray(counts)
    cm = _lib.i32(cm)
    G, Cn = cm.shape
    mm, _lt, sq = api.pack_models(models, clamp=False)
    K = int(length_out) + 1
    x, y, lp, gw = (np.empty(K) for _ in range(4))
    ctx = context or _lib.default_context()
    _lib.check(_lib.lib().scde_b200_expression_prior(
        ctx.handle, _lib.p_f64(mm), Cn, _lib.p_i32(cm), G, sq, int(length_out), float(pseudo_count), float(bw),
        float(max_quantile), float("nan") if max_value is None else float(max_value), _lib.p_f64(x), _lib.p_f64(y),
        _lib.p_f64(lp), _lib.p_f64(gw)))
    return pd.DataFrame({"x": x, "y": y, "lp": lp, "grid.weight": gw})
