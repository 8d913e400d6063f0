"""numpy restatement of scde.expression.prior / scde.failure.probability (R/functions.R:225-254, 725-750) and of R's
density.default (linear binning onto 2n points, FFT convolution with the Gaussian kernel, linear interpolation).

TEST INFRASTRUCTURE: builds the es.mef.small prior for the CPU-only fixtures (tests/helpers.py), is compared with the
oracle's C restatement (tests/test_host.py) and with the device path (tests/test_gpu_parity.py).  The product path is
scde_b200.prior (C ABI, csrc/prior.cu); nothing in the package imports this module.
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def expression_magnitude_host(models: pd.DataFrame, counts: np.ndarray) -> np.ndarray:
    """(log(counts) - corr.b)/corr.a; numpy."""
    with np.errstate(divide="ignore"):
        return (np.log(counts.astype(np.float64)) - models["corr.b"].to_numpy()[None, :]) / \
            models["corr.a"].to_numpy()[None, :]


def scde_failure_probability_host(models: pd.DataFrame, magnitudes=None, counts=None) -> np.ndarray:
    """Drop-out probability per gene (rows) and cell (columns) (R/functions.R:725-750)."""
    if magnitudes is None:
        if counts is None:
            raise ValueError("ERROR: either magnitudes or counts should be provided")
        magnitudes = expression_magnitude_host(models, np.asarray(counts))
    m = np.asarray(magnitudes, dtype=np.float64)
    ca, cb = models["conc.a"].to_numpy(), models["conc.b"].to_numpy()
    with np.errstate(over="ignore", invalid="ignore"):
        if m.ndim == 2:
            eta = m * ca[None, :] + cb[None, :]
            if "conc.a2" in models.columns:
                eta = eta + (m ** 2) * models["conc.a2"].to_numpy()[None, :]
        else:
            eta = np.outer(m, ca) + cb[None, :]
            if "conc.a2" in models.columns:
                eta = eta + np.outer(m ** 2, models["conc.a2"].to_numpy())
        x = 1.0 / (np.exp(eta) + 1.0)
    x[np.isnan(x)] = 0
    return x


def _bin_dist(x, w, lo, up, n):
    """R's BinDist: linear binning of weighted points onto n grid points, returned zero-padded to 2n."""
    y = np.zeros(2 * n)
    ixmin, ixmax = 0, n - 2
    delta = (up - lo) / (n - 1)
    ok = np.isfinite(x)
    xpos = (x[ok] - lo) / delta
    ix = np.floor(xpos).astype(np.int64)
    fx = xpos - ix
    wi = w[ok]
    inside = (ix >= ixmin) & (ix <= ixmax)
    np.add.at(y, ix[inside], wi[inside] * (1 - fx[inside]))
    np.add.at(y, ix[inside] + 1, wi[inside] * fx[inside])
    left = ix == -1
    y[0] += np.sum(wi[left] * fx[left])
    right = ix == ixmax + 1
    np.add.at(y, ix[right], wi[right] * (1 - fx[right]))
    return y


def density_gaussian(x, weights, bw, n, lo_from, hi_to):
    """stats::density(x, bw = bw, weights = weights, n = n, from = lo_from, to = hi_to) -> (x grid, y)."""
    n_user = n
    n = max(n, 512)
    if n > 512:
        n = int(2 ** np.ceil(np.log2(n)))
    lo, up = lo_from - 4 * bw, hi_to + 4 * bw
    y = _bin_dist(np.asarray(x, dtype=np.float64), np.asarray(weights, dtype=np.float64), lo, up, n)
    kords = np.linspace(0, 2 * (up - lo), 2 * n)
    kords[n + 1:2 * n] = -kords[n - 1:0:-1]
    kords = np.exp(-0.5 * (kords / bw) ** 2) / (bw * np.sqrt(2 * np.pi))
    conv = np.fft.ifft(np.fft.fft(y) * np.conj(np.fft.fft(kords))) * len(y)  # R's inverse fft is unnormalised
    kords = np.maximum(0, conv.real[:n] / len(y))
    xords = np.linspace(lo, up, n)
    xout = np.linspace(lo_from, hi_to, n_user)
    return xout, np.interp(xout, xords, kords)


def scde_expression_prior_host(models: pd.DataFrame, counts, length_out: int = 400, show_plot: bool = False,
                               pseudo_count: float = 1, bw: float = 0.1, max_quantile: float = 1, max_value=None):
    """Expression-magnitude grid (``x``, log10 scale) and prior (``y``) (R/functions.R:225-254)."""
    if isinstance(counts, pd.DataFrame):
        cm = counts.loc[:, list(models.index)].to_numpy()
    else:
        cm = np.asarray(counts)
    mag = expression_magnitude_host(models, cm)
    fail = scde_failure_probability_host(models, magnitudes=mag)
    with np.errstate(over="ignore"):
        fpkm = np.log10(np.exp(mag) + 1)
    xv = fpkm.ravel(order="F")
    wts = (1 - fail).ravel(order="F")
    wts = wts / wts.sum()
    if max_value is None:
        fin = xv[xv < np.inf]
        max_value = float(np.quantile(fin, max_quantile))  # R's default type-7 quantile
    gx, gy = density_gaussian(np.concatenate([-xv, xv]), np.concatenate([wts / 2, wts / 2]), bw, 2 * length_out + 1,
                              -max_value, max_value)
    x, y = gx[length_out:], gy[length_out:].copy()
    y[np.isnan(y)] = 0
    y = y + pseudo_count / fpkm.shape[0]
    y = y / y.sum()
    with np.errstate(divide="ignore"):
        lp = np.log(y)
    gw = np.diff(np.power(10.0, np.concatenate([[x[0]], x + np.concatenate([np.diff(x) / 2, [0]])])) - 1)
    return pd.DataFrame({"x": x, "y": y, "lp": lp, "grid.weight": gw})
