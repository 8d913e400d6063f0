"""ctypes binding of the CPU oracle (``oracle/scde_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of
``bench.py``; never by the product package ``scde_b200``.  Parity status (pinned bit for bit against the reference's own
C++ compiled against a header shim, ``oracle/ref.py``; R nmath / R-level code stay restated): see the header of
``scde_oracle.c`` and DESIGN.md.

All matrices follow R's column-major convention, expressed here as Fortran-ordered numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libscde_oracle.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "scde_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        for name in ("orc_stirlerr", "orc_qnorm_upper", "orc_pnorm_upper"):
            getattr(_lib, name).restype = C.c_double
            getattr(_lib, name).argtypes = [C.c_double]
        _lib.orc_bd0.restype = C.c_double
        _lib.orc_bd0.argtypes = [C.c_double, C.c_double]
        _lib.orc_dnbinom_log.restype = C.c_double
        _lib.orc_dnbinom_log.argtypes = [C.c_double] * 3
        _lib.orc_dpois_log.restype = C.c_double
        _lib.orc_dpois_log.argtypes = [C.c_double] * 2
    return _lib


def _d(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


def _f64(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _i32(a):
    return np.asfortranarray(np.asarray(a, dtype=np.int32))


# ---------------------------------------------------------------- scalar functions
def dnbinom_log(x, size, prob):
    return lib().orc_dnbinom_log(float(x), float(size), float(prob))


def dpois_log(x, lam):
    return lib().orc_dpois_log(float(x), float(lam))


def qnorm_upper(p):
    return lib().orc_qnorm_upper(float(p))


def pnorm_upper(x):
    return lib().orc_pnorm_upper(float(x))


def stirlerr(n):
    return lib().orc_stirlerr(float(n))


def bd0(x, np_):
    return lib().orc_bd0(float(x), float(np_))


# ---------------------------------------------------------------- bootstrap indices (libc srand/rand)
def boot_indices(seed: int, n: int, nboot: int) -> np.ndarray:
    out = np.empty((nboot, n), dtype=np.int32)
    lib().orc_boot_indices(C.c_int(seed), C.c_int(n), C.c_int(nboot), _i(out))
    return out


def flatten_pools(pools):
    off = np.zeros(len(pools) + 1, dtype=np.int32)
    for k, p in enumerate(pools):
        off[k + 1] = off[k] + len(p)
    cells = np.concatenate([np.asarray(p, dtype=np.int32) for p in pools]) if len(pools) else np.zeros(0, np.int32)
    return off, np.ascontiguousarray(cells, dtype=np.int32)


def batch_boot_indices(seed: int, pools, comp, nboot: int) -> np.ndarray:
    off, cells = flatten_pools(pools)
    comp = _i32(comp)
    D = int(np.sum(comp[comp > 0]))
    out = np.empty((nboot, D), dtype=np.int32)
    lib().orc_batch_boot_indices(C.c_int(seed), C.c_int(len(pools)), _i(off), _i(cells), _i(comp), C.c_int(nboot), _i(out))
    return out


# ---------------------------------------------------------------- host prep
def unique_counts(counts):
    """``ucl`` (flat + offsets) and ``uci`` as scde.posteriors builds them (R/functions.R:631-632)."""
    counts = _i32(counts)
    G, Cn = counts.shape
    flat = np.empty(G * Cn, dtype=np.int32)
    off = np.empty(Cn + 1, dtype=np.int32)
    uci = np.empty((G, Cn), dtype=np.int32, order="F")
    lib().orc_unique_counts(_i(counts), C.c_int(G), C.c_int(Cn), _i(flat), _i(off), _i(uci))
    return flat[: off[-1]].copy(), off, uci


def cell_table(model_row12, uc, mag, localtheta=0, sqlogit=0, ncells_for_clamp=1):
    """One cell's log-posterior table: returns (K x U array, modes[U])."""
    m = _f64(model_row12).reshape(12)
    uc = _i32(uc)
    mag = _f64(mag)
    K, U = len(mag), len(uc)
    out = np.empty((K, U), dtype=np.float64, order="F")
    modes = np.empty(U, dtype=np.int32)
    lib().orc_cell_table(_d(m), _i(uc), C.c_int(U), _d(mag), C.c_int(K), C.c_int(localtheta), C.c_int(sqlogit),
                         C.c_int(ncells_for_clamp), _d(out), _i(modes))
    return out, modes


# ---------------------------------------------------------------- native entry points
def log_boot_posterior(models, ucl_flat, ucl_off, uci, mag, nboot, seed=1, boot_idx=None, returnpost=0,
                       localtheta=0, sqlogit=0, ensemble=0):
    """logBootPosterior (src/jpmatLogBoot.cpp:100).  Returns dict(jp[, modes][, post])."""
    models = _f64(models)
    ncells = models.shape[0]
    assert models.shape[1] == 12
    uci = _i32(uci)
    G = uci.shape[0]
    mag = _f64(mag)
    K = len(mag)
    ucl_flat, ucl_off = _i32(ucl_flat), _i32(ucl_off)
    bi = None if boot_idx is None else np.ascontiguousarray(boot_idx, dtype=np.int32)
    jp = np.empty((G, K), dtype=np.float64, order="F")
    modes = np.empty((G, ncells), dtype=np.float64, order="F") if returnpost in (1, 3) else None
    post = np.empty((ncells, K, G), dtype=np.float64) if returnpost in (2, 3) else None
    lib().orc_log_boot_posterior(_d(models), C.c_int(ncells), _i(ucl_flat), _i(ucl_off), _i(uci), C.c_int(G), _d(mag),
                                 C.c_int(K), C.c_int(nboot), C.c_int(seed), _i(bi), C.c_int(returnpost),
                                 C.c_int(localtheta), C.c_int(sqlogit), C.c_int(ensemble), _d(jp), _d(modes), _d(post))
    out = {"jp": jp}
    if modes is not None:
        out["modes"] = modes
    if post is not None:  # ncells matrices, each G x K column-major
        out["post"] = [post[i].T for i in range(ncells)]
    return out


def log_boot_batch_posterior(models, ucl_flat, ucl_off, uci, mag, pools, comp, nboot, seed=1, boot_idx=None,
                             returnpost=0, localtheta=0, sqlogit=0):
    """logBootBatchPosterior (src/jpmatLogBoot.cpp:343)."""
    models = _f64(models)
    ncells = models.shape[0]
    uci = _i32(uci)
    G = uci.shape[0]
    mag = _f64(mag)
    K = len(mag)
    ucl_flat, ucl_off = _i32(ucl_flat), _i32(ucl_off)
    off, cells = flatten_pools(pools)
    comp = _i32(comp)
    bi = None if boot_idx is None else np.ascontiguousarray(boot_idx, dtype=np.int32)
    jp = np.empty((G, K), dtype=np.float64, order="F")
    modes = np.empty((G, ncells), dtype=np.float64, order="F") if returnpost == 1 else None
    post = np.empty((ncells, K, G), dtype=np.float64) if returnpost in (2, 3) else None
    lib().orc_log_boot_batch_posterior(_d(models), C.c_int(ncells), _i(ucl_flat), _i(ucl_off), _i(uci), C.c_int(G),
                                       _d(mag), C.c_int(K), C.c_int(len(pools)), _i(off), _i(cells), _i(comp),
                                       C.c_int(nboot), C.c_int(seed), _i(bi), C.c_int(returnpost), C.c_int(localtheta),
                                       C.c_int(sqlogit), _d(jp), _d(modes), _d(post))
    out = {"jp": jp}
    if modes is not None:
        out["modes"] = modes
    if post is not None:
        out["post"] = [post[i].T for i in range(ncells)]
    return out


def jpmat_log_boot(matl, nboot, seed=1, boot_idx=None):
    """Legacy jpmatLogBoot (src/jpmatLogBoot.cpp:11): matl = list of nrows x ncols log matrices."""
    nmat = len(matl)
    nrows, ncols = matl[0].shape
    stack = np.empty((nmat, ncols, nrows), dtype=np.float64)
    for i, m in enumerate(matl):
        stack[i] = np.asarray(m, dtype=np.float64).T
    bi = None if boot_idx is None else np.ascontiguousarray(boot_idx, dtype=np.int32)
    jp = np.empty((nrows, ncols), dtype=np.float64, order="F")
    lib().orc_jpmat_log_boot(_d(stack), C.c_int(nmat), C.c_int(nrows), C.c_int(ncols), C.c_int(nboot), C.c_int(seed),
                             _i(bi), _d(jp))
    return jp


def jpmat_log_batch_boot(matll, comp, nboot, seed=1, boot_idx=None):
    """Legacy jpmatLogBatchBoot (src/jpmatLogBoot.cpp:48): matll = list (per pool) of lists of matrices."""
    flat = [m for pool in matll for m in pool]
    off = np.zeros(len(matll) + 1, dtype=np.int32)
    for k, pool in enumerate(matll):
        off[k + 1] = off[k] + len(pool)
    nrows, ncols = flat[0].shape
    stack = np.empty((len(flat), ncols, nrows), dtype=np.float64)
    for i, m in enumerate(flat):
        stack[i] = np.asarray(m, dtype=np.float64).T
    comp = _i32(comp)
    bi = None if boot_idx is None else np.ascontiguousarray(boot_idx, dtype=np.int32)
    jp = np.empty((nrows, ncols), dtype=np.float64, order="F")
    lib().orc_jpmat_log_batch_boot(_d(stack), C.c_int(len(matll)), _i(off), _i(comp), C.c_int(nrows), C.c_int(ncols),
                                   C.c_int(nboot), C.c_int(seed), _i(bi), _d(jp))
    return jp


def mat_slide_mult(m1, m2):
    m1, m2 = _f64(m1), _f64(m2)
    nrows, n = m1.shape
    out = np.empty((nrows, 2 * n - 1), dtype=np.float64, order="F")
    lib().orc_mat_slide_mult(_d(m1), _d(m2), C.c_int(nrows), C.c_int(n), _d(out))
    return out


def ratio_posterior(pmat1, pmat2, prior_y=None):
    """calculate.ratio.posterior (R/functions.R:3491); prior_y=None == skip.prior.adjustment."""
    p1, p2 = _f64(pmat1), _f64(pmat2)
    G, K = p1.shape
    py = None if prior_y is None else _f64(prior_y)
    out = np.empty((G, 2 * K - 1), dtype=np.float64, order="F")
    lib().orc_ratio_posterior(_d(p1), _d(p2), C.c_int(G), C.c_int(K), _d(py), _d(out))
    return out


def fold_change_grid(x):
    """seq(x[1]-x[n], x[n]-x[1], length = 2n-1) as R computes it (from + i*by), R/functions.R:3506."""
    x = np.asarray(x, dtype=np.float64)
    n = 2 * len(x) - 1
    lo, hi = x[0] - x[-1], x[-1] - x[0]
    by = (hi - lo) / (n - 1)
    return lo + np.arange(n, dtype=np.float64) * by


def distribution_summary(s_bdiffp, diffv, expectation=0.0):
    """quick.distribution.summary (R/functions.R:5039).  Returns (G x 6 array, G x 3 int index array)."""
    s = _f64(s_bdiffp)
    G, n = s.shape
    diffv = _f64(diffv)
    ex = np.atleast_1d(np.asarray(expectation, dtype=np.float64))
    out = np.empty((G, 6), dtype=np.float64, order="F")
    idx = np.empty((G, 3), dtype=np.int32, order="F")
    lib().orc_distribution_summary(_d(s), C.c_int(G), C.c_int(n), _d(diffv), _d(ex), C.c_int(len(ex)), _d(out), _i(idx))
    return out, idx


def p_adjust_bh(p):
    p = _f64(p)
    out = np.empty_like(p)
    lib().orc_p_adjust_bh(_d(p), C.c_int(len(p)), _d(out))
    return out


def expression_magnitude(counts, corr_b, corr_a):
    counts = _i32(counts)
    G, Cn = counts.shape
    out = np.empty((G, Cn), dtype=np.float64, order="F")
    lib().orc_expression_magnitude(_i(counts), C.c_int(G), C.c_int(Cn), _d(_f64(corr_b)), _d(_f64(corr_a)), _d(out))
    return out


def density_gaussian(x, w, bw, n, lo_from, hi_to):
    """stats::density(x, bw = bw, weights = w, n = n, from = lo_from, to = hi_to) -> (x grid, y); published algorithm
    restated (R is not in the reference tree)."""
    x, w = _f64(x), _f64(w)
    xo, yo = np.empty(n), np.empty(n)
    lib().orc_density_gaussian(_d(x), _d(w), C.c_long(len(x)), C.c_double(bw), C.c_int(n), C.c_double(lo_from),
                               C.c_double(hi_to), _d(xo), _d(yo))
    return xo, yo


def failure_probability(mag, conc_a, conc_b, conc_a2=None):
    """scde.failure.probability with a genes x cells magnitude matrix (R/functions.R:736-748)."""
    mag = _f64(mag)
    G, Cn = mag.shape
    out = np.empty((G, Cn), dtype=np.float64, order="F")
    a2 = None if conc_a2 is None else _f64(conc_a2)
    lib().orc_failure_probability(_d(mag), C.c_int(G), C.c_int(Cn), _d(_f64(conc_a)), _d(_f64(conc_b)), _d(a2), _d(out))
    return out


def expression_prior(models_df, counts, length_out=400, pseudo_count=1.0, bw=0.1, max_quantile=1.0, max_value=None):
    """scde.expression.prior (R/functions.R:225-254) -> dict(x, y, lp, grid.weight)."""
    counts = _i32(counts)
    G, Cn = counts.shape
    mm, lt, sq = pack_models(models_df)
    K = length_out + 1
    x, y, lp, gw = (np.empty(K) for _ in range(4))
    lib().orc_expression_prior(_i(counts), C.c_int(G), C.c_int(Cn), _d(mm), C.c_int(sq), C.c_int(length_out),
                               C.c_double(pseudo_count), C.c_double(bw), C.c_double(max_quantile),
                               C.c_double(np.nan if max_value is None else max_value), _d(x), _d(y), _d(lp), _d(gw))
    return {"x": x, "y": y, "lp": lp, "grid.weight": gw}


def posteriors_chunked(models, counts, mag, nboot, boot_idx, nthreads, return_times=False):
    """Gene-chunked CPU arm (R/functions.R:606-617 semantics with shared boot_idx)."""
    models = _f64(models)
    counts = _i32(counts)
    G, Cn = counts.shape
    mag = _f64(mag)
    bi = np.ascontiguousarray(boot_idx, dtype=np.int32)
    jp = np.empty((G, len(mag)), dtype=np.float64, order="F")
    times = np.zeros(2)
    lib().orc_posteriors_chunked(_d(models), C.c_int(Cn), _i(counts), C.c_int(G), _d(mag), C.c_int(len(mag)),
                                 C.c_int(nboot), _i(bi), C.c_int(bi.shape[1]), C.c_int(nthreads), _d(jp), _d(times))
    return (jp, times) if return_times else jp


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ---------------------------------------------------------------- R-level composition (oracle side)
def marginals_from_prior_x(x):
    """R/functions.R:575-577."""
    m = np.power(10.0, np.asarray(x, dtype=np.float64)) - 1.0
    m[m < 0] = 0
    with np.errstate(divide="ignore"):
        return np.log(m)


def pack_models(models_df):
    """R/functions.R:579-583, 601-604: clamp corr.a, 12-column matrix with NaN for absent columns."""
    mn = ["conc.b", "conc.a", "fail.r", "corr.b", "corr.a", "corr.theta", "corr.ltheta.b", "corr.ltheta.t",
          "corr.ltheta.m", "corr.ltheta.s", "corr.ltheta.r", "conc.a2"]
    mm = np.full((len(models_df), 12), np.nan, dtype=np.float64, order="F")
    for j, name in enumerate(mn):
        if name in models_df.columns:
            mm[:, j] = np.asarray(models_df[name], dtype=np.float64)
    ca = mm[:, 4]
    ca[ca < 1e-10] = 1e-10
    return mm, int("corr.ltheta.b" in models_df.columns), int("conc.a2" in models_df.columns)


def expression_difference(models_df, counts, prior_x, prior_y, group_idx, nboot=100, seed=1, batch_codes=None,
                          expectation=0.0, batch_models_df=None):
    """scde.expression.difference with n.cores = 1 (R/functions.R:304-408) on numpy inputs.

    counts: G x C (cells ordered as model rows); group_idx: (idx0, idx1) integer arrays of the two factor levels;
    batch_codes: optional per-cell integer batch level codes (0..L-1; negative = NA, dropped from the pools and the
    composition as tapply / table do); batch_models_df: batch.models (default: the same models).
    """
    counts = _i32(counts)
    mm, lt, sq = pack_models(models_df)
    mag = marginals_from_prior_x(prior_x)
    jpl = []
    for ii in group_idx:
        sub = np.asfortranarray(counts[:, ii])
        flat, off, uci = unique_counts(sub)
        jpl.append(log_boot_posterior(np.asfortranarray(mm[ii]), flat, off, uci, mag, nboot, seed=seed,
                                      localtheta=lt, sqlogit=sq)["jp"])
    bdiffp = ratio_posterior(jpl[0], jpl[1], prior_y)
    diffv = fold_change_grid(prior_x)
    res, idx = distribution_summary(bdiffp, diffv, expectation)
    out = {"results": res, "idx": idx, "difference.posterior": bdiffp, "joint.posteriors": jpl}
    if batch_codes is not None:
        batch_codes = np.asarray(batch_codes)
        L = int(batch_codes.max()) + 1
        pools = [np.nonzero(batch_codes == l)[0].astype(np.int32) for l in range(L)]
        flat, off, uci = unique_counts(counts)
        bjpl = []
        bmm, blt, bsq = (mm, lt, sq) if batch_models_df is None else pack_models(batch_models_df)
        for ii in group_idx:
            bc = batch_codes[ii]
            comp = np.bincount(bc[bc >= 0], minlength=L).astype(np.int32)
            bjpl.append(log_boot_batch_posterior(bmm, flat, off, uci, mag, pools, comp, nboot, seed=seed,
                                                 localtheta=blt, sqlogit=bsq)["jp"])
        bb = ratio_posterior(bjpl[0], bjpl[1], prior_y)
        bres, bidx = distribution_summary(bb, diffv, 0.0)
        ab = ratio_posterior(bdiffp, bb, None)
        # as.numeric(colnames(bdiffp)) feeds seq() again: R/functions.R:391,3506
        adiffv = fold_change_grid(diffv)
        ares, aidx = distribution_summary(ab, adiffv, expectation)
        out.update({"batch.effect": bres, "batch.effect.idx": bidx, "batch.adjusted": ares,
                    "batch.adjusted.idx": aidx, "batch.adjusted.difference.posterior": ab,
                    "batch.joint.posteriors": bjpl, "batch.difference.posterior": bb})
    return out
