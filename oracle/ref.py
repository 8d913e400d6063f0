"""ctypes binding of oracle/_ref/libscde_ref.so -- THE REFERENCE'S OWN C++ (src/jpmatLogBoot.cpp, src/matSlideMult.cpp),
compiled unmodified against the header shim in oracle/shim/ (recipe: ``make -C oracle ref``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, by ``tests/golden/make_ref_fixtures.py`` and by the CPU legs of
``bench.py`` (``--impl reference`` / ``cpu_baseline``); never by the product package ``scde_b200``.

The library is built in this container, where /root/reference exists, and travels to the GPU box prebuilt
(``oracle/_ref/`` is git-ignored, not gpurun-ignored).  ``available()`` says whether it is there.
Function signatures mirror ``oracle.oracle`` so the two can be called side by side; the draws always come from the
reference's own ``srand(seed)`` / ``rand()`` calls (there is no boot_idx argument in the reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libscde_ref.so")
REF_SRC = os.environ.get("SCDE_REF_SRC", "/root/reference/src")
_lib = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def can_build() -> bool:
    return os.path.exists(os.path.join(REF_SRC, "jpmatLogBoot.cpp")) and os.path.exists(os.path.join(REF_SRC, "matSlideMult.cpp"))


def build(force: bool = False) -> str | None:
    """Compile the reference sources where they lie (only possible where /root/reference exists)."""
    if not can_build():
        return _LIB_PATH if os.path.exists(_LIB_PATH) else None
    if force and os.path.exists(_LIB_PATH):
        os.remove(_LIB_PATH)
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref", f"REF_SRC={REF_SRC}"], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    return _LIB_PATH


def available() -> bool:
    return os.path.exists(_LIB_PATH) or can_build()


def lib():
    global _lib
    if _lib is None:
        if can_build():
            build()
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing and the reference sources are not here to build it")
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _d(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


def _f64(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _i32(a):
    return np.asfortranarray(np.asarray(a, dtype=np.int32))


def log_boot_posterior(models, ucl_flat, ucl_off, uci, mag, nboot, seed=1, returnpost=0, localtheta=0, sqlogit=0,
                       ensemble=0):
    """The reference's logBootPosterior (src/jpmatLogBoot.cpp:100).  Returns dict(jp[, modes][, post])."""
    models = _f64(models)
    ncells = models.shape[0]
    assert models.shape[1] == 12
    uci = _i32(uci)
    G = uci.shape[0]
    mag = _f64(mag)
    K = len(mag)
    ucl_flat, ucl_off = _i32(ucl_flat), _i32(ucl_off)
    jp = np.empty((G, K), dtype=np.float64, order="F")
    modes = np.empty((G, ncells), dtype=np.float64, order="F") if returnpost in (1, 3) else None
    post = np.empty((ncells, K, G), dtype=np.float64) if returnpost in (2, 3) else None
    rc = lib().ref_log_boot_posterior(_d(models), C.c_int(ncells), _i(ucl_flat), _i(ucl_off), _i(uci), C.c_int(G), _d(mag),
                                      C.c_int(K), C.c_int(nboot), C.c_int(seed), C.c_int(returnpost), C.c_int(localtheta),
                                      C.c_int(sqlogit), C.c_int(ensemble), _d(jp), _d(modes), _d(post))
    assert rc == 0
    out = {"jp": jp}
    if modes is not None:
        out["modes"] = modes
    if post is not None:
        out["post"] = [post[i].T for i in range(ncells)]
    return out


def log_boot_batch_posterior(models, ucl_flat, ucl_off, uci, mag, pools, comp, nboot, seed=1, returnpost=0, localtheta=0,
                             sqlogit=0):
    """The reference's logBootBatchPosterior (src/jpmatLogBoot.cpp:343)."""
    from .oracle import flatten_pools

    models = _f64(models)
    ncells = models.shape[0]
    uci = _i32(uci)
    G = uci.shape[0]
    mag = _f64(mag)
    K = len(mag)
    ucl_flat, ucl_off = _i32(ucl_flat), _i32(ucl_off)
    off, cells = flatten_pools(pools)
    comp = _i32(comp)
    jp = np.empty((G, K), dtype=np.float64, order="F")
    modes = np.empty((G, ncells), dtype=np.float64, order="F") if returnpost == 1 else None
    post = np.empty((ncells, K, G), dtype=np.float64) if returnpost == 2 else None
    rc = lib().ref_log_boot_batch_posterior(_d(models), C.c_int(ncells), _i(ucl_flat), _i(ucl_off), _i(uci), C.c_int(G),
                                            _d(mag), C.c_int(K), C.c_int(len(pools)), _i(off), _i(cells), _i(comp),
                                            C.c_int(nboot), C.c_int(seed), C.c_int(returnpost), C.c_int(localtheta),
                                            C.c_int(sqlogit), _d(jp), _d(modes), _d(post))
    assert rc == 0
    out = {"jp": jp}
    if modes is not None:
        out["modes"] = modes
    if post is not None:
        out["post"] = [post[i].T for i in range(ncells)]
    return out


def _stack(mats):
    nrows, ncols = mats[0].shape
    stack = np.empty((len(mats), ncols, nrows), dtype=np.float64)
    for i, m in enumerate(mats):
        stack[i] = np.asarray(m, dtype=np.float64).T
    return stack, nrows, ncols


def jpmat_log_boot(matl, nboot, seed=1):
    stack, nrows, ncols = _stack(matl)
    jp = np.empty((nrows, ncols), dtype=np.float64, order="F")
    lib().ref_jpmat_log_boot(_d(stack), C.c_int(len(matl)), C.c_int(nrows), C.c_int(ncols), C.c_int(nboot), C.c_int(seed), _d(jp))
    return jp


def jpmat_log_batch_boot(matll, comp, nboot, seed=1):
    flat = [m for pool in matll for m in pool]
    off = np.zeros(len(matll) + 1, dtype=np.int32)
    for k, pool in enumerate(matll):
        off[k + 1] = off[k] + len(pool)
    stack, nrows, ncols = _stack(flat)
    comp = _i32(comp)
    jp = np.empty((nrows, ncols), dtype=np.float64, order="F")
    lib().ref_jpmat_log_batch_boot(_d(stack), C.c_int(len(matll)), _i(off), _i(comp), C.c_int(nrows), C.c_int(ncols),
                                   C.c_int(nboot), C.c_int(seed), _d(jp))
    return jp


def mat_slide_mult(m1, m2):
    m1, m2 = _f64(m1), _f64(m2)
    nrows, n = m1.shape
    out = np.empty((nrows, 2 * n - 1), dtype=np.float64, order="F")
    lib().ref_mat_slide_mult(_d(m1), _d(m2), C.c_int(nrows), C.c_int(n), _d(out))
    return out
