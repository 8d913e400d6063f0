"""Probe of the tcgen05 (kind::i8) contraction kernel on random integer operands, checked exactly with numpy.

    python tools/probe_i8.py [--genes 7] [--cells 90] [--rows 300] [--grid 401]

Prints whether T matches and (if not) where it differs.  Used by tests/test_gpu_i8.py."""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NV, PW, PIECE, WB, FRAC = 5, 102, 512, 128, 29
WP, KP = 104, 416
SENTINEL = -1.0e300


def pack_rows(planes, n_grid):
    """planes[n_rows][NV][n_grid] -> the kernel's row layout [piece][plane][102] (+ 2 zero bytes per piece)"""
    n_rows = planes.shape[0]
    n_pieces = (n_grid + PW - 1) // PW
    q = np.zeros((n_rows, n_pieces * PIECE), dtype=np.int8)
    for c in range(n_pieces):
        w = min(PW, n_grid - c * PW)
        for p in range(NV):
            q[:, c * PIECE + p * PW: c * PIECE + p * PW + w] = planes[:, p, c * PW: c * PW + w]
    return q


def make_problem(n_genes, n_cells, n_rows, n_grid, seed=0, max_w=5, sentinel_frac=0.3, full_lists=False):
    """sentinel_frac: fraction of the rows whose non-sentinel range is a proper sub-interval of the grid"""
    rng = np.random.default_rng(seed)
    planes = rng.integers(-128, 128, size=(n_rows, NV, n_grid), dtype=np.int64)
    lo = np.zeros(n_rows, dtype=np.int64)
    hi = np.full(n_rows, n_grid - 1, dtype=np.int64)
    clip = rng.random(n_rows) < sentinel_frac
    lo[clip] = rng.integers(0, max(1, n_grid // 3), size=int(clip.sum()))
    hi[clip] = n_grid - 1 - rng.integers(0, max(1, n_grid // 3), size=int(clip.sum())) * (rng.random(int(clip.sum())) < 0.3)
    k = np.arange(n_grid)
    dead = (k[None, :] < lo[:, None]) | (k[None, :] > hi[:, None])
    planes[np.broadcast_to(dead[:, None, :], planes.shape)] = 0  # "log 0" points carry zero digits
    row_range = (lo | (hi << 16)).astype(np.uint32)
    n_w_rows = (n_cells + 1 + 15) // 16 * 16
    w8 = np.zeros((n_w_rows, WB), dtype=np.int8)
    w8[:n_cells, :WP] = rng.integers(0, max_w + 1, size=(n_cells, WP))
    ld = (n_cells + 31) // 32 * 32
    lst_row = np.zeros((n_genes, ld), dtype=np.int32)
    lst_cell = np.full((n_genes, ld), n_cells, dtype=np.int32)  # zero W row
    lst_len = np.zeros(n_genes, dtype=np.int32)
    for g in range(n_genes):
        n = n_cells if full_lists else int(rng.integers(0, n_cells + 1))
        if g == 0:
            n = n_cells
        if g == 1 and n_genes > 2:
            n = 0
        lst_len[g] = n
        lst_row[g, :n] = rng.integers(0, n_rows, size=n)
        lst_cell[g, :n] = rng.permutation(n_cells)[:n]
    return dict(planes=planes, q=pack_rows(planes, n_grid), row_range=row_range, lo=lo, hi=hi, w8=w8, lst_row=lst_row,
                lst_cell=lst_cell, lst_len=lst_len, n_grid=n_grid, n_w_rows=n_w_rows, ld=ld)


def expected(pr):
    """(exact integer value of every (gene, boot, grid point), mask of the points some drawn row marks "log 0")"""
    G = pr["lst_len"].shape[0]
    K = pr["n_grid"]
    val = np.zeros((G, WP, K), dtype=np.int64)
    dead = np.zeros((G, WP, K), dtype=bool)
    k = np.arange(K)
    for g in range(G):
        n = pr["lst_len"][g]
        rows, cells = pr["lst_row"][g, :n], pr["lst_cell"][g, :n]
        W = pr["w8"][cells, :WP].astype(np.int64)  # [n][104]
        acc = np.zeros((WP, K), dtype=np.int64)
        for p in range(NV - 1, -1, -1):
            acc = acc * 256 + W.T @ pr["planes"][rows, p, :]
        val[g] = acc
        if n:
            drawn = W.T > 0  # [104][n]
            lo = np.where(drawn, pr["lo"][rows][None, :], 0).max(axis=1)
            hi = np.where(drawn, pr["hi"][rows][None, :], K - 1).min(axis=1)
            dead[g] = (k[None, :] < lo[:, None]) | (k[None, :] > hi[:, None])
    return val, dead


def run(ctx, pr, with_ranges=True):
    from scde_b200 import _lib

    L = _lib.lib()
    G = pr["lst_len"].shape[0]
    out = np.zeros((G, WP, KP), dtype=np.float64)
    i8p, i32p, u32p, f64p = C.POINTER(C.c_int8), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    L.scde_b200_probe_contract_i8.argtypes = [C.c_void_p, i8p, C.c_int32, C.c_int32, u32p, i8p, C.c_int32, i32p, i32p, i32p,
                                              C.c_int32, C.c_int32, f64p, i32p]
    q, w8 = np.ascontiguousarray(pr["q"]), np.ascontiguousarray(pr["w8"])
    rr = np.ascontiguousarray(pr["row_range"])
    flags = C.c_int32(0)
    r = L.scde_b200_probe_contract_i8(ctx._h, q.ctypes.data_as(i8p), q.shape[0], pr["n_grid"],
                                      rr.ctypes.data_as(u32p) if with_ranges else None, w8.ctypes.data_as(i8p),
                                      pr["n_w_rows"], pr["lst_row"].ctypes.data_as(i32p),
                                      pr["lst_cell"].ctypes.data_as(i32p), pr["lst_len"].ctypes.data_as(i32p), G,
                                      pr["ld"], out.ctypes.data_as(f64p), C.byref(flags))
    _lib.check(r)
    pr["flags"] = int(flags.value)
    return out


def compare(pr, out, with_ranges=True):
    val, dead = expected(pr)
    K = pr["n_grid"]
    want = val.astype(np.float64) * 2.0 ** -FRAC
    got = out[:, :, :K]
    clean = ~dead if with_ranges else np.ones_like(dead)
    ok_clean = got[clean] == want[clean]
    sent_ok = bool(np.all(got[~clean] == SENTINEL))
    # flag 4 exactly when some (gene, real boot) has no admissible grid point left
    empty = bool(np.any(dead[:, :WP, :].all(axis=2))) if with_ranges else False
    flag_ok = bool(((pr.get("flags", 0) & 4) != 0) == empty)
    return bool(ok_clean.all()) and sent_ok and flag_ok, got, want, clean


def report(pr, out, label):
    ok, got, want, clean = compare(pr, out)
    print(f"[{label}] match={ok} flags={pr.get('flags')}")
    if ok:
        return True
    bad = (got != want) & clean
    print(f"   mismatching clean elements: {bad.sum()} of {clean.sum()}; sentinel elements wrong: "
          f"{int((got[~clean] != SENTINEL).sum())} of {int((~clean).sum())}")
    G = got.shape[0]
    for g in range(min(G, 4)):
        bg = bad[g]
        print(f"   gene {g} len {pr['lst_len'][g]}: bad boots {np.unique(np.nonzero(bg)[0])[:12]} ... "
              f"bad grid {np.unique(np.nonzero(bg)[1])[:24]} ...")
        if bg.any():
            b, k = np.argwhere(bg)[0]
            print(f"      first bad (b={b}, k={k}): got {got[g, b, k] * 2.0 ** FRAC:.0f} want {want[g, b, k] * 2.0 ** FRAC:.0f}")
    g = 0
    print("   got[0, 0, :8] * 2^29 =", (got[g, 0, :8] * 2.0 ** FRAC).astype(np.int64))
    print("   want[0, 0, :8] * 2^29 =", (want[g, 0, :8] * 2.0 ** FRAC).astype(np.int64))
    return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=7)
    ap.add_argument("--cells", type=int, default=90)
    ap.add_argument("--rows", type=int, default=300)
    ap.add_argument("--grid", type=int, default=401)
    ap.add_argument("--simple", action="store_true", help="planes: only plane 0 non-zero")
    a = ap.parse_args()
    from scde_b200 import _lib

    ctx = _lib.Context(0)
    pr = make_problem(a.genes, a.cells, a.rows, a.grid)
    if a.simple:
        pr["planes"][:, 1:, :] = 0
        pr["q"] = pack_rows(pr["planes"], a.grid)
    out = run(ctx, pr)
    return 0 if report(pr, out, "sw128") else 1


if __name__ == "__main__":
    sys.exit(main())
