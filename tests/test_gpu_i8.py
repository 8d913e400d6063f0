"""The tcgen05 (kind::i8) contraction kernel: exact integer checks of the kernel alone (random operands through the probe
entry of the C ABI), and the fixed-point path against the FP64 kernel / the oracle at the 1e-6 contract."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O
from scde_b200 import _lib, api, synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import probe_i8 as P  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("genes,cells,rows,grid", [
    (7, 90, 300, 401),     # the default grid: four 512-byte pieces (102 + 102 + 102 + 95 grid points)
    (3, 700, 900, 401),    # lists of 22 stages: the 10-slot ring wraps twice within an item
    (400, 40, 64, 401),    # 1600 items on 148 CTAs: many items per CTA, accumulator hand-over every item
    (5, 33, 50, 201),      # two pieces
    (5, 33, 50, 102),      # exactly one piece
    (5, 33, 50, 16),       # a single, mostly empty piece
    (4, 64, 40, 408),      # the largest grid the kernel takes
    (2, 1, 3, 401),        # one-entry lists
])
def test_kernel_exact_on_random_integers(ctx, genes, cells, rows, grid):
    pr = P.make_problem(genes, cells, rows, grid, seed=genes + cells)
    out = P.run(ctx, pr)
    ok, got, want, clean = P.compare(pr, out)
    assert ok, (f"{int(((got != want) & clean).sum())} of {int(clean.sum())} elements differ, "
                f"{int((got[~clean] != P.SENTINEL).sum())} sentinel points wrong, flags {pr['flags']}")


def test_kernel_without_ranges(ctx):
    """row_range = NULL: no row holds a sentinel, no range kernel runs"""
    pr = P.make_problem(6, 70, 100, 401, seed=9, sentinel_frac=0.0)
    out = P.run(ctx, pr, with_ranges=False)
    ok, got, want, clean = P.compare(pr, out, with_ranges=False)
    assert ok


def test_kernel_extreme_digits(ctx):
    """all digits at -128 / 127 and multiplicities at 127: the int32 accumulators hold |sum| <= 128 * 127 * entries"""
    pr = P.make_problem(3, 200, 16, 401, seed=5, max_w=127, sentinel_frac=0.0, full_lists=True)
    pr["planes"][:] = np.where(np.random.default_rng(1).random(pr["planes"].shape) < 0.5, -128, 127)
    pr["q"] = P.pack_rows(pr["planes"], 401)
    out = P.run(ctx, pr)
    ok, got, want, clean = P.compare(pr, out)
    assert ok


def test_kernel_flags_empty_sentinel_intersection(ctx):
    """two rows whose admissible grid ranges do not overlap, both drawn by every boot: flag 4 (the caller reruns in FP64)"""
    pr = P.make_problem(2, 4, 4, 401, seed=1, max_w=1, sentinel_frac=0.0, full_lists=True)
    pr["w8"][:4, :P.WP] = 1
    pr["lo"][:] = [0, 300, 0, 0]
    pr["hi"][:] = [100, 400, 400, 400]
    pr["row_range"] = (pr["lo"] | (pr["hi"] << 16)).astype(np.uint32)
    pr["lst_row"][:, :4] = [0, 1, 2, 3]
    out = P.run(ctx, pr)
    assert pr["flags"] & 4
    ok, got, want, clean = P.compare(pr, out)
    assert ok


def _err(a, b, floor=1e-290):
    big = (a > floor) & (b > floor)
    la, lb = np.log(a[big]), np.log(b[big])
    return float(np.max(np.abs(la - lb) / np.maximum(np.abs(lb), 1.0)))


@pytest.mark.parametrize("G,Cn", [(200, 64), (60, 400)])
def test_fixed_point_path_vs_fp64_kernel(ctx, G, Cn):
    """same inputs, same draws: joint posteriors of the fixed-point path within 1e-6 (relative, log scale) of the FP64
    kernel, grid indices of the summary identical, and the run is reproducible bit for bit"""
    w = synth.make_workload(3, n_genes=G, n_cells=Cn, seed=11)
    res = {}
    for kernel in (2, 3, 3):
        ctx.set_contract_kernel(kernel)
        try:
            r = api.scde_expression_difference(w.models, w.counts, w.prior, groups=w.groups, n_randomizations=100,
                                               return_posteriors=True, context=ctx)
        finally:
            ctx.set_contract_kernel(0)
        res.setdefault(kernel, []).append(r)
    a, b, b2 = res[2][0], res[3][0], res[3][1]
    for lev in ("g1", "g2"):
        e = _err(b["joint.posteriors"][lev].to_numpy(), a["joint.posteriors"][lev].to_numpy())
        assert e < 1e-6, e
        assert np.array_equal(b["joint.posteriors"][lev].to_numpy(), b2["joint.posteriors"][lev].to_numpy())
    assert np.array_equal(a["results"][["lb", "mle", "ub"]].to_numpy(), b["results"][["lb", "mle", "ub"]].to_numpy())
    za, zb = a["results"]["Z"].to_numpy(), b["results"]["Z"].to_numpy()
    assert np.all(np.abs(za - zb) <= np.where(za < -6.0, 2e-4, 1e-6 * np.maximum(1.0, np.abs(za)) + 1e-9))


def test_large_counts_sentinel_ranges(ctx):
    """counts in the tens of thousands put the reference's "log 0" clamp (-DBL_MAX/n/1.1) on the low end of the grid: the
    per-row ranges have to reproduce the FP64 path's hard exclusion of those grid points"""
    w = synth.make_workload(3, n_genes=80, n_cells=30, seed=3)
    counts = np.array(w.counts, copy=True)
    rng = np.random.default_rng(0)
    counts[:40] = rng.integers(0, 60000, size=counts[:40].shape)
    counts[40:60, ::3] = 0
    mm, lt, sq = api.pack_models(w.models)
    mag = api.marginals_from_prior(w.prior)
    flat, off, uci = O.unique_counts(counts)
    bi = O.boot_indices(1, counts.shape[1], 100)
    want = O.log_boot_posterior(mm, flat, off, uci, mag, 100, boot_idx=bi)["jp"]
    ctx.set_contract_kernel(3)
    try:
        got = api.scde_posteriors(w.models, counts, w.prior, n_randomizations=100, context=ctx).to_numpy()
    finally:
        ctx.set_contract_kernel(0)
    assert _err(got, want) < 1e-6
    assert np.all(got[want == 0.0] < 1e-280)


def test_multiplicity_above_127_falls_back_to_fp64(ctx):
    """a cell drawn 150 times in every randomization does not fit the int8 operand: the library reruns on the FP64 kernel"""
    w = synth.make_workload(3, n_genes=20, n_cells=150, seed=2)
    bi = np.zeros((10, 150), dtype=np.int32)  # every draw hits cell 0
    res = {}
    for kernel in (0, 2):
        ctx.set_contract_kernel(kernel)
        try:
            res[kernel] = api.scde_posteriors(w.models, w.counts, w.prior, n_randomizations=10, boot_idx=bi,
                                              context=ctx).to_numpy()
        finally:
            ctx.set_contract_kernel(0)
    assert np.array_equal(res[0], res[2])


@pytest.mark.parametrize("length_out", [400, 200, 407])
def test_register_resident_row_kernel_equals_per_element_kernel(ctx, length_out):
    """lp_rows_q_kernel (cell vectors in registers, 13 grid points per lane) against lp_rows_fast_kernel (the same rows
    built with per-element loads; scde_b200_options.lp_rows_kernel = 1): the fixed-point tables may differ by one unit of 2^-29 where a
    value sits on a rounding boundary (the row sum is associated differently), so the joint posteriors agree to 1e-7;
    counts up to 60000 exercise the snap point on every part of the grid, the sentinel ranges and the slow band; grids
    of 201 and 408 points exercise the lanes without a point of their own and the largest grid of the fixed-point form."""
    w = synth.make_workload(3, n_genes=120, n_cells=40, seed=11)
    prior = synth.make_prior(120, length_out=length_out)
    counts = np.array(w.counts, copy=True)
    rng = np.random.default_rng(5)
    counts[:30] = rng.integers(0, 60000, size=counts[:30].shape)
    counts[30:50] = rng.integers(0, 40, size=counts[30:50].shape)
    res = {}
    for old in (1, 0):
        keep = ctx.set_options(lp_rows_kernel=old)
        try:
            res[old] = api.scde_posteriors(w.models, counts, prior, n_randomizations=50, context=ctx).to_numpy()
        finally:
            ctx.restore_options(keep)
    assert _err(res[0], res[1]) < 1e-7
    assert np.array_equal(res[0] == 0.0, res[1] == 0.0)


def test_item_order_and_l2_hints_do_not_change_results(ctx):
    """The tcgen05 kernel's schedule (gene-major or piece-major items) and the L2 eviction hints on the table loads
    (scde_b200_options.item_order / hot_rank / cold_evict_first) are performance knobs: the integer sums, and therefore
    the posteriors, are bit-identical.  137 randomizations = two passes of 104 boots with a ragged last group."""
    w = synth.make_workload(3, n_genes=150, n_cells=48, seed=4)
    counts = np.array(w.counts, copy=True)
    counts[:20] = np.random.default_rng(2).integers(0, 30000, size=counts[:20].shape)
    res = []
    for kw in (dict(item_order=0, hot_rank=-1), dict(item_order=1, hot_rank=-1), dict(item_order=1, hot_rank=8),
               dict(item_order=0, hot_rank=2, cold_evict_first=0)):
        keep = ctx.set_options(**kw)
        try:
            res.append(api.scde_posteriors(w.models, counts, w.prior, n_randomizations=137, context=ctx).to_numpy())
        finally:
            ctx.restore_options(keep)
    for r in res[1:]:
        assert np.array_equal(res[0], r)


@pytest.mark.parametrize("item_order", [0, 1])
def test_short_and_long_lists_mixed_in_any_order(ctx, item_order):
    """Genes whose entry lists take 0, 1, 2 and 3 ring stages (0 .. 70 non-zero cells), shuffled so that the schedule meets
    them in every succession, in both item orders: with two producer groups the group without a stage of its own in a run
    of short items ran ahead of the MMA thread and a parity wait aliased (deadlock caught by the kernel's watchdog in
    round 2); the single producer group cannot.  Results against the FP64 kernel."""
    w = synth.make_workload(3, n_genes=900, n_cells=72, seed=13)
    counts = np.array(w.counts, copy=True)
    rng = np.random.default_rng(7)
    keep = rng.integers(0, 72, size=900)          # number of non-zero cells per gene: 0 .. 71
    for g in range(900):
        zero = rng.permutation(72)[keep[g]:]
        counts[g, zero] = 0
    res = {}
    for kernel in (3, 2):
        keep_opt = ctx.set_options(item_order=item_order, contract_kernel=kernel)
        try:
            res[kernel] = api.scde_posteriors(w.models, counts, w.prior, n_randomizations=60, context=ctx).to_numpy()
        finally:
            ctx.restore_options(keep_opt)
    assert _err(res[3], res[2]) < 1e-6


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_shapes_tcgen05_against_fp64_kernel(ctx, seed):
    """Randomised shapes and sparsity patterns (cells 1 .. 200, genes 1 .. 400, 1 .. 137 randomizations, per-gene density
    anywhere between all-zero and dense, a few huge counts) through the tcgen05 kernel against the FP64 DMMA kernel on
    the same table: every combination of ring stages per item, empty lists, one- and two-pass runs."""
    rng = np.random.default_rng(1000 + seed)
    n_cells = int(rng.integers(1, 201))
    n_genes = int(rng.integers(1, 401))
    n_boot = int(rng.choice([1, 7, 100, 104, 105, 137]))
    w = synth.make_workload(3, n_genes=n_genes, n_cells=max(n_cells, 2), seed=50 + seed)
    models = w.models.iloc[:n_cells]
    counts = np.array(w.counts[:, :n_cells], copy=True)
    dens = rng.uniform(0, 1, size=n_genes)
    counts[rng.uniform(size=counts.shape) > dens[:, None]] = 0
    big = rng.uniform(size=counts.shape) < 0.01
    counts[big] = rng.integers(1, 200000, size=int(big.sum()))
    res = {}
    for kernel in (3, 2):
        keep = ctx.set_options(contract_kernel=kernel)
        try:
            res[kernel] = api.scde_posteriors(models, counts, w.prior, n_randomizations=n_boot, context=ctx).to_numpy()
        finally:
            ctx.restore_options(keep)
    assert _err(res[3], res[2]) < 1e-6, (n_cells, n_genes, n_boot)
