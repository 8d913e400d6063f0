"""The oracle's restatement (oracle/scde_oracle.c) against THE REFERENCE ITSELF: /root/reference/src/jpmatLogBoot.cpp and
src/matSlideMult.cpp compiled unmodified against the header shim oracle/shim/RcppArmadillo.h (oracle/_ref/libscde_ref.so,
`make -C oracle ref`).  Bit-exact on every entry point: the loop nests, the summation orders (Armadillo's two-accumulator
`accumulate`, column-wise `sum(.., 1)`), the libc `srand/rand` rejection loop, the snap rule, the clamp, the modes and
individual posteriors, the batch sampler, the ensemble and no-bootstrap forms.  What the shim cannot pin is R nmath's
`dnbinom` / `dpois` (R is absent): both sides use the oracle's restatement, which tests/test_oracle.py pins with mpmath.

Where /root/reference is absent (the GPU box) the prebuilt library is used; if neither exists the module is skipped and
the committed fixtures of tests/golden/ref_fixtures.npz (written by tests/golden/make_ref_fixtures.py from this library)
still pin the oracle in tests/test_oracle.py::test_oracle_matches_reference_fixtures.
"""
import numpy as np
import pytest

import helpers
from oracle import oracle as O
from oracle import ref as R
from scde_b200 import synth

pytestmark = pytest.mark.skipif(not R.available(), reason="neither the reference sources nor a prebuilt oracle/_ref")


def _prep(w):
    mm, lt, sq = O.pack_models(w.models)
    mag = O.marginals_from_prior_x(w.prior["x"].to_numpy())
    flat, off, uci = O.unique_counts(w.counts)
    return mm, lt, sq, mag, flat, off, uci


@pytest.mark.parametrize("G,Cn,B,seed", [(97, 23, 100, 1), (5, 1, 7, 3), (33, 40, 150, 1), (1, 9, 100, 12143)])
@pytest.mark.parametrize("returnpost", [0, 1, 2, 3])
def test_log_boot_posterior_bit_exact(G, Cn, B, seed, returnpost):
    w = synth.make_workload(3, n_genes=G, n_cells=Cn, seed=5)
    mm, lt, sq, mag, flat, off, uci = _prep(w)
    a = O.log_boot_posterior(mm, flat, off, uci, mag, B, seed=seed, returnpost=returnpost)
    b = R.log_boot_posterior(mm, flat, off, uci, mag, B, seed=seed, returnpost=returnpost)
    assert set(a) == set(b)
    assert np.array_equal(a["jp"], b["jp"])
    if "modes" in a:
        assert np.array_equal(a["modes"], b["modes"])
    if "post" in a:
        for x, y in zip(a["post"], b["post"]):
            assert np.array_equal(x, y)


def test_explicit_draws_equal_the_reference_rand_stream():
    """the oracle fed with the draws of orc_boot_indices (what the GPU path is fed with) == the reference drawing inside"""
    w = synth.make_workload(3, n_genes=40, n_cells=31, seed=2)
    mm, lt, sq, mag, flat, off, uci = _prep(w)
    for seed in (1, 7, 40):
        bi = O.boot_indices(seed, 31, 60)
        a = O.log_boot_posterior(mm, flat, off, uci, mag, 60, seed=999, boot_idx=bi)
        b = R.log_boot_posterior(mm, flat, off, uci, mag, 60, seed=seed)
        assert np.array_equal(a["jp"], b["jp"])


def test_no_bootstrap_and_ensemble_bit_exact():
    w = synth.make_workload(3, n_genes=60, n_cells=17, seed=8)
    mm, lt, sq, mag, flat, off, uci = _prep(w)
    for kw in (dict(nboot=0), dict(nboot=10, ensemble=1), dict(nboot=0, ensemble=1)):
        a = O.log_boot_posterior(mm, flat, off, uci, mag, seed=1, **kw)
        b = R.log_boot_posterior(mm, flat, off, uci, mag, seed=1, **kw)
        assert np.array_equal(a["jp"], b["jp"]), kw


@pytest.mark.parametrize("returnpost", [0, 1, 2])
def test_log_boot_batch_posterior_bit_exact(returnpost):
    w = synth.make_workload(5, n_genes=80, n_cells=36, seed=3, batch=True)
    mm, lt, sq, mag, flat, off, uci = _prep(w)
    bc = np.asarray(w.batch.codes)
    pools = [np.nonzero(bc == l)[0].astype(np.int32) for l in range(2)]
    codes = np.asarray(w.groups.codes)
    for lev in (0, 1):
        comp = np.bincount(bc[codes == lev], minlength=2).astype(np.int32)
        a = O.log_boot_batch_posterior(mm, flat, off, uci, mag, pools, comp, 100, seed=1, returnpost=returnpost)
        b = R.log_boot_batch_posterior(mm, flat, off, uci, mag, pools, comp, 100, seed=1, returnpost=returnpost)
        assert np.array_equal(a["jp"], b["jp"])
        if "modes" in a:
            assert np.array_equal(a["modes"], b["modes"])
        if "post" in a:
            for x, y in zip(a["post"], b["post"]):
                assert np.array_equal(x, y)
    # a level with zero composition is skipped by both (src/jpmatLogBoot.cpp:472)
    comp = np.array([0, 9], dtype=np.int32)
    a = O.log_boot_batch_posterior(mm, flat, off, uci, mag, pools, comp, 30, seed=5)
    b = R.log_boot_batch_posterior(mm, flat, off, uci, mag, pools, comp, 30, seed=5)
    assert np.array_equal(a["jp"], b["jp"])
    bi = O.batch_boot_indices(5, pools, comp, 30)
    c = O.log_boot_batch_posterior(mm, flat, off, uci, mag, pools, comp, 30, seed=77, boot_idx=bi)
    assert np.array_equal(c["jp"], b["jp"])


def test_local_theta_and_square_logit_models_bit_exact():
    """knn.error.models output (data/knn.rda, 12 columns: corr.ltheta.*, conc.a2) -- src/jpmatLogBoot.cpp:136-138,148-176"""
    knn = helpers.knn_models().iloc[:12]
    rng = np.random.default_rng(3)
    counts = rng.negative_binomial(0.8, 0.02, size=(70, len(knn))).astype(np.int32)
    counts[rng.uniform(size=counts.shape) < 0.4] = 0
    mm, lt, sq = O.pack_models(knn)
    assert lt == 1 and sq == 1
    mag = O.marginals_from_prior_x(np.linspace(0, 4.8, 401))
    flat, off, uci = O.unique_counts(counts)
    a = O.log_boot_posterior(mm, flat, off, uci, mag, 50, seed=1, returnpost=3, localtheta=lt, sqlogit=sq)
    b = R.log_boot_posterior(mm, flat, off, uci, mag, 50, seed=1, returnpost=3, localtheta=lt, sqlogit=sq)
    assert np.array_equal(a["jp"], b["jp"]) and np.array_equal(a["modes"], b["modes"])
    for x, y in zip(a["post"], b["post"]):
        assert np.array_equal(x, y)


def test_es_mef_small_slice_bit_exact():
    """config 1's data: the bundled es.mef.small counts + o.ifm models, a 300-gene slice, both groups"""
    cd, ifm, prior, groups = helpers.es_mef_inputs("tests")
    sub = cd.iloc[1000:1300]
    mm, lt, sq = O.pack_models(ifm)
    mag = O.marginals_from_prior_x(prior["x"].to_numpy())
    codes = np.asarray(groups.codes)
    for lev in (0, 1):
        ii = np.nonzero(codes == lev)[0]
        flat, off, uci = O.unique_counts(np.asfortranarray(sub.to_numpy()[:, ii]))
        a = O.log_boot_posterior(np.asfortranarray(mm[ii]), flat, off, uci, mag, 100, seed=1, returnpost=1)
        b = R.log_boot_posterior(np.asfortranarray(mm[ii]), flat, off, uci, mag, 100, seed=1, returnpost=1)
        assert np.array_equal(a["jp"], b["jp"]) and np.array_equal(a["modes"], b["modes"])


def test_legacy_jpmat_forms_bit_exact():
    rng = np.random.default_rng(11)
    matl = [np.asfortranarray(-rng.gamma(2.0, 3.0, size=(13, 29))) for _ in range(9)]
    assert np.array_equal(O.jpmat_log_boot(matl, 40, seed=3), R.jpmat_log_boot(matl, 40, seed=3))
    matll = [matl[:4], matl[4:]]
    comp = np.array([3, 6], dtype=np.int32)
    assert np.array_equal(O.jpmat_log_batch_boot(matll, comp, 25, seed=2), R.jpmat_log_batch_boot(matll, comp, 25, seed=2))
    comp0 = np.array([0, 5], dtype=np.int32)
    assert np.array_equal(O.jpmat_log_batch_boot(matll, comp0, 25, seed=2), R.jpmat_log_batch_boot(matll, comp0, 25, seed=2))


@pytest.mark.parametrize("n", [1, 2, 33, 401])
def test_mat_slide_mult_bit_exact(n):
    rng = np.random.default_rng(n)
    m1, m2 = rng.random((11, n)), rng.random((11, n))
    assert np.array_equal(O.mat_slide_mult(m1, m2), R.mat_slide_mult(m1, m2))
