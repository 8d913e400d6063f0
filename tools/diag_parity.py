"""Diagnostic (GPU box): tiled vs generic contraction kernel vs CPU oracle at many cells per group."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from scde_b200 import _lib, api, synth

G, C, NCPU = int(sys.argv[1]) if len(sys.argv) > 1 else 128, int(sys.argv[2]) if len(sys.argv) > 2 else 10000, 32
w = synth.make_workload(4, n_genes=G, n_cells=C)
ctx = _lib.default_context(0)
out = {}
for kern in (2, 1):
    ctx.set_contract_kernel(kern)
    out[kern] = api.scde_expression_difference(w.models, w.counts, w.prior, groups=w.groups, n_randomizations=100,
                                               return_posteriors=True, context=ctx)
ctx.set_contract_kernel(0)
def lerr(a, b):
    big = (a > 1e-290) & (b > 1e-290)
    return float(np.max(np.abs(np.log(a[big]) - np.log(b[big])) / np.maximum(1, np.abs(np.log(b[big])))))
for lev in ("g1", "g2"):
    a, b = out[2]["joint.posteriors"][lev].to_numpy(), out[1]["joint.posteriors"][lev].to_numpy()
    print("tiled vs generic jp", lev, lerr(a, b), "max abs", float(np.abs(a - b).max()))
print("tiled vs generic dZ", float(np.abs(out[2]["results"]["Z"] - out[1]["results"]["Z"]).max()))
codes = np.asarray(w.groups.codes)
t = time.time()
want = O.expression_difference(w.models, w.counts[:NCPU], w.prior["x"].to_numpy(), w.prior["y"].to_numpy(),
                               (np.nonzero(codes == 0)[0], np.nonzero(codes == 1)[0]), nboot=100, seed=1)
print("oracle s", time.time() - t)
for i, lev in enumerate(("g1", "g2")):
    a = out[2]["joint.posteriors"][lev].to_numpy()[:NCPU]
    print("tiled vs oracle jp", lev, lerr(a, want["joint.posteriors"][i]), "max abs", float(np.abs(a - want["joint.posteriors"][i]).max()))
dz = np.abs(out[2]["results"]["Z"].to_numpy()[:NCPU] - want["results"][:, 4])
j = int(np.argmax(dz))
print("tiled vs oracle dZ", float(dz.max()), "gene", j, out[2]["results"]["Z"].iloc[j], want["results"][j, 4])
print("dp err", lerr(out[2]["difference.posterior"].to_numpy()[:NCPU], want["difference.posterior"]))
# per-row table check on a few cells
mm, lt, sq = O.pack_models(w.models)
mag = O.marginals_from_prior_x(w.prior["x"].to_numpy())
worst = 0
for cell in (0, 17, 5000, 9999):
    uc = np.unique(w.counts[:, cell]).astype(np.int32)
    want_t, _ = O.cell_table(mm[cell], uc, mag, ncells_for_clamp=C)
    got = np.empty((len(uc), len(mag)))
    _lib.check(_lib.lib().scde_b200_cell_table(ctx.handle, _lib.p_f64(_lib.f64(mm[cell])), _lib.p_i32(uc), len(uc),
                                               _lib.p_f64(_lib.f64(mag)), len(mag), 0, 0, C, _lib.p_f64(got), None))
    wt = want_t.T
    fin = wt > -1e300
    d = np.abs(got[fin] - wt[fin])
    k = np.unravel_index(np.argmax(np.where(fin, np.abs(got - wt), 0)), wt.shape)
    print("cell", cell, "rows", len(uc), "max abs row diff", float(d.max()), "at count", int(uc[k[0]]), "k", int(k[1]), wt[k], got[k])
