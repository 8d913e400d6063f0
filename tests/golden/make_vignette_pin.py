"""Runs the CPU oracle end to end on the vignette variant of es.mef.small (G = 12142, max.quantile = 0.999,
n.randomizations = 100, Seed = 1; ~1 minute) and stores the G x 6 summary as tests/golden/es_mef_vignette_oracle.npz.
tests/test_oracle.py compares (a) a slice recomputed on the spot and (b) the rows printed in the reference's
vignettes/diffexp.md:113-119 against this file."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import helpers  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    cd, ifm, prior, groups = helpers.es_mef_inputs("vignette")
    codes = np.asarray(groups.codes)
    res = O.expression_difference(ifm, cd.to_numpy(), prior.x.to_numpy(), prior.y.to_numpy(),
                                  (np.nonzero(codes == 0)[0], np.nonzero(codes == 1)[0]), nboot=100, seed=1)
    np.savez_compressed(os.path.join(HERE, "es_mef_vignette_oracle.npz"), results=res["results"], idx=res["idx"],
                        genes=np.array(cd.index, dtype=str))


if __name__ == "__main__":
    main()
