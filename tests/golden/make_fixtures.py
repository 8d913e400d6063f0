"""Regenerates tests/golden/*.npz from the reference's bundled data files (run in the authoring container only;
/root/reference does not exist on the GPU box).

  es_mef_small.npz  counts (14897 x 40 int32), gene and cell names      <- data/es.mef.small.rda
  o_ifm.npz         40 x 6 error-model coefficients, row names, groups   <- data/o.ifm.rda
  knn.npz           64 x 12 error-model coefficients (local theta form)  <- data/knn.rda

The binary fixtures are decoded with scde_b200.rdata (stdlib XDR reader); no reference source code is involved.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from scde_b200.rdata import as_data_frame, read_rda  # noqa: E402

REF = "/root/reference/data"


def main():
    cd = as_data_frame(read_rda(os.path.join(REF, "es.mef.small.rda"))["es.mef.small"])
    np.savez_compressed(os.path.join(HERE, "es_mef_small.npz"), counts=cd.to_numpy().astype(np.int32),
                        genes=np.array(cd.index, dtype=str), cells=np.array(cd.columns, dtype=str))
    ifm = as_data_frame(read_rda(os.path.join(REF, "o.ifm.rda"))["o.ifm"])
    np.savez_compressed(os.path.join(HERE, "o_ifm.npz"), values=ifm.to_numpy().astype(np.float64),
                        columns=np.array(ifm.columns, dtype=str), cells=np.array(ifm.index, dtype=str),
                        groups=np.array(list(ifm.attrs["groups"]), dtype=str))
    knn = as_data_frame(read_rda(os.path.join(REF, "knn.rda"))["knn"])
    np.savez_compressed(os.path.join(HERE, "knn.npz"), values=knn.to_numpy().astype(np.float64),
                        columns=np.array(knn.columns, dtype=str), cells=np.array(knn.index, dtype=str))
    for f in ("es_mef_small.npz", "o_ifm.npz", "knn.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
