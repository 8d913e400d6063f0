"""diagnostic: scde.posteriors over all 40 es.mef.small cells with different item orders / gene counts"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers
from scde_b200 import _lib, api

cd, ifm, prior, groups = helpers.es_mef_inputs("tests")
for order in (0, 1):
    for n in (2000, 8000, 13788):
        for ncell in (20, 40):
            ctx = _lib.Context(0)
            ctx.set_options(item_order=order)
            try:
                r = api.scde_posteriors(ifm.iloc[:ncell], cd.iloc[:n, :ncell], prior, n_randomizations=100, context=ctx)
                print("order", order, "genes", n, "cells", ncell, "ok", float(r.to_numpy().sum()))
            except Exception as e:
                print("order", order, "genes", n, "cells", ncell, "FAILED", str(e)[:120])
            ctx.close()
