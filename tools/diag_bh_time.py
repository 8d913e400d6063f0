import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bench
from scde_b200 import _lib, api
dev = torch.device("cuda", 0)
models, counts, prior, group, batch = bench.workload_host(4, 30000, 10000, dev)
ctx = _lib.Context(0)
mm, lt, sq = api.pack_models(models)
x, y = prior["x"].to_numpy(), prior["y"].to_numpy()
r = api.expression_difference_call(ctx, counts, mm, x, y, group, 100, 1)
z = r["z"]; cz = np.empty_like(z)
L = _lib.lib()
for _ in range(3): L.scde_b200_bh_cz(_lib.p_f64(z), len(z), _lib.p_f64(cz))
t0 = time.perf_counter()
for _ in range(20): L.scde_b200_bh_cz(_lib.p_f64(z), len(z), _lib.p_f64(cz))
print("bh_cz on the run's Z (30000 genes): %.3f ms; distinct |Z|: %d; cores %d" % ((time.perf_counter() - t0) / 20 * 1e3, len(np.unique(np.abs(z))), len(__import__("os").sched_getaffinity(0))))
t0 = time.perf_counter()
for _ in range(5): r = api.expression_difference_call(ctx, counts, mm, x, y, group, 100, 1)
print("one-shot call wall: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
