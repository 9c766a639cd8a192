"""GPU timing probe (developer tool, not the bench): times vmult and the merged CG
iteration of the C-ABI path for one (degree, s).  Uses the oracle only to build inputs."""
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import bp4_oracle as O
from mf_data_locality_b200 import capi

p, s = int(sys.argv[1]), int(sys.argv[2])
variant = int(sys.argv[3]) if len(sys.argv) > 3 else None  # None: the library's per-degree default
t0 = time.time()
rd = O.build_problem(p, s)[0]
print(f"setup {time.time()-t0:.1f}s cells={rd.n_cells} dofs={rd.n_owned}", flush=True)
ctx = capi.Context(p, rd.entity_index, rd.vertices, rd.n_owned, 0, rd.constrained)
if variant is not None:
    ctx.set_merged_variant(variant)
n = rd.n_owned
src, dst = ctx.vector(data=rd.rhs), ctx.vector()
for _ in range(3):
    ctx.vmult(dst, src)
ctx.synchronize()
ctx.profile_reset()
ctx.profile_enable(True)
t0 = time.time()
for _ in range(20):
    ctx.vmult(dst, src)
ctx.synchronize()
wall = (time.time() - t0) / 20
ms, cnt = ctx.profile_get(capi.K_VMULT)
print(f"vmult: wall {wall*1e3:.3f} ms, kernel {ms/cnt:.3f} ms -> {n/(ms/cnt)*1e-6:.2f} GDoF/s (kernel), {n/wall*1e-9:.2f} GDoF/s (wall)")
prec = ctx.inverse_diagonal()
x, g, d, h = ctx.vector(), ctx.vector(), ctx.vector(), ctx.vector()
ctx.equ(g, -1.0, src)
al = be = ao = bo = 0.0
ctx.profile_reset()
t0 = time.time()
its = 20
for it in range(1, its + 1):
    S = ctx.vmult_merged(x, g, d, h, prec, al, be, ao if it % 2 == 1 else 0.0, bo)
    ao, bo = al, be
    al = S[6] / S[0]
    be = al * (S[4] + al * S[5]) / S[6]
ctx.synchronize()
wall = (time.time() - t0) / its
parts = {k: ctx.profile_get(i) for k, i in (("vmult", capi.K_VMULT), ("merged", capi.K_MERGED), ("pre", capi.K_PRE), ("post", capi.K_POST))}
print("merged iteration: wall %.3f ms -> %.2f GDoF/s; kernels: %s" % (wall * 1e3, n / wall * 1e-9, {k: round(v[0] / max(v[1], 1), 3) for k, v in parts.items()}))
print("residual-ish", S)
