"""GPU timing probe (developer tool, not the bench): the merged CG iteration of the C-ABI path
with the vector updates inside the cell loop (fused, default) and streamed (unfused), plus the
plain operator apply, per kernel.  Set-up through the C++ host mirror; no oracle involved.

  python scripts/probe_merged.py P S [iters]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from mf_data_locality_b200 import capi, host

p, s = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
t0 = time.time()
prob = host.Problem(p, s, plugin="merged", device=0)
ctx = capi.Context.from_handle(prob.ctx_handle(), p, prob.n_cells, prob.n_owned, prob.n_ghost)
n = prob.n_owned
fused, n_priv, n_units = ctx.fused_info()
print(f"Q{p} s={s}: setup {time.time()-t0:.1f}s cells={prob.n_cells} dofs={n} private={n_priv} "
      f"({100.0*n_priv/n:.1f}%) units={n_units}", flush=True)
names = {capi.K_VMULT: "cells", capi.K_MERGED: "cells+upd", capi.K_PRE: "pre", capi.K_POST: "post", capi.K_BLAS1: "blas1"}

src, dst = ctx.vector(data=prob.rhs()), ctx.vector()
for _ in range(3):
    ctx.vmult(dst, src)
ctx.synchronize()
ctx.profile_reset()
ctx.profile_enable(True)
t0 = time.time()
for _ in range(iters):
    ctx.vmult(dst, src)
ctx.synchronize()
wall = (time.time() - t0) / iters
ms, cnt = ctx.profile_get(capi.K_VMULT)
print(f"  vmult : wall {wall*1e3:.3f} ms, cell kernel {ms/cnt:.3f} ms -> {n/(ms/cnt)*1e-6:.2f} GDoF/s (kernel), "
      f"{n/wall*1e-9:.2f} GDoF/s (wall)", flush=True)
ctx.profile_enable(False)

prec = ctx.inverse_diagonal()
for mode in ([True, False] if n_priv > 0 else [False]):
    if n_priv > 0:
        ctx.set_fused(mode)
    x, g, d, h = ctx.vector(), ctx.vector(), ctx.vector(), ctx.vector()
    ctx.equ(g, -1.0, src)
    al = be = ao = bo = 0.0
    def step(it):
        global al, be, ao, bo
        S = ctx.vmult_merged(x, g, d, h, prec, al, be, ao if it % 2 == 1 else 0.0, bo)
        ao, bo = al, be
        al = S[6] / S[0]
        be = al * (S[4] + al * S[5]) / S[6]
    for it in range(1, 4):
        step(it)
    ctx.synchronize()
    ctx.profile_reset()
    ctx.profile_enable(True)
    t0 = time.time()
    for it in range(4, 4 + iters):
        step(it)
    ctx.synchronize()
    wall = (time.time() - t0) / iters
    parts = []
    for kid, nm in names.items():
        ms, cnt = ctx.profile_get(kid)
        if cnt:
            parts.append(f"{nm} {ms/iters:.3f}")
    ctx.profile_enable(False)
    byts = (58.0 + 2.0 / 3.0) * n + 300.0 * prob.n_cells
    print(f"  merged iteration ({'fused' if mode else 'unfused'}): wall {wall*1e3:.3f} ms -> {n/wall*1e-9:.2f} GDoF/s, "
          f"{byts/wall*1e-9:.0f} GB/s algorithmic | ms/it: " + ", ".join(parts), flush=True)
    for v in (x, g, d, h):
        v.free()
prob.close()
