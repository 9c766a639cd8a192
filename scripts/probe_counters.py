"""Developer tool for the ncu counter pass (profiles/kernel_counters.json): for every degree one
plain operator apply and one merged iteration with the in-loop vector updates, so that ncu sees
exactly one cell_kernel<P,CPB,false> and one cell_kernel<P,CPB,true> launch per degree (after a
warm-up launch of each).  Prints the DoF counts the post-processing needs.

  ncu --metrics ... -k regex:cell_kernel --csv --log-file gpurun_out/counters.csv \
      python scripts/probe_counters.py [s]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mf_data_locality_b200 import capi, host

s = int(sys.argv[1]) if len(sys.argv) > 1 else 15
info = {}
for p in range(2, 9):
    prob = host.Problem(p, s, plugin="merged", device=0)
    ctx = capi.Context.from_handle(prob.ctx_handle(), p, prob.n_cells, prob.n_owned, prob.n_ghost)
    src, dst = ctx.vector(data=prob.rhs()), ctx.vector()
    prec = ctx.inverse_diagonal()
    x, g, d, h = ctx.vector(), ctx.vector(), ctx.vector(), ctx.vector()
    ctx.equ(g, -1.0, src)
    for rep in range(2):                 # launch 0 of each kind warms up, launch 1 is the one to read
        ctx.vmult(dst, src)
    if ctx.fused_info()[1] > 0:
        ctx.set_fused(True)
    S = ctx.vmult_merged(x, g, d, h, prec, 0.0, 0.0, 0.0, 0.0)
    al = S[6] / S[0]
    be = al * (S[4] + al * S[5]) / S[6]
    S = ctx.vmult_merged(x, g, d, h, prec, al, be, 0.0, 0.0)
    ctx.synchronize()
    info[p] = {"s": s, "n_dofs": prob.n_owned, "n_cells": prob.n_cells, "n_private": ctx.fused_info()[1]}
    prob.close()
print("COUNTER_INFO " + json.dumps(info), flush=True)
