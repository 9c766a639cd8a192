"""Regenerates tests/golden/bp4_golden.npz.

The reference ships NO golden vectors and cannot be run here, so these are
SELF-GENERATED regression fixtures: outputs of the numpy oracle (which tests/test_oracle.py pins
against dense brute-force assembly) on tiny seeded problems.  They freeze the oracle's numbering
and arithmetic so that accidental changes show up, and give the GPU tests a committed target.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bp4_oracle as O  # noqa: E402

CASES = [(2, 3, 1), (3, 3, 1), (4, 2, 1), (3, 4, 2), (5, 1, 1)]


def main():
    out = {}
    for p, s, n_ranks in CASES:
        rds = O.build_problem(p, s, n_ranks=n_ranks)
        t = O.make_tables(p)
        for rd in rds:
            key = f"p{p}_s{s}_r{n_ranks}_{rd.rank}"
            out[key + "_node_of_local"] = rd.node_of_local.astype(np.int64)
            out[key + "_entity_index"] = rd.entity_index
            out[key + "_constrained"] = rd.constrained
        if n_ranks == 1:
            rd = rds[0]
            rng = np.random.default_rng(1000 * p + s)
            v = rng.standard_normal(rd.n_owned)
            out[f"p{p}_s{s}_src"] = v
            out[f"p{p}_s{s}_vmult"] = O.vmult(rd, t, v)
            out[f"p{p}_s{s}_invdiag"] = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
            ctl = O.ReductionControl(100, 1e-15, 1e-8)
            x = O.solver_cg_merged(lambda w: O.vmult_cells(rd, t, w), np.zeros(rd.n_owned), rd.rhs,
                                   out[f"p{p}_s{s}_invdiag"], ctl)
            out[f"p{p}_s{s}_cg_x"] = x
            out[f"p{p}_s{s}_cg_its"] = np.array([ctl.last_step])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bp4_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
