"""GPU tests of the full drop-in path: C++ host mirror (Renumber -> LaplaceOperator::initialize
-> run_cg_solver plugin) over the C ABI, against the CPU oracle."""
import numpy as np
import pytest

from oracle import bp4_oracle as O

from helpers import rel_l2, single

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p,s", [(3, 6), (4, 7), (6, 5), (2, 8)])
def test_host_vmult_and_tables(p, s, bp4_lib, c_oracle_lib):
    from mf_data_locality_b200 import host
    rd, co = single(p, s)
    prob = host.Problem(p, s, plugin="merged", device=0)
    assert np.array_equal(prob.entity_index(), rd.entity_index)
    assert np.array_equal(prob.constrained(), rd.constrained)
    assert np.array_equal(prob.rhs(), rd.rhs)
    v = np.random.default_rng(p).standard_normal(rd.n_owned)
    assert rel_l2(prob.vmult(v), co.vmult(v)) <= 1e-12
    want = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
    assert rel_l2(prob.diagonal(), want) <= 1e-12
    prob.close()


@pytest.mark.parametrize("plugin", ["plain", "merged"])
@pytest.mark.parametrize("p,s,rel_tol", [(3, 6, 1e-8), (4, 6, 1e-8), (4, 6, 1e-10), (3, 9, 1e-8)])
def test_run_cg_solver_plugin(plugin, p, s, rel_tol, bp4_lib, c_oracle_lib):
    """iteration count within +-1 and solution within 1e-8 of the oracle (north_star)"""
    from mf_data_locality_b200 import host
    rd, co = single(p, s)
    prob = host.Problem(p, s, plugin=plugin, device=0)
    prob.set_solver(100, 1e-15, rel_tol)
    prec = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
    x, it = prob.run_cg_solver()          # default RHS i % 8
    xo, ito, hist = co.cg(rd.rhs, prec, merged=(plugin == "merged"), reduce=rel_tol)
    assert abs(it - ito) <= 1
    assert rel_l2(x, xo) <= (1e-8 if ito < 100 else 1e-6)
    # solving twice gives the same answer (x0 = 0 each time, benchmark.h:191)
    x2, it2 = prob.run_cg_solver()
    assert it2 == it and rel_l2(x2, x) <= 1e-9
    prob.set_solver()
    prob.close()


def test_cli_verbose_output(bp4_lib):
    """`bench <degree> <s> 0`: the non-compact prints of run_templated (benchmark.h:149-154,
    178-182) -- the preconditioner's diagonal norm (checked against the oracle) and the setup time;
    `bench <degree> <s>`: the one-line table row (benchmark.h:217-225)."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "mf_data_locality_b200", "benchmark_precond_merged", "bench")
    out = subprocess.run([exe, "3", "9", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    m = re.search(r"Norm of diagonal for preconditioner: ([0-9.eE+-]+)", out.stdout)
    assert m and "Setup time:" in out.stdout
    rd, _ = single(3, 9)
    want = np.linalg.norm(O.finish_inverse_diagonal(O.inverse_diagonal(rd)))
    assert abs(float(m.group(1)) - want) <= 1e-5 * want  # printed with six significant digits
    row = subprocess.run([exe, "3", "9"], capture_output=True, text=True, timeout=300)
    assert row.returncode == 0, row.stderr
    cols = [c.strip() for c in row.stdout.strip().splitlines()[-1].split("|")]
    assert cols[0] == "3" and cols[1] == "5" and int(cols[2]) == rd.n_cells and int(cols[3]) == rd.n_owned
    assert 1 <= int(cols[6]) <= 100
