"""Worker of the multi-GPU parity test (launched by torchrun, one rank per GPU): the partitioned
CUDA path with NCCL ghost exchange against the oracle run with the same number of virtual ranks."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from mf_data_locality_b200 import capi, host
from oracle import bp4_oracle as O
from oracle.c_oracle import COracle


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    L = capi.lib()
    for (p, s, plugin) in [(3, 6, "merged"), (4, 7, "merged"), (3, 7, "plain"), (2, 9, "merged")]:
        rds = O.build_problem(p, s, n_ranks=world)
        rd = rds[rank]
        idt = torch.tensor(list(capi.unique_id() if rank == 0 else bytes(128)), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, src=0)
        prob = host.Problem(p, s, plugin=plugin, device=lr, n_ranks=world, rank=rank,
                            nccl_id=bytes(idt.cpu().tolist()))
        assert np.array_equal(prob.entity_index(), rd.entity_index)
        assert np.array_equal(prob.node_of_local(), rd.node_of_local.astype(np.uint64))
        t = O.make_tables(p)
        cos = [COracle(r) for r in rds]
        op = lambda r, v: cos[r.rank].vmult_cells(v)
        rng = np.random.default_rng(5)
        srcs = [np.concatenate([rng.standard_normal(r.n_owned), np.zeros(r.n_ghost)]) for r in rds]
        want = O.multi_vmult(rds, t, srcs, cell_op=op)[rank][: rd.n_owned]
        got = prob.vmult(srcs[rank][: rd.n_owned])
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert err <= 1e-12, ("vmult", p, s, err)
        diag = O.multi_inverse_diagonal(rds)
        errd = np.linalg.norm(prob.diagonal() - diag[rank]) / np.linalg.norm(diag[rank])
        assert errd <= 1e-12, ("diag", errd)
        if plugin == "merged":
            ctl = O.ReductionControl(100, 1e-15, 1e-8)
            xs = O.multi_cg_merged(rds, t, [r.rhs for r in rds], diag, ctl, cell_op=op)
            x, it = prob.run_cg_solver(rd.rhs)      # vmult() above overwrote the input vector
            assert abs(it - ctl.last_step) <= 1, (it, ctl.last_step)
            errx = np.linalg.norm(x - xs[rank][: rd.n_owned]) / np.linalg.norm(xs[rank][: rd.n_owned])
            assert errx <= (1e-8 if ctl.last_step < 100 else 1e-6), ("x", errx)
        else:
            x, it = prob.run_cg_solver(rd.rhs)
            assert 0 < it <= 100
        if rank == 0:
            print(f"mgpu ok: Q{p} s={s} {plugin} world={world} vmult {err:.1e} diag {errd:.1e} it {it}", flush=True)
        prob.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
