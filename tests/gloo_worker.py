"""Worker of the CPU multi-process test (gloo, world_size 2): every rank builds its own partition
with the C++ host mirror (tables only, no GPU) and the ghost exchange plan is exercised with real
point-to-point messages: owners -> ghosts (update_ghost_values) and ghosts -> owners (compress add)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from mf_data_locality_b200 import host


def exchange(plan, send_of_peer, recv_len_of_peer):
    reqs, out = [], {}
    for k, peer in enumerate(plan["rank"]):
        peer = int(peer)
        if len(send_of_peer[k]):
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(send_of_peer[k])), peer))
        if recv_len_of_peer[k]:
            out[k] = torch.empty(recv_len_of_peer[k], dtype=torch.float64)
            reqs.append(dist.irecv(out[k], peer))
    for r in reqs:
        r.wait()
    return {k: v.numpy() for k, v in out.items()}


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for (p, s) in [(3, 6), (4, 6), (2, 8)]:
        prob = host.Problem(p, s, device=-1, n_ranks=world, rank=rank)
        plan = prob.plan()
        nl = prob.node_of_local().astype(np.int64)
        n_owned, n_ghost = prob.n_owned, prob.n_ghost
        f = lambda node, c: np.sin(0.001 * node) + 0.25 * c      # a field defined on lattice nodes
        v = np.zeros(n_owned + n_ghost)
        loc = np.arange(n_owned)
        v[:n_owned] = f(nl[loc // 3], loc % 3)
        io, eo, ex = plan["import_offset"].astype(int), plan["export_offset"].astype(int), plan["export_index"]
        npeer = len(plan["rank"])
        # update_ghost_values
        got = exchange(plan, [v[ex[eo[k]:eo[k + 1]]] for k in range(npeer)], [io[k + 1] - io[k] for k in range(npeer)])
        for k, data in got.items():
            v[n_owned + io[k]: n_owned + io[k + 1]] = data
        gl = np.arange(n_owned, n_owned + n_ghost)
        assert np.array_equal(v[gl], f(nl[gl // 3], gl % 3)), "ghost values do not match their lattice nodes"
        # compress(add): every rank contributes 1 per local copy -> owners end with #ranks holding the DoF
        w = np.ones(n_owned + n_ghost)
        got = exchange(plan, [w[n_owned + io[k]: n_owned + io[k + 1]] for k in range(npeer)],
                       [eo[k + 1] - eo[k] for k in range(npeer)])
        for k, data in got.items():
            np.add.at(w, ex[eo[k]:eo[k + 1]], data)
        holders = torch.from_numpy(w[:n_owned].copy())
        total = torch.tensor([float(holders.sum())])
        dist.all_reduce(total)
        # sum over owners of (number of ranks holding the DoF) == total number of local copies
        copies = torch.tensor([float(n_owned + n_ghost)])
        dist.all_reduce(copies)
        assert total.item() == copies.item(), (total.item(), copies.item())
        # global sizes are consistent
        owned = torch.tensor([float(n_owned)])
        dist.all_reduce(owned)
        assert int(owned.item()) == prob.n_dofs
        prob.close()
    if rank == 0:
        print("gloo ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
