"""Host logic of the multi-GPU path (no GPU): partition views, halo lists, and a real world_size-2 exchange over gloo."""
import os
import sys

import numpy as np
import pytest

from amaru_jl_b200 import Block, FEModel, LinearElastic, MechContext, MechSolid, Mesh, NodeBC
from amaru_jl_b200.partition import local_flat, partition_mesh, rcb_partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def model(shape="HEX20", n=4):
    mesh = Mesh(Block([[0, 0, 0], [1, 1, 2]], nx=n, ny=n, nz=2 * n, cellshape=shape, tag="s"))
    return FEModel(mesh, [("s", MechSolid, LinearElastic, dict(E=1.0, nu=0.3))], MechContext())


@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
@pytest.mark.parametrize("shape", ["HEX8", "HEX20", "TET10"])
def test_views_are_consistent(shape, nranks):
    m = model(shape, 3)
    views = [partition_mesh(m.coords, m.conn, nranks, r) for r in range(nranks)]
    part = rcb_partition(m.coords[m.conn].mean(axis=1), nranks)
    cnt = np.bincount(part, minlength=nranks)
    assert cnt.max() - cnt.min() <= nranks                      # balanced
    owned_all = np.concatenate([v.node_gid[:v.nowned] for v in views])
    assert np.array_equal(np.sort(owned_all), np.arange(m.nnodes))       # every node owned exactly once
    eown_all = np.concatenate([v.elem_gid[v.elem_owned] for v in views])
    assert np.array_equal(np.sort(eown_all), np.arange(m.nelems))        # every element owned exactly once
    # node -> elements adjacency, globally
    adj = [[] for _ in range(m.nnodes)]
    for e, c in enumerate(m.conn):
        for n in c:
            adj[n].append(e)
    g = np.arange(m.nnodes, dtype=np.float64) * 1.5 + 7.0                # a global nodal field
    for v in views:
        local_e = set(v.elem_gid.tolist())
        for n in v.node_gid[:v.nowned]:
            assert set(adj[n]) <= local_e                                # owned rows can be assembled locally
        assert (v.conn >= 0).all() and v.conn.max() < v.node_gid.size
        assert np.array_equal(v.node_gid[v.conn], m.conn[v.elem_gid])
    # simulated halo exchange: ghosts get exactly the owners' values
    for p, v in enumerate(views):
        x = np.full(v.node_gid.size, np.nan)
        x[:v.nowned] = g[v.node_gid[:v.nowned]]
        for i, q in enumerate(v.neigh):
            w = views[q]
            j = list(w.neigh).index(p)
            sent = g[w.node_gid[w.send_nodes[w.send_ptr[j]:w.send_ptr[j + 1]]]]
            assert sent.size == v.recv_count[i]
            x[v.recv_start[i]:v.recv_start[i] + v.recv_count[i]] = sent
        assert np.array_equal(x, g[v.node_gid])


def test_local_flat_restricts_arrays():
    m = model("HEX8", 3)
    eqid, nu, _ = m.configure_dofs([("z==0", NodeBC(ux=0, uy=0, uz=0))])
    v = partition_mesh(m.coords, m.conn, 2, 1)
    lf, eql = local_flat(m.flatten(), eqid, v)
    assert lf["coords"].shape[0] == v.node_gid.size and lf["conn"].shape == v.conn.shape
    assert np.array_equal(eql, eqid[v.node_gid])
    assert lf["batch_nelem"][0] == v.elem_gid.size


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from amaru_jl_b200 import Block, Mesh
from amaru_jl_b200.partition import partition_mesh
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
mesh = Mesh(Block([[0, 0, 0], [1, 1, 2]], nx=3, ny=3, nz=6, cellshape="HEX20"))
v = partition_mesh(mesh.coords, mesh.conn, world, rank)
g = np.arange(mesh.nnodes, dtype=np.float64) * 0.25 - 3.0
x = torch.full((v.node_gid.size, 3), float("nan"), dtype=torch.float64)
x[:v.nowned] = torch.from_numpy(g[v.node_gid[:v.nowned]])[:, None] * torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
reqs, keep = [], []
for i, q in enumerate(v.neigh):
    sb = x[torch.from_numpy(v.send_nodes[v.send_ptr[i]:v.send_ptr[i + 1]].astype(np.int64))].contiguous()
    keep.append(sb)
    reqs.append(dist.isend(sb, int(q)))
    reqs.append(dist.irecv(x[int(v.recv_start[i]):int(v.recv_start[i] + v.recv_count[i])], int(q)))
for r in reqs:
    r.wait()
ref = torch.from_numpy(g[v.node_gid])[:, None] * torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
ok = torch.equal(x, ref)
# a distributed dot product over owned rows equals the global one (what the PCG all-reduces)
part = torch.tensor([float((x[:v.nowned] ** 2).sum())], dtype=torch.float64)
dist.all_reduce(part)
ok = ok and abs(part.item() - float(((g[:, None] * np.array([1.0, 2.0, 3.0])) ** 2).sum())) < 1e-6 * part.item()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_halo_exchange_gloo_world2(tmp_path):
    import subprocess
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    codes = [p.wait(timeout=180) for p in procs]
    assert codes == [0, 0]


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_peer_memory_halo_addressing(nranks):
    """The peer-memory halo push writes rank p's k-th send entry for neighbour q to q's local node
    peer_recv_start + k: that slot must be exactly the ghost copy of the same global node (amaru_p2p_connect contract)."""
    from amaru_jl_b200.partition import peer_recv_starts
    m = Mesh(Block([[0, 0, 0], [1, 1, 2]], nx=4, ny=4, nz=8, cellshape="HEX20"))
    views = [partition_mesh(m.coords, m.conn, nranks, r) for r in range(nranks)]
    everyone = [dict(neigh=[int(q) for q in v.neigh], recv_start=[int(s) for s in v.recv_start]) for v in views]
    for p, v in enumerate(views):
        starts = peer_recv_starts(p, v.neigh, everyone)
        for i, q in enumerate(v.neigh):
            sent = v.node_gid[v.send_nodes[v.send_ptr[i]:v.send_ptr[i + 1]]]
            w = views[int(q)]
            dst = w.node_gid[starts[i]:starts[i] + sent.size]
            assert starts[i] >= w.nowned and np.array_equal(sent, dst)
