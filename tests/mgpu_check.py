"""Multi-GPU parity check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py

Every rank builds the same model, takes its partition (amaru_jl_b200/partition.py) and runs assemble_K -> solve ->
update_state on its partitioned handle; rank 0 also runs the single-GPU handle and compares U, reactions, dFin and the
IP state of the elements it owns.  Exit code 0 = parity.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from amaru_jl_b200 import Block, FEModel, MechContext, MechSolid, Mesh, NodeBC, VonMises  # noqa: E402
from amaru_jl_b200 import lib as L  # noqa: E402
from amaru_jl_b200.partition import local_flat, partition_mesh  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shape = sys.argv[1] if len(sys.argv) > 1 else "HEX20"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    mesh = Mesh(Block([[0, 0, 0], [1, 1, 2]], nx=n, ny=n, nz=2 * n, cellshape=shape, tag="s"))
    model = FEModel(mesh, [("s", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=1e6))], MechContext())
    bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==2 and x>=0.3 and x<=0.7", NodeBC(uz=-0.006))]
    eqid, nu, setup = model.configure_dofs(bcs)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    ndofs = int(eqid.size)
    flat = model.flatten()
    view = partition_mesh(model.coords, model.conn, world, rank)
    lf, eql = local_flat(flat, eqid, view)
    uid = [L.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    dm = L.DeviceModel(lf, eql, ndofs, nu, device=local, view=view, nccl_uid=uid[0])
    if os.environ.get("AMARU_P2P", "0") == "1":

        def gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        assert dm.p2p_connect(gather), "peer-memory path could not be enabled"
    ok = True
    if rank == 0:
        print(f"operator kernel: {dm.spmv_kernel}; AMARU_P2P={os.environ.get('AMARU_P2P', '0')} "
              f"AMARU_P2P_FUSED={os.environ.get('AMARU_P2P_FUSED', '1')}", flush=True)
    ref = L.DeviceModel(flat, eqid, ndofs, nu, device=local) if rank == 0 else None
    for it in range(2):                                    # second pass: tangent on the plastic trial state
        dm.assemble_K()
        U, F = Uex.copy(), Fex.copy()
        iters, rr = dm.solve(U, F, cg_rtol=1e-12)
        dm.state_restore() if it else None
        dF = dm.update_state(U)
        st = dm.get_state()
        if rank == 0:
            ref.assemble_K()
            U0, F0 = Uex.copy(), Fex.copy()
            it0, rr0 = ref.solve(U0, F0, cg_rtol=1e-12)
            ref.state_restore() if it else None
            dF0 = ref.update_state(U0)
            s0 = ref.get_state()
            nip = model.nip
            own = view.elem_gid[view.elem_owned]
            loc = np.nonzero(view.elem_owned)[0]
            gi = (own[:, None] * nip + np.arange(nip)).reshape(-1)
            li = (loc[:, None] * nip + np.arange(nip)).reshape(-1)

            def rel(a, b):
                return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
            eu, ef, ed, es = rel(U, U0), rel(F[nu:], F0[nu:]), rel(dF, dF0), rel(st["sigma"][li], s0["sigma"][gi])
            print(f"[{world} ranks, {shape} {n}] it{it}: cg {iters} vs {it0} iters | rel diff U {eu:.1e} react {ef:.1e} "
                  f"dFin {ed:.1e} sigma(owned) {es:.1e} | plastic IPs {(s0['dlam'] > 0).sum()}", flush=True)
            ok = ok and eu < 1e-8 and ef < 1e-7 and ed < 1e-7 and es < 1e-7 and abs(iters - it0) <= max(3, it0 // 20)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dm.close()
    if ref is not None:
        ref.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
