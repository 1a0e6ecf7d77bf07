import sys, numpy as np
sys.path.insert(0,'/root/repo')
from amaru_jl_b200 import *
from amaru_jl_b200 import lib as L
n=int(sys.argv[1])
def run(mult, multi_foot, clamp_sides=False):
    mx,my,mz=mult
    mesh = Mesh(Block([[0,0,0],[mx,my,mz]], nx=n*mx, ny=n*my, nz=n*mz, cellshape="HEX20", tag="s"))
    model = FEModel(mesh, [("s", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=0.0))], MechContext())
    bcs=[("z==0", NodeBC(ux=0,uy=0,uz=0))]
    if multi_foot:
        for i in range(mx):
            for j in range(my):
                bcs.append((f"z=={mz} and x>={i+0.4} and x<={i+0.6} and y>={j+0.4} and y<={j+0.6}", NodeBC(uz=-0.01)))
    else:
        cx,cy=mx/2,my/2
        bcs.append((f"z=={mz} and x>={cx-0.1} and x<={cx+0.1} and y>={cy-0.1} and y<={cy+0.1}", NodeBC(uz=-0.01)))
    eqid,nu,setup=model.configure_dofs(bcs); U,F=model.get_bc_vals(eqid,setup)
    dm=L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
    dm.assemble_K(); u=0.1*U; f=0.1*F
    it,rr=dm.solve(u,f,1e-10)
    dm.update_state(u); dm.assemble_K(); u2=0.1*U; f2=0.1*F; it2,_=dm.solve(u2,f2,1e-10)
    dm.close()
    return it,it2
for mult in [(1,1,1),(2,1,1),(2,2,1),(2,2,2),(1,1,2)]:
    print(mult, "single footing", run(mult,False), " multi footing", run(mult,True), flush=True)
