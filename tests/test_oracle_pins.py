"""Pins of the CPU oracle that do NOT depend on the builder's reading of the reference's return mapping:

  * yield consistency: after every plastic update the new state sits on the yield surface, |f(σ, εpa)| ~ 0 — for the
    perfectly plastic cases (H = 0) of von Mises (von-mises.jl:104-156) and Drucker-Prager on the cone AND at the apex
    (drucker-prager.jl:76-149).  (With H > 0 the reference's Δλ = f/(3G + √1.5·H) does not return exactly onto its own
    f = √(3J2) − fy − H·εpa; that quirk is kept and therefore not asserted here.)
  * tangent vs finite differences: from a state ON the yield surface, a small loading increment t·δε must change the
    stress by t·D(σ, Δλ)·δε + O(t²) with D = calcD of the reference (von-mises.jl:112-125, drucker-prager.jl:86-109);
  * the one known answer of the reference's element tests that round 1 did not reproduce: the EdgeBC case of
    test/mech/elem/elastic-hex8.jl (qy = 2 on the edge y==1 && z==1: uz = 3.32088, 3.1998, -4.6002, -4.32047).
"""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import oracle_loads as OL

LE, VM, DP = 1, 2, 3
SR2 = 2.0 ** 0.5


def mandel(s):
    """(xx, yy, zz, yz, xz, xy) tensor components -> Mandel vector (tensors.jl:24-25)."""
    s = np.asarray(s, dtype=np.float64)
    return np.array([s[0], s[1], s[2], SR2 * s[3], SR2 * s[4], SR2 * s[5]])


def f_vm(sig, epa, fy, H):
    return np.sqrt(3.0 * O.J2(sig)) - fy - H * epa


def f_dp(sig, epa, alpha, kappa, H):
    return alpha * sig[:3].sum() + np.sqrt(O.J2(sig)) - kappa - H * epa


def test_J2_is_the_second_deviatoric_invariant():
    rng = np.random.default_rng(0)
    for _ in range(10):
        t = rng.normal(size=6)
        T = np.array([[t[0], t[5], t[4]], [t[5], t[1], t[3]], [t[4], t[3], t[2]]])
        dev = T - np.trace(T) / 3 * np.eye(3)
        assert abs(O.J2(mandel(t)) - 0.5 * (dev * dev).sum()) < 1e-13


@pytest.mark.parametrize("seed", range(6))
def test_von_mises_returns_onto_the_yield_surface(seed):
    rng = np.random.default_rng(seed)
    E, nu, fy = 210e6, 0.3, 240e3
    par = [E, nu, fy, 0.0]
    sig, eps, epa, dlam = np.zeros(6), np.zeros(6), 0.0, 0.0
    nplastic = 0
    for step in range(6):
        deps = mandel(rng.normal(size=6)) * 1.5e-3
        st, sig, eps, epa, dlam, _ = O.update_ip(VM, par, sig, eps, epa, dlam, deps)
        assert st == 0
        if dlam > 0:
            nplastic += 1
            assert abs(f_vm(sig, epa, fy, 0.0)) <= 1e-8 * fy
        else:
            assert f_vm(sig, epa, fy, 0.0) < 1e-8
    assert nplastic >= 3


@pytest.mark.parametrize("seed", range(6))
def test_drucker_prager_cone_return_is_consistent(seed):
    rng = np.random.default_rng(seed)
    E, nu, alpha, kappa = 100.0, 0.25, 0.05, 0.1
    par = [E, nu, alpha, kappa, 0.0]
    sig, eps, epa, dlam = np.zeros(6), np.zeros(6), 0.0, 0.0
    nplastic = 0
    for step in range(6):
        d = rng.normal(size=6)
        d[:3] -= d[:3].mean() + 0.3                     # compressive mean strain keeps the trial state off the apex
        deps = mandel(d) * 2e-3
        st, sig, eps, epa, dlam, _ = O.update_ip(DP, par, sig, eps, epa, dlam, deps)
        assert st == 0
        if dlam > 0:
            nplastic += 1
            assert O.J2(sig) > 0
            assert abs(f_dp(sig, epa, alpha, kappa, 0.0)) <= 1e-10
    assert nplastic >= 3


def test_drucker_prager_apex_return_is_consistent():
    """Hydrostatic tension beyond κ/α: the second plastic step takes the apex branch (the switch uses the PREVIOUS Δγ,
    drucker-prager.jl:130) and lands on the apex: J2 = 0, α·J1 = κ."""
    E, nu, alpha, kappa = 100.0, 0.25, 0.05, 0.1
    par = [E, nu, alpha, kappa, 0.0]
    sig, eps, epa, dlam = np.zeros(6), np.zeros(6), 0.0, 0.0
    seen_apex = False
    for step in range(4):
        deps = np.array([1.0, 1.0, 1.0, 1e-3, 0.0, 0.0]) * 8e-3
        st, sig, eps, epa, dlam, _ = O.update_ip(DP, par, sig, eps, epa, dlam, deps)
        assert st == 0
        if dlam > 0 and O.J2(sig) == 0.0:
            seen_apex = True
            assert abs(alpha * sig[:3].sum() - kappa) <= 1e-12
            assert abs(f_dp(sig, epa, alpha, kappa, 0.0)) <= 1e-12
    assert seen_apex


@pytest.mark.parametrize("kind,par", [(VM, [210e6, 0.3, 240e3, 0.0]), (DP, [100.0, 0.25, 0.05, 0.1, 0.0])])
def test_tangent_matches_finite_differences_of_the_update(kind, par):
    """From a state on the yield surface, σ(t·δε) − σ(0) = t·D·δε + O(t²) for loading directions δε."""
    rng = np.random.default_rng(11)
    scale = 2e-3 if kind == VM else 6e-3
    # drive the point onto the surface with two plastic steps
    sig, eps, epa, dlam = np.zeros(6), np.zeros(6), 0.0, 0.0
    base = np.array([1.0, -0.4, -0.9, 0.3, -0.2, 0.5]) if kind == VM else np.array([-1.0, 0.3, 0.5, 0.3, -0.2, 0.4])
    if kind == VM:
        for _ in range(2):
            st, sig, eps, epa, dlam, _ = O.update_ip(kind, par, sig, eps, epa, dlam, mandel(base) * scale)
            assert st == 0 and dlam > 0
    else:
        # Drucker-Prager chooses cone / apex with the PREVIOUS Δγ (drucker-prager.jl:130, kept): after a large plastic step a
        # small one would be sent to the apex.  Step just past first yield instead (bisection), so that Δγ stays small.
        lo, hi = 0.0, 1.0
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            _, s_, _, _, dl_, _ = O.update_ip(kind, par, np.zeros(6), np.zeros(6), 0.0, 0.0, mandel(base) * scale * mid)
            lo, hi = (lo, mid) if dl_ > 0 else (mid, hi)
        st, sig, eps, epa, dlam, _ = O.update_ip(kind, par, sig, eps, epa, dlam, mandel(base) * scale * hi * (1 + 1e-4))
        assert st == 0 and dlam > 0 and O.J2(sig) > 0
    D, st = O.calcD(kind, par, sig, dlam)
    assert st == 0
    De = O.calcDe(par[0], par[1])
    assert np.abs(D - De).max() > 1e-3 * np.abs(De).max()                 # really the elasto-plastic tangent
    checked = 0
    for _ in range(12):
        de = mandel(base + 0.4 * rng.normal(size=6))
        errs = []
        for t in (1e-6, 5e-7):
            st, s1, _, _, dl1, _ = O.update_ip(kind, par, sig, eps, epa, dlam, de * t * scale)
            assert st == 0
            if dl1 == 0.0:                                                # elastic unloading direction: not this test
                errs = None
                break
            lin = D @ (de * t * scale)
            errs.append(np.abs((s1 - sig) - lin).max() / np.abs(lin).max())
        if errs is None:
            continue
        checked += 1
        assert errs[0] < 5e-4 and errs[1] < 0.7 * errs[0] + 1e-9          # first-order consistent: the error shrinks with t
    assert checked >= 6


def test_edgebc_known_answer_elastic_hex8():
    """reference test/mech/elem/elastic-hex8.jl, load case 2: EdgeBC(qy=2) on y==1 && z==1 (mech_boundary_forces on a
    LIN2 edge in 3D: th = 1, coef = |J|·w, distributed.jl:88-147)."""
    from amaru_jl_b200 import Block, FEModel, LinearElastic, MechContext, MechSolid, Mesh, NodeBC
    mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=1, ny=1, nz=1, cellshape="HEX8", tag="solid"), native=False)
    model = FEModel(mesh, [("solid", MechSolid, LinearElastic, dict(E=1.0, nu=0.3))], MechContext())
    bcs = [("x==0 && y==0 && z==0", NodeBC(ux=0, uy=0)), ("x==1 && y==0 && z==0", NodeBC(uy=0)),
           ("x==0 && y==1 && z==0", NodeBC(ux=0)), ("z==0", NodeBC(uz=0))]
    eqid, nu, setup = model.configure_dofs(bcs, native=False)
    U, F = model.get_bc_vals(eqid, setup)
    X = model.coords
    edge = np.nonzero((np.abs(X[:, 1] - 1) < 1e-9) & (np.abs(X[:, 2] - 1) < 1e-9))[0]
    assert edge.size == 2
    OL.apply(OL.LIN2, X, edge.reshape(1, 2), eqid, 3, 1.0, OL.KEY_Y, 2.0, F)
    om = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
    st, K = om.mount_K()
    ok, _ = O.solve_system(K, U, F, nu)
    assert st == 0 and ok
    uz = U[eqid[:, 2]]
    assert np.abs(uz - np.array([0, 0, 0, 0, 3.32088, 3.1998, -4.6002, -4.32047])).max() < 1e-5
