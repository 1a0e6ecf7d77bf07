/*
 * amaru_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, CPU restatement of the reference algorithm (NumSoftware/Amaru.jl v0.7.1, pure Julia) for the
 * mechanical Newton-iteration hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libamaru_b200.so) never does.
 *
 * PARITY PINNING: the reference is Julia-only and Julia is not installed in the build container, so this
 * restatement cannot be diffed against reference outputs at 1e-12.  It is pinned against every known answer
 * the reference's own tests hold for this path (tests/test_oracle_golden.py): elastic-quad4 (0.3125/-0.9375),
 * elastic-hex8 load cases (uz = 4, 1.51044/-2.4501/..., -0.5), elastic-elems (-0.012,-0.095), vm-3d
 * (fz ~ -30 +- 0.7), dp (.success), structured node counts, shape partition-of-unity / finite differences,
 * tensor invariants.  Below that (1e-5 elastic, ~2 % plastic) parity with the reference is UNPINNED.
 *
 * Each function cites the reference file:line it follows.  Layouts: dense matrices row-major; Mandel order
 * (xx,yy,zz,sqrt2*yz,sqrt2*xz,sqrt2*xy); state arrays [nip_total*6] / [nip_total], element-major then IP.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SHAPE_QUAD4 1
#define SHAPE_QUAD8 2
#define SHAPE_HEX8 3
#define SHAPE_HEX20 4
#define SHAPE_TET10 5
#define MAT_LE 1
#define MAT_VM 2
#define MAT_DP 3
#define MAT_LE_PS 4 /* LinearElastic under stressmodel = :planestress (oracle-side kind only) */
#define NPAR 8
#define MAXNN 20
#define MAXNE 60

static const double SR2 = 1.4142135623730951; /* src/tools/constants.jl:3 */

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------ shapes */
int orc_shape_nn(int shape) {
    switch (shape) {
    case SHAPE_QUAD4: return 4;
    case SHAPE_QUAD8: return 8;
    case SHAPE_HEX8: return 8;
    case SHAPE_HEX20: return 20;
    case SHAPE_TET10: return 10;
    }
    return 0;
}
int orc_shape_ndim(int shape) { return (shape == SHAPE_QUAD4 || shape == SHAPE_QUAD8) ? 2 : 3; }

/* default quadrature: src/shape/quadrature.jl:62-66 (QUAD_IP4), :110-114 (TET_IP4), :167-175 (HEX_IP8);
 * defaults per shape: solids2d.jl:327,427, solids3d.jl:206,519,706.  ips rows = (r,s,t,w). */
int orc_quadrature(int shape, double *ips) {
    const double g = 0.577350269189626;
    if (shape == SHAPE_QUAD4 || shape == SHAPE_QUAD8) {
        int q = 0;
        for (int j = -1; j <= 1; j += 2)
            for (int i = -1; i <= 1; i += 2, q++) {
                ips[4 * q + 0] = i * g; ips[4 * q + 1] = j * g; ips[4 * q + 2] = 0.0; ips[4 * q + 3] = 1.0;
            }
        return 4;
    }
    if (shape == SHAPE_HEX8 || shape == SHAPE_HEX20) {
        int q = 0;
        for (int k = -1; k <= 1; k += 2)
            for (int j = -1; j <= 1; j += 2)
                for (int i = -1; i <= 1; i += 2, q++) {
                    ips[4 * q + 0] = i * g; ips[4 * q + 1] = j * g; ips[4 * q + 2] = k * g; ips[4 * q + 3] = 1.0;
                }
        return 8;
    }
    if (shape == SHAPE_TET10) {
        const double a = 0.5854101966249685, b = 0.1381966011250105, w = 0.04166666666666667;
        const double t[4][4] = {{a, b, b, w}, {b, a, b, w}, {b, b, a, w}, {b, b, b, w}};
        memcpy(ips, t, sizeof t);
        return 4;
    }
    return 0;
}

/* N(R): solids2d.jl:295-302 (QUAD4), :371-386 (QUAD8); solids3d.jl:124-147 (TET10), :472-484 (HEX8),
 * :584-613 (HEX20) */
int orc_shape_func(int shape, const double *R, double *N) {
    const double r = R[0], s = R[1], t = R[2];
    if (shape == SHAPE_QUAD4) {
        N[0] = 0.25 * (1.0 - r - s + r * s); N[1] = 0.25 * (1.0 + r - s - r * s);
        N[2] = 0.25 * (1.0 + r + s + r * s); N[3] = 0.25 * (1.0 - r + s - r * s);
        return 0;
    }
    if (shape == SHAPE_QUAD8) {
        const double rp = 1.0 + r, rm = 1.0 - r, sp = 1.0 + s, sm = 1.0 - s;
        N[0] = 0.25 * rm * sm * (rm + sm - 3.0); N[1] = 0.25 * rp * sm * (rp + sm - 3.0);
        N[2] = 0.25 * rp * sp * (rp + sp - 3.0); N[3] = 0.25 * rm * sp * (rm + sp - 3.0);
        N[4] = 0.5 * sm * (1.0 - r * r); N[5] = 0.5 * rp * (1.0 - s * s);
        N[6] = 0.5 * sp * (1.0 - r * r); N[7] = 0.5 * rm * (1.0 - s * s);
        return 0;
    }
    if (shape == SHAPE_HEX8) {
        const double sg[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                 {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
        for (int i = 0; i < 8; i++) {
            const double a = sg[i][0], b = sg[i][1], c = sg[i][2];
            /* expanded trilinear form of solids3d.jl:475-482 */
            N[i] = 0.125 * (1.0 + a * r + b * s + a * b * r * s + c * t + b * c * s * t + a * c * r * t + a * b * c * r * s * t);
        }
        return 0;
    }
    if (shape == SHAPE_HEX20) {
        const double rp = 1.0 + r, rm = 1.0 - r, sp = 1.0 + s, sm = 1.0 - s, tp = 1.0 + t, tm = 1.0 - t;
        N[0] = 0.125 * rm * sm * tm * (-r - s - t - 2.0); N[1] = 0.125 * rp * sm * tm * (r - s - t - 2.0);
        N[2] = 0.125 * rp * sp * tm * (r + s - t - 2.0);  N[3] = 0.125 * rm * sp * tm * (-r + s - t - 2.0);
        N[4] = 0.125 * rm * sm * tp * (-r - s + t - 2.0); N[5] = 0.125 * rp * sm * tp * (r - s + t - 2.0);
        N[6] = 0.125 * rp * sp * tp * (r + s + t - 2.0);  N[7] = 0.125 * rm * sp * tp * (-r + s + t - 2.0);
        N[8] = 0.25 * (1.0 - r * r) * sm * tm;  N[9] = 0.25 * rp * (1.0 - s * s) * tm;
        N[10] = 0.25 * (1.0 - r * r) * sp * tm; N[11] = 0.25 * rm * (1.0 - s * s) * tm;
        N[12] = 0.25 * (1.0 - r * r) * sm * tp; N[13] = 0.25 * rp * (1.0 - s * s) * tp;
        N[14] = 0.25 * (1.0 - r * r) * sp * tp; N[15] = 0.25 * rm * (1.0 - s * s) * tp;
        N[16] = 0.25 * rm * sm * (1.0 - t * t); N[17] = 0.25 * rp * sm * (1.0 - t * t);
        N[18] = 0.25 * rp * sp * (1.0 - t * t); N[19] = 0.25 * rm * sp * (1.0 - t * t);
        return 0;
    }
    if (shape == SHAPE_TET10) {
        const double u = 1.0 - r - s - t;
        N[0] = u * (2.0 * u - 1.0); N[1] = r * (2.0 * r - 1.0); N[2] = s * (2.0 * s - 1.0); N[3] = t * (2.0 * t - 1.0);
        N[4] = 4.0 * u * r; N[5] = 4.0 * r * s; N[6] = 4.0 * s * u; N[7] = 4.0 * u * t; N[8] = 4.0 * r * t; N[9] = 4.0 * s * t;
        return 0;
    }
    return -1;
}

/* dN/dR (nn x nd, row-major): solids2d.jl:305-312 (QUAD4), :388-412 (QUAD8); solids3d.jl:149-191 (TET10),
 * :487-504 (HEX8), :616-691 (HEX20) */
int orc_shape_deriv(int shape, const double *R, double *D) {
    const double r = R[0], s = R[1], t = R[2];
    if (shape == SHAPE_QUAD4) {
        const double d[4][2] = {{0.25 * (-1.0 + s), 0.25 * (-1.0 + r)}, {0.25 * (1.0 - s), 0.25 * (-1.0 - r)},
                                {0.25 * (1.0 + s), 0.25 * (1.0 + r)},   {0.25 * (-1.0 - s), 0.25 * (1.0 - r)}};
        memcpy(D, d, sizeof d);
        return 0;
    }
    if (shape == SHAPE_QUAD8) {
        const double rp = 1.0 + r, rm = 1.0 - r, sp = 1.0 + s, sm = 1.0 - s;
        D[0] = -0.25 * sm * (rm + rm + sm - 3.0); D[1] = -0.25 * rm * (sm + rm + sm - 3.0);
        D[2] = 0.25 * sm * (rp + rp + sm - 3.0);  D[3] = -0.25 * rp * (sm + rp + sm - 3.0);
        D[4] = 0.25 * sp * (rp + rp + sp - 3.0);  D[5] = 0.25 * rp * (sp + rp + sp - 3.0);
        D[6] = -0.25 * sp * (rm + rm + sp - 3.0); D[7] = 0.25 * rm * (sp + rm + sp - 3.0);
        D[8] = -r * sm;                D[9] = -0.5 * (1.0 - r * r);
        D[10] = 0.5 * (1.0 - s * s);   D[11] = -s * rp;
        D[12] = -r * sp;               D[13] = 0.5 * (1.0 - r * r);
        D[14] = -0.5 * (1.0 - s * s);  D[15] = -s * rm;
        return 0;
    }
    if (shape == SHAPE_HEX8) {
        const double sg[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                 {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
        const double st = s * t, rt = r * t, rs = r * s;
        for (int i = 0; i < 8; i++) {
            const double a = sg[i][0], b = sg[i][1], c = sg[i][2];
            /* solids3d.jl:492-499: D[i,1] = a + ab*s + ac*t + abc*st, etc., then D = 0.125*D (:501) */
            D[3 * i + 0] = 0.125 * (a + a * b * s + a * c * t + a * b * c * st);
            D[3 * i + 1] = 0.125 * (b + a * b * r + b * c * t + a * b * c * rt);
            D[3 * i + 2] = 0.125 * (c + a * c * r + b * c * s + a * b * c * rs);
        }
        return 0;
    }
    if (shape == SHAPE_HEX20) {
        const double rp = 1.0 + r, rm = 1.0 - r, sp = 1.0 + s, sm = 1.0 - s, tp = 1.0 + t, tm = 1.0 - t;
#define DD(i, j) D[3 * (i) + (j)]
        /* d/dr */
        DD(0, 0) = -0.125 * sm * tm * (-r - s - t - 2) - 0.125 * rm * sm * tm;
        DD(1, 0) = 0.125 * sm * tm * (r - s - t - 2) + 0.125 * rp * sm * tm;
        DD(2, 0) = 0.125 * sp * tm * (r + s - t - 2) + 0.125 * rp * sp * tm;
        DD(3, 0) = -0.125 * sp * tm * (-r + s - t - 2) - 0.125 * rm * sp * tm;
        DD(4, 0) = -0.125 * sm * tp * (-r - s + t - 2) - 0.125 * rm * sm * tp;
        DD(5, 0) = 0.125 * sm * tp * (r - s + t - 2) + 0.125 * rp * sm * tp;
        DD(6, 0) = 0.125 * sp * tp * (r + s + t - 2) + 0.125 * rp * sp * tp;
        DD(7, 0) = -0.125 * sp * tp * (-r + s + t - 2) - 0.125 * rm * sp * tp;
        DD(8, 0) = -0.5 * r * sm * tm;  DD(9, 0) = 0.25 * (1 - s * s) * tm;
        DD(10, 0) = -0.5 * r * sp * tm; DD(11, 0) = -0.25 * (1 - s * s) * tm;
        DD(12, 0) = -0.5 * r * sm * tp; DD(13, 0) = 0.25 * (1 - s * s) * tp;
        DD(14, 0) = -0.5 * r * sp * tp; DD(15, 0) = -0.25 * (1 - s * s) * tp;
        DD(16, 0) = -0.25 * sm * (1 - t * t); DD(17, 0) = 0.25 * sm * (1 - t * t);
        DD(18, 0) = 0.25 * sp * (1 - t * t);  DD(19, 0) = -0.25 * sp * (1 - t * t);
        /* d/ds */
        DD(0, 1) = -0.125 * rm * tm * (-r - s - t - 2) - 0.125 * rm * sm * tm;
        DD(1, 1) = -0.125 * rp * tm * (r - s - t - 2) - 0.125 * rp * sm * tm;
        DD(2, 1) = 0.125 * rp * tm * (r + s - t - 2) + 0.125 * rp * sp * tm;
        DD(3, 1) = 0.125 * rm * tm * (-r + s - t - 2) + 0.125 * rm * sp * tm;
        DD(4, 1) = -0.125 * rm * tp * (-r - s + t - 2) - 0.125 * rm * sm * tp;
        DD(5, 1) = -0.125 * rp * tp * (r - s + t - 2) - 0.125 * rp * sm * tp;
        DD(6, 1) = 0.125 * rp * tp * (r + s + t - 2) + 0.125 * rp * sp * tp;
        DD(7, 1) = 0.125 * rm * tp * (-r + s + t - 2) + 0.125 * rm * sp * tp;
        DD(8, 1) = -0.25 * (1 - r * r) * tm;  DD(9, 1) = -0.5 * s * rp * tm;
        DD(10, 1) = 0.25 * (1 - r * r) * tm;  DD(11, 1) = -0.5 * s * rm * tm;
        DD(12, 1) = -0.25 * (1 - r * r) * tp; DD(13, 1) = -0.5 * s * rp * tp;
        DD(14, 1) = 0.25 * (1 - r * r) * tp;  DD(15, 1) = -0.5 * s * rm * tp;
        DD(16, 1) = -0.25 * rm * (1 - t * t); DD(17, 1) = -0.25 * rp * (1 - t * t);
        DD(18, 1) = 0.25 * rp * (1 - t * t);  DD(19, 1) = 0.25 * rm * (1 - t * t);
        /* d/dt */
        DD(0, 2) = -0.125 * rm * sm * (-r - s - t - 2) - 0.125 * rm * sm * tm;
        DD(1, 2) = -0.125 * rp * sm * (r - s - t - 2) - 0.125 * rp * sm * tm;
        DD(2, 2) = -0.125 * rp * sp * (r + s - t - 2) - 0.125 * rp * sp * tm;
        DD(3, 2) = -0.125 * rm * sp * (-r + s - t - 2) - 0.125 * rm * sp * tm;
        DD(4, 2) = 0.125 * rm * sm * (-r - s + t - 2) + 0.125 * rm * sm * tp;
        DD(5, 2) = 0.125 * rp * sm * (r - s + t - 2) + 0.125 * rp * sm * tp;
        DD(6, 2) = 0.125 * rp * sp * (r + s + t - 2) + 0.125 * rp * sp * tp;
        DD(7, 2) = 0.125 * rm * sp * (-r + s + t - 2) + 0.125 * rm * sp * tp;
        DD(8, 2) = -0.25 * (1 - r * r) * sm;  DD(9, 2) = -0.25 * rp * (1 - s * s);
        DD(10, 2) = -0.25 * (1 - r * r) * sp; DD(11, 2) = -0.25 * rm * (1 - s * s);
        DD(12, 2) = 0.25 * (1 - r * r) * sm;  DD(13, 2) = 0.25 * rp * (1 - s * s);
        DD(14, 2) = 0.25 * (1 - r * r) * sp;  DD(15, 2) = 0.25 * rm * (1 - s * s);
        DD(16, 2) = -0.5 * t * rm * sm; DD(17, 2) = -0.5 * t * rp * sm;
        DD(18, 2) = -0.5 * t * rp * sp; DD(19, 2) = -0.5 * t * rm * sp;
#undef DD
        return 0;
    }
    if (shape == SHAPE_TET10) {
        const double q = 4.0 * (r + s + t) - 3.0;
        const double d[10][3] = {{q, q, q},
                                 {4.0 * r - 1.0, 0.0, 0.0},
                                 {0.0, 4.0 * s - 1.0, 0.0},
                                 {0.0, 0.0, 4.0 * t - 1.0},
                                 {4.0 - 8.0 * r - 4.0 * s - 4.0 * t, -4.0 * r, -4.0 * r},
                                 {4.0 * s, 4.0 * r, 0.0},
                                 {-4.0 * s, 4.0 - 4.0 * r - 8.0 * s - 4.0 * t, -4.0 * s},
                                 {-4.0 * t, -4.0 * t, 4.0 - 4.0 * r - 4.0 * s - 8.0 * t},
                                 {4.0 * t, 0.0, 4.0 * r},
                                 {0.0, 4.0 * t, 4.0 * s}};
        memcpy(D, d, sizeof d);
        return 0;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------------ tensors */
/* src/tools/tensors.jl:21 (tr), :24-27 (J2), :10-18 (dev = Psd*T), :136 (norm) */
static double tr6(const double *s) { return s[0] + s[1] + s[2]; }
double orc_J2(const double *s) {
    const double t23 = s[3] / SR2, t13 = s[4] / SR2, t12 = s[5] / SR2;
    const double a = s[0] - s[1], b = s[1] - s[2], c = s[2] - s[0];
    return 1.0 / 6.0 * (a * a + b * b + c * c) + t23 * t23 + t13 * t13 + t12 * t12;
}
void orc_dev(const double *s, double *d) {
    const double a = 2.0 / 3.0, b = -1.0 / 3.0;
    d[0] = a * s[0] + b * s[1] + b * s[2];
    d[1] = b * s[0] + a * s[1] + b * s[2];
    d[2] = b * s[0] + b * s[1] + a * s[2];
    d[3] = s[3]; d[4] = s[4]; d[5] = s[5];
}
static double norm6(const double *s) {
    double a = 0;
    for (int i = 0; i < 6; i++) a += s[i] * s[i];
    return sqrt(a);
}

/* ------------------------------------------------------------------------------------------------ materials */
/* calcDe, 3D / plane-strain branch: src/mech/mat/linear-elastic.jl:98-119 */
void orc_calcDe(double E, double nu, double *D) {
    const double c = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
    memset(D, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) D[6 * i + j] = (i == j) ? c * (1.0 - nu) : c * nu;
    for (int i = 3; i < 6; i++) D[6 * i + i] = c * (1.0 - 2.0 * nu);
}

/* calcDe, plane-stress branch: src/mech/mat/linear-elastic.jl:99-108 (the zz row and column are zero) */
void orc_calcDe_ps(double E, double nu, double *D) {
    const double c = E / (1.0 - nu * nu);
    memset(D, 0, 36 * sizeof(double));
    D[0] = c; D[1] = c * nu;
    D[6] = c * nu; D[7] = c;
    for (int i = 3; i < 6; i++) D[6 * i + i] = c * (1.0 - nu);
}

static void matvec6(const double *D, const double *x, double *y) {
    for (int i = 0; i < 6; i++) {
        double a = 0;
        for (int j = 0; j < 6; j++) a += D[6 * i + j] * x[j];
        y[i] = a;
    }
}

/* calcD: linear-elastic.jl:125-127; von-mises.jl:112-125; drucker-prager.jl:86-109.
 * returns 0 ok, 6 if the J2>0 assertion of von-mises.jl:117 fails. */
int orc_calcD(int kind, const double *par, const double *sig, double dlam, double *D) {
    const double E = par[0], nu = par[1];
    if (kind == MAT_LE_PS) {
        orc_calcDe_ps(E, nu, D);
        return 0;
    }
    orc_calcDe(E, nu, D);
    if (kind == MAT_LE) return 0;
    if (dlam == 0.0) return 0;
    if (kind == MAT_VM) {
        const double H = par[3];
        if (!(orc_J2(sig) > 0)) return 6;
        double s[6], n[6], Dn[6];
        orc_dev(sig, s);
        const double ns = norm6(s);
        for (int i = 0; i < 6; i++) n[i] = sqrt(1.5) * s[i] / ns;    /* dfdσ = √1.5 s/‖s‖ */
        const double dfdep = -H;
        matvec6(D, n, Dn);                                            /* De*dfdσ (De symmetric) */
        double den = 0;
        for (int i = 0; i < 6; i++) den += n[i] * Dn[i];
        den -= sqrt(1.5) * dfdep;
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) D[6 * i + j] -= Dn[i] * Dn[j] / den;
        return 0;
    }
    if (kind == MAT_DP) {
        const double alpha = par[2], H = par[4];
        double V[6], Nu[6];
        const double j2 = orc_J2(sig);
        if (j2 != 0.0) {
            double s[6];
            orc_dev(sig, s);
            const double ns = norm6(s);
            for (int i = 0; i < 6; i++) V[i] = alpha * (i < 3 ? 1.0 : 0.0) + (s[i] / ns) / sqrt(2.0);
            const double nv = norm6(V);
            for (int i = 0; i < 6; i++) Nu[i] = V[i] / nv;
        } else { /* apex */
            for (int i = 0; i < 6; i++) Nu[i] = V[i] = (i < 3 ? 1.0 / sqrt(3.0) : 0.0);
        }
        double DNu[6], VD[6];
        matvec6(D, Nu, DNu);
        matvec6(D, V, VD); /* V'*De = (De*V)' */
        double den = H;
        for (int i = 0; i < 6; i++) den += VD[i] * Nu[i];
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) D[6 * i + j] -= DNu[i] * VD[j] / den;
        return 0;
    }
    return -1;
}

/* update_state!: linear-elastic.jl:130-137; von-mises.jl:128-156; drucker-prager.jl:112-149.
 * sig, eps (6), epa, dlam are updated in place; dsig receives Δσ.  returns 0 ok, 1 failure. */
int orc_update_ip(int kind, const double *par, double *sig, double *eps, double *epa, double *dlam,
                  const double *deps, double *dsig) {
    const double E = par[0], nu = par[1];
    double De[36], ds[6], sini[6], str[6];
    if (kind == MAT_LE_PS) orc_calcDe_ps(E, nu, De);
    else orc_calcDe(E, nu, De);
    matvec6(De, deps, ds);
    memcpy(sini, sig, sizeof sini);
    if (kind == MAT_LE || kind == MAT_LE_PS) {
        for (int i = 0; i < 6; i++) { eps[i] += deps[i]; sig[i] += ds[i]; dsig[i] = ds[i]; }
        return 0;
    }
    for (int i = 0; i < 6; i++) str[i] = sig[i] + ds[i];
    if (kind == MAT_VM) {
        const double fy = par[2], H = par[3];
        const double ftr = sqrt(3.0 * orc_J2(str)) - fy - H * (*epa);          /* yield_func :104-109 */
        if (ftr < 1e-8) {
            *dlam = 0.0;
            memcpy(sig, str, sizeof str);
        } else {
            const double G = E / (2.0 * (1.0 + nu));
            const double j2tr = orc_J2(str);
            *dlam = ftr / (3.0 * G + sqrt(1.5) * H);
            if (!(sqrt(j2tr) - (*dlam) * sqrt(3.0) * G >= 0.0)) return 1;      /* :146 (state.Δλ already set) */
            double s[6];
            orc_dev(str, s);
            const double f = 1.0 - sqrt(3.0) * G * (*dlam) / sqrt(j2tr);
            for (int i = 0; i < 6; i++) s[i] *= f;
            const double ns = norm6(s);
            for (int i = 0; i < 6; i++) sig[i] = str[i] - sqrt(6.0) * G * (*dlam) * s[i] / ns;
            *epa += *dlam;
        }
    } else if (kind == MAT_DP) {
        const double alpha = par[2], kappa = par[3], H = par[4];
        const double ftr = alpha * tr6(str) + sqrt(orc_J2(str)) - kappa - H * (*epa);  /* :76-83 */
        if (ftr < 1.e-8) {
            *dlam = 0.0;
            memcpy(sig, str, sizeof str);
        } else {
            const double K = E / (3.0 * (1.0 - 2.0 * nu)), G = E / (2.0 * (1.0 + nu));
            const double n = 1.0 / sqrt(3.0 * alpha * alpha + 0.5);
            const double j1tr = tr6(str), j2tr = orc_J2(str);
            double s[6];
            orc_dev(str, s);
            if (sqrt(j2tr) - (*dlam) * n * G > 0.0) {                          /* :130 uses the PREVIOUS Δγ */
                *dlam = ftr / (9 * alpha * alpha * n * K + n * G + H);
                const double j1 = j1tr - 9 * (*dlam) * alpha * n * K;
                const double m = 1.0 - (*dlam) * n * G / sqrt(j2tr);
                for (int i = 0; i < 6; i++) sig[i] = m * s[i] + (i < 3 ? j1 / 3.0 : 0.0);
            } else { /* apex */
                *dlam = (alpha * j1tr - kappa - H * (*epa)) / (3 * sqrt(3.0) * alpha * K + H);
                const double j1 = j1tr - 3 * sqrt(3.0) * (*dlam) * K;
                for (int i = 0; i < 6; i++) sig[i] = (i < 3 ? j1 / 3.0 : 0.0);
            }
            *epa += *dlam;
        }
    } else
        return -1;
    for (int i = 0; i < 6; i++) { eps[i] += deps[i]; dsig[i] = sig[i] - sini[i]; }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ element */
/* J = C'*dNdR, dNdX = dNdR*inv(J), detJ  (mech-solid.jl:146-150).  C nn x nd, dNdR nn x nd. */
static double jacobian(int nn, int nd, const double *C, const double *dNdR, double *dNdX) {
    double J[9] = {0}, Ji[9];
    for (int a = 0; a < nn; a++)
        for (int i = 0; i < nd; i++)
            for (int j = 0; j < nd; j++) J[nd * i + j] += C[nd * a + i] * dNdR[nd * a + j];
    double det;
    if (nd == 2) {
        det = J[0] * J[3] - J[1] * J[2];
        Ji[0] = J[3] / det; Ji[1] = -J[1] / det; Ji[2] = -J[2] / det; Ji[3] = J[0] / det;
    } else {
        const double c0 = J[4] * J[8] - J[5] * J[7], c1 = J[5] * J[6] - J[3] * J[8], c2 = J[3] * J[7] - J[4] * J[6];
        det = J[0] * c0 + J[1] * c1 + J[2] * c2;
        Ji[0] = c0 / det; Ji[1] = (J[2] * J[7] - J[1] * J[8]) / det; Ji[2] = (J[1] * J[5] - J[2] * J[4]) / det;
        Ji[3] = c1 / det; Ji[4] = (J[0] * J[8] - J[2] * J[6]) / det; Ji[5] = (J[2] * J[3] - J[0] * J[5]) / det;
        Ji[6] = c2 / det; Ji[7] = (J[1] * J[6] - J[0] * J[7]) / det; Ji[8] = (J[0] * J[4] - J[1] * J[3]) / det;
    }
    for (int a = 0; a < nn; a++)
        for (int j = 0; j < nd; j++) {
            double v = 0;
            for (int k = 0; k < nd; k++) v += dNdR[nd * a + k] * Ji[nd * k + j];
            dNdX[nd * a + j] = v;
        }
    return det;
}

/* stressmodel = :axisymmetric (mech-solver.jl:9): the element routines below read this switch, set per model by
 * orc_set_axisymmetric (the python wrapper sets it before every call into this file). */
static int g_axi = 0;
void orc_set_axisymmetric(int on) { g_axi = on; }

/* ip.coord.x = (C'N)[1]: the radius of the integration point (ip coordinates are set from the shape functions, fe-model.jl) */
static double ip_radius(int nn, int nd, const double *C, const double *N) {
    double r = 0;
    for (int a = 0; a < nn; a++) r += N[a] * C[nd * a];
    return r;
}

/* setB (mech-solid.jl:82-121).  B is 6 x ne row-major, pre-zeroed.  N, r: shape functions and radius of the integration
 * point, used by the axisymmetric branch only (:94-108: hoop strain N_i/r in row 3). */
static void setB(int nn, int nd, const double *dNdX, double *B, const double *N, double r) {
    const int ne = nn * nd;
    if (nd == 2) {
        for (int i = 0; i < nn; i++) {
            B[0 * ne + 0 + i * 2] = dNdX[2 * i + 0];
            B[1 * ne + 1 + i * 2] = dNdX[2 * i + 1];
            if (g_axi) B[2 * ne + 0 + i * 2] = N[i] / r;
            B[5 * ne + 0 + i * 2] = dNdX[2 * i + 1] / SR2;
            B[5 * ne + 1 + i * 2] = dNdX[2 * i + 0] / SR2;
        }
    } else {
        for (int i = 0; i < nn; i++) {
            const double dx = dNdX[3 * i], dy = dNdX[3 * i + 1], dz = dNdX[3 * i + 2];
            B[0 * ne + 0 + i * 3] = dx;
            B[1 * ne + 1 + i * 3] = dy;
            B[2 * ne + 2 + i * 3] = dz;
            B[3 * ne + 1 + i * 3] = dz / SR2; B[3 * ne + 2 + i * 3] = dy / SR2;
            B[4 * ne + 0 + i * 3] = dz / SR2; B[4 * ne + 2 + i * 3] = dx / SR2;
            B[5 * ne + 0 + i * 3] = dy / SR2; B[5 * ne + 1 + i * 3] = dx / SR2;
        }
    }
}

/* elem_stiffness (mech-solid.jl:124-166): K = Σ_ip (detJ*w*th) * B' * (D*B).
 * C nn x nd; state of this element's IPs: sig [nip*6], dlam [nip].  K ne x ne row-major.
 * returns 0, 4 (detJ<=0), 6 (tangent assertion). */
int orc_elem_stiffness(int shape, double th, const double *C, int kind, const double *par, const double *sig,
                       const double *dlam, double *K) {
    const int nn = orc_shape_nn(shape), nd = orc_shape_ndim(shape), ne = nn * nd;
    double ips[8 * 4], dNdR[MAXNN * 3], dNdX[MAXNN * 3], B[6 * MAXNE], DB[6 * MAXNE], D[36], N[MAXNN];
    const int nip = orc_quadrature(shape, ips);
    memset(K, 0, sizeof(double) * ne * ne);
    memset(B, 0, sizeof B);
    for (int q = 0; q < nip; q++) {
        orc_shape_func(shape, &ips[4 * q], N);
        const double r = ip_radius(nn, nd, C, N);
        if (g_axi) th = 2.0 * M_PI * r;                                /* mech-solid.jl:143 */
        orc_shape_deriv(shape, &ips[4 * q], dNdR);
        const double detJ = jacobian(nn, nd, C, dNdR, dNdX);
        if (!(detJ > 0.0)) return 4;
        setB(nn, nd, dNdX, B, N, r);
        const double coef = detJ * ips[4 * q + 3] * th;
        const int st = orc_calcD(kind, par, &sig[6 * q], dlam[q], D);
        if (st) return st;
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < ne; j++) {
                double v = 0;
                for (int k = 0; k < 6; k++) v += D[6 * i + k] * B[k * ne + j];
                DB[i * ne + j] = v;
            }
        for (int i = 0; i < ne; i++)
            for (int j = 0; j < ne; j++) {
                double v = 0;
                for (int k = 0; k < 6; k++) v += B[k * ne + i] * DB[k * ne + j];
                K[i * ne + j] += coef * v;
            }
    }
    return 0;
}

/* elem_mass (mech-solid.jl:169-205): M = Σ_ip (ρ*detJ*w*th) N'N */
int orc_elem_mass(int shape, double th, double rho, const double *C, double *M) {
    const int nn = orc_shape_nn(shape), nd = orc_shape_ndim(shape), ne = nn * nd;
    double ips[8 * 4], dNdR[MAXNN * 3], dNdX[MAXNN * 3], N[MAXNN];
    const int nip = orc_quadrature(shape, ips);
    memset(M, 0, sizeof(double) * ne * ne);
    for (int q = 0; q < nip; q++) {
        orc_shape_func(shape, &ips[4 * q], N);
        if (g_axi) th = 2.0 * M_PI * ip_radius(nn, nd, C, N);          /* mech-solid.jl:180 */
        orc_shape_deriv(shape, &ips[4 * q], dNdR);
        const double detJ = jacobian(nn, nd, C, dNdR, dNdX);
        if (!(detJ > 0.0)) return 4;
        const double coef = rho * detJ * ips[4 * q + 3] * th;
        for (int a = 0; a < nn; a++)
            for (int b = 0; b < nn; b++)
                for (int d = 0; d < nd; d++) M[(a * nd + d) * ne + b * nd + d] += coef * N[a] * N[b];
    }
    return 0;
}

/* update_elem! (mech-solid.jl:243-279): dU (ne) -> dF (ne), IP state updated in place.
 * returns 0 or the failing status; on failure dF holds the partial sum like the reference (:273). */
int orc_update_elem(int shape, double th, const double *C, int kind, const double *par, double *sig,
                    double *eps, double *epa, double *dlam, const double *dU, double *dF) {
    const int nn = orc_shape_nn(shape), nd = orc_shape_ndim(shape), ne = nn * nd;
    double ips[8 * 4], dNdR[MAXNN * 3], dNdX[MAXNN * 3], B[6 * MAXNE], N[MAXNN];
    const int nip = orc_quadrature(shape, ips);
    memset(dF, 0, sizeof(double) * ne);
    memset(B, 0, sizeof B);
    for (int q = 0; q < nip; q++) {
        orc_shape_func(shape, &ips[4 * q], N);
        const double r = ip_radius(nn, nd, C, N);
        if (g_axi) th = 2.0 * M_PI * r;                                /* mech-solid.jl:260-262 */
        orc_shape_deriv(shape, &ips[4 * q], dNdR);
        const double detJ = jacobian(nn, nd, C, dNdR, dNdX);
        setB(nn, nd, dNdX, B, N, r);
        double de[6], ds[6];
        for (int i = 0; i < 6; i++) {
            double v = 0;
            for (int j = 0; j < ne; j++) v += B[i * ne + j] * dU[j];
            de[i] = v;
        }
        const int st = orc_update_ip(kind, par, &sig[6 * q], &eps[6 * q], &epa[q], &dlam[q], de, ds);
        if (st) return st;
        const double coef = detJ * ips[4 * q + 3] * th;
        for (int j = 0; j < ne; j++) {
            double v = 0;
            for (int i = 0; i < 6; i++) v += B[i * ne + j] * ds[i];
            dF[j] += coef * v;
        }
    }
    return 0;
}

/* elem_internal_forces (mech-solid.jl:208-240): dF = Σ coef B'σ */
int orc_elem_internal_forces(int shape, double th, const double *C, const double *sig, double *dF) {
    const int nn = orc_shape_nn(shape), nd = orc_shape_ndim(shape), ne = nn * nd;
    double ips[8 * 4], dNdR[MAXNN * 3], dNdX[MAXNN * 3], B[6 * MAXNE], N[MAXNN];
    const int nip = orc_quadrature(shape, ips);
    memset(dF, 0, sizeof(double) * ne);
    memset(B, 0, sizeof B);
    for (int q = 0; q < nip; q++) {
        orc_shape_func(shape, &ips[4 * q], N);
        const double r = ip_radius(nn, nd, C, N);
        if (g_axi) th = 2.0 * M_PI * r;                                /* mech-solid.jl:222-224 */
        orc_shape_deriv(shape, &ips[4 * q], dNdR);
        const double detJ = jacobian(nn, nd, C, dNdR, dNdX);
        setB(nn, nd, dNdX, B, N, r);
        const double coef = detJ * ips[4 * q + 3] * th;
        for (int j = 0; j < ne; j++) {
            double v = 0;
            for (int i = 0; i < 6; i++) v += B[i * ne + j] * sig[6 * q + i];
            dF[j] += coef * v;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ model loops */
static void gather_elem(int nn, int nd, const double *coords, const int32_t *conn_e, const int32_t *eqid,
                        double *C, int64_t *map) {
    for (int a = 0; a < nn; a++) {
        const int64_t n = conn_e[a];
        for (int d = 0; d < nd; d++) {
            C[a * nd + d] = coords[3 * n + d];                         /* getcoords, element.jl:112-116 */
            map[a * nd + d] = eqid[n * nd + d];                        /* mech-solid.jl:160-161 */
        }
    }
}

/* mount_K (mech-solver.jl:78-110), COO part: triplets in element order, i then j (:86-95).
 * mode 0: Ke = stiffness; mode 1: Ke = mass with rho[e].
 * filter != 0 drops |v| < eps() like :90.  rows/cols/vals must hold nelem*ne*ne entries.
 * `@withthreads` is mimicked with static contiguous chunks whose triplets are concatenated in thread order
 * (src/tools/threads.jl:97-106,51-57).  returns status; *ntrip = number of triplets written. */
int orc_mount_coo(int mode, int shape, double th, int64_t nelem, const double *coords, const int32_t *conn,
                  const int32_t *elem_mat, const int32_t *mat_kind, const double *mat_par, const double *rho,
                  const int32_t *eqid, const double *sig, const double *dlam, int filter, int64_t *rows,
                  int64_t *cols, double *vals, int64_t *ntrip) {
    const int nn = orc_shape_nn(shape), nd = orc_shape_ndim(shape), ne = nn * nd;
    double ips[32];
    const int nip = orc_quadrature(shape, ips);
    int status = 0;
    int64_t *cnt = (int64_t *)calloc((size_t)nelem + 1, sizeof(int64_t));
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < nelem; e++) {
        double C[MAXNN * 3], K[MAXNE * MAXNE];
        int64_t map[MAXNE];
        gather_elem(nn, nd, coords, &conn[e * nn], eqid, C, map);
        const int m = elem_mat[e];
        int st;
        if (mode == 0)
            st = orc_elem_stiffness(shape, th, C, mat_kind[m], &mat_par[NPAR * m], &sig[6 * nip * e], &dlam[nip * e], K);
        else
            st = orc_elem_mass(shape, th, rho[e], C, K);
        if (st) {
#pragma omp critical
            status = st;
            continue;
        }
        /* written at the unfiltered offset; compacted below */
        int64_t o = e * ne * ne, c = 0;
        for (int i = 0; i < ne; i++)
            for (int j = 0; j < ne; j++) {
                const double v = K[i * ne + j];
                if (filter && fabs(v) < 2.220446049250313e-16) continue;
                rows[o + c] = map[i]; cols[o + c] = map[j]; vals[o + c] = v;
                c++;
            }
        cnt[e] = c;
    }
    int64_t w = 0;
    for (int64_t e = 0; e < nelem; e++) {
        const int64_t o = e * ne * ne, c = cnt[e];
        if (w != o) {
            memmove(&rows[w], &rows[o], c * sizeof(int64_t));
            memmove(&cols[w], &cols[o], c * sizeof(int64_t));
            memmove(&vals[w], &vals[o], c * sizeof(double));
        }
        w += c;
    }
    free(cnt);
    *ntrip = w;
    return status;
}

/* update_state! (mech-solver.jl:124-144): ΔFin[map] += ΔF in element order per thread chunk, chunks reduced in
 * thread order (threads.jl:51-57).  State arrays updated in place.  returns 0 / 1 (material) / 2 (NaN). */
int orc_update_state(int shape, double th, int64_t nelem, const double *coords, const int32_t *conn,
                     const int32_t *elem_mat, const int32_t *mat_kind, const double *mat_par,
                     const int32_t *eqid, int64_t ndofs, double *sig, double *eps, double *epa, double *dlam,
                     const double *dU, double *dFin) {
    const int nn = orc_shape_nn(shape), nd = orc_shape_ndim(shape), ne = nn * nd;
    double ips[32];
    const int nip = orc_quadrature(shape, ips);
    int status = 0;
    int nth = orc_num_threads();
    if (nth > nelem) nth = nelem > 0 ? (int)nelem : 1;
    double *priv = (double *)calloc((size_t)nth * (size_t)ndofs, sizeof(double));
#pragma omp parallel for schedule(static) num_threads(nth)
    for (int t = 0; t < nth; t++) {
        const int64_t lo = nelem * t / nth, hi = nelem * (t + 1) / nth;
        double *F = priv + (size_t)t * ndofs;
        for (int64_t e = lo; e < hi; e++) {
            double C[MAXNN * 3], dUe[MAXNE], dFe[MAXNE];
            int64_t map[MAXNE];
            gather_elem(nn, nd, coords, &conn[e * nn], eqid, C, map);
            for (int i = 0; i < ne; i++) dUe[i] = dU[map[i]];
            const int m = elem_mat[e];
            const int st = orc_update_elem(shape, th, C, mat_kind[m], &mat_par[NPAR * m], &sig[6 * nip * e],
                                           &eps[6 * nip * e], &epa[nip * e], &dlam[nip * e], dUe, dFe);
            if (st) {
#pragma omp critical
                status = st;
                break;                                              /* mech-solver.jl:131-134 */
            }
            for (int i = 0; i < ne; i++) F[map[i]] += dFe[i];
        }
    }
    memset(dFin, 0, sizeof(double) * ndofs);
    for (int t = 0; t < nth; t++)
        for (int64_t i = 0; i < ndofs; i++) dFin[i] += priv[(size_t)t * ndofs + i];
    free(priv);
    if (status) return status;
    for (int64_t i = 0; i < ndofs; i++)
        if (isnan(dFin[i])) return 2;                               /* mech-solver.jl:142 */
    return 0;
}

/* Σ_e elem_internal_forces scattered to Fin[ndofs] */
int orc_internal_forces(int shape, double th, int64_t nelem, const double *coords, const int32_t *conn,
                        const int32_t *eqid, int64_t ndofs, const double *sig, double *Fin) {
    const int nn = orc_shape_nn(shape), nd = orc_shape_ndim(shape), ne = nn * nd;
    double ips[32];
    const int nip = orc_quadrature(shape, ips);
    memset(Fin, 0, sizeof(double) * ndofs);
    for (int64_t e = 0; e < nelem; e++) {
        double C[MAXNN * 3], dFe[MAXNE];
        int64_t map[MAXNE];
        gather_elem(nn, nd, coords, &conn[e * nn], eqid, C, map);
        orc_elem_internal_forces(shape, th, C, &sig[6 * nip * e], dFe);
        for (int i = 0; i < ne; i++) Fin[map[i]] += dFe[i];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ CPU PCG
 * Not part of the reference (which uses lu(K11), solver.jl:42-43): the same Jacobi-PCG the GPU path runs, on
 * CSR, used only as the "same algorithm on host cores" leg of the CPU baseline. */
int orc_pcg_jacobi(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const double *b,
                   double *x, double rtol, int maxit, int *iters, double *relres) {
    double *r = malloc(n * sizeof(double)), *z = malloc(n * sizeof(double)), *p = malloc(n * sizeof(double)),
           *q = malloc(n * sizeof(double)), *dinv = malloc(n * sizeof(double));
    double bb = 0, rz = 0;
#pragma omp parallel for reduction(+ : bb, rz) schedule(static)
    for (int64_t i = 0; i < n; i++) {
        double d = 1.0;
        for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++)
            if (col[k] == i) d = val[k];
        dinv[i] = 1.0 / d;
        x[i] = 0.0; r[i] = b[i]; z[i] = dinv[i] * r[i]; p[i] = z[i];
        bb += b[i] * b[i]; rz += r[i] * z[i];
    }
    int it = 0;
    double rr = bb;
    while (it < maxit && sqrt(rr) > rtol * sqrt(bb)) {
        double pq = 0;
#pragma omp parallel for reduction(+ : pq) schedule(static)
        for (int64_t i = 0; i < n; i++) {
            double a = 0;
            for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) a += val[k] * p[col[k]];
            q[i] = a; pq += p[i] * a;
        }
        const double alpha = rz / pq;
        double rz2 = 0; rr = 0;
#pragma omp parallel for reduction(+ : rz2, rr) schedule(static)
        for (int64_t i = 0; i < n; i++) {
            x[i] += alpha * p[i]; r[i] -= alpha * q[i]; z[i] = dinv[i] * r[i];
            rz2 += r[i] * z[i]; rr += r[i] * r[i];
        }
        const double beta = rz2 / rz;
        rz = rz2;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
        it++;
    }
    *iters = it;
    *relres = bb > 0 ? sqrt(rr / bb) : 0.0;
    free(r); free(z); free(p); free(q); free(dinv);
    return 0;
}
