// Device-side constitutive routines (FP64, Mandel notation) of the three materials on the hot path.
// Semantics follow the reference exactly, including its quirks:
//   calcDe                 src/mech/mat/linear-elastic.jl:98-119 (3D / plane-strain branch)
//   LinearElastic          src/mech/mat/linear-elastic.jl:125-137
//   VonMises               src/mech/mat/von-mises.jl:104-156   (absolute yield tolerance 1e-8, √1.5·H, Δλ reset on elastic steps)
//   DruckerPrager          src/mech/mat/drucker-prager.jl:76-149 (cone/apex switch uses the PREVIOUS Δγ, :130)
//   tr, J2, dev, norm      src/tools/tensors.jl:10-27,136
#pragma once
#include "amaru_internal.h"

#define AM_SR2 1.4142135623730951   // src/tools/constants.jl:3
#define AM_ISR2 0.70710678118654746 // 1/SR2 is NOT used for B (the reference divides); kept for invariants only

struct MatPar {
    int kind;
    double E, nu, p2, p3, p4;
};

__device__ __forceinline__ MatPar load_mat(const int32_t *kind, const double *par, int m) {
    MatPar r;
    r.kind = kind[m];
    const double *p = par + (size_t)m * AMARU_MAT_NPARAMS;
    r.E = p[0]; r.nu = p[1]; r.p2 = p[2]; r.p3 = p[3]; r.p4 = p[4];
    return r;
}

__device__ __forceinline__ double am_tr(const double *s) { return s[0] + s[1] + s[2]; }

__device__ __forceinline__ double am_J2(const double *s) {
    const double t23 = s[3] / AM_SR2, t13 = s[4] / AM_SR2, t12 = s[5] / AM_SR2;
    const double a = s[0] - s[1], b = s[1] - s[2], c = s[2] - s[0];
    return 1.0 / 6.0 * (a * a + b * b + c * c) + t23 * t23 + t13 * t13 + t12 * t12;
}

__device__ __forceinline__ void am_dev(const double *s, double *d) {
    const double a = 2.0 / 3.0, b = -1.0 / 3.0;
    d[0] = a * s[0] + b * s[1] + b * s[2];
    d[1] = b * s[0] + a * s[1] + b * s[2];
    d[2] = b * s[0] + b * s[1] + a * s[2];
    d[3] = s[3]; d[4] = s[4]; d[5] = s[5];
}

__device__ __forceinline__ double am_norm(const double *s) {
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < 6; i++) a += s[i] * s[i];
    return sqrt(a);
}

// y = De*x for the isotropic Mandel matrix (diag c(1-ν), off-diag cν, shear c(1-2ν))
__device__ __forceinline__ void am_De_mul(double E, double nu, const double *x, double *y) {
    const double c = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
    const double d = c * (1.0 - nu), o = c * nu, g = c * (1.0 - 2.0 * nu);
    y[0] = d * x[0] + o * x[1] + o * x[2];
    y[1] = o * x[0] + d * x[1] + o * x[2];
    y[2] = o * x[0] + o * x[1] + d * x[2];
    y[3] = g * x[3]; y[4] = g * x[4]; y[5] = g * x[5];
}

// Tangent D (6x6, row-major) = calcD(mat, state).  returns 0 or AMARU_FAIL_TANGENT.
__device__ inline int am_calcD(const MatPar &mp, const double *sig, double dlam, double *D) {
    const double c = mp.E / ((1.0 + mp.nu) * (1.0 - 2.0 * mp.nu));
#pragma unroll
    for (int i = 0; i < 36; i++) D[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) D[6 * i + j] = (i == j) ? c * (1.0 - mp.nu) : c * mp.nu;
        D[6 * (i + 3) + i + 3] = c * (1.0 - 2.0 * mp.nu);
    }
    if (mp.kind == AMARU_MAT_LINEAR_ELASTIC || dlam == 0.0) return 0;
    if (mp.kind == AMARU_MAT_VON_MISES) {
        const double H = mp.p3;
        if (!(am_J2(sig) > 0.0)) return AMARU_FAIL_TANGENT;  // @assert j2d>0, von-mises.jl:117
        double s[6], n[6], Dn[6];
        am_dev(sig, s);
        const double ns = am_norm(s);
#pragma unroll
        for (int i = 0; i < 6; i++) n[i] = sqrt(1.5) * s[i] / ns;
        am_De_mul(mp.E, mp.nu, n, Dn);
        double den = 0.0;
#pragma unroll
        for (int i = 0; i < 6; i++) den += n[i] * Dn[i];
        den -= sqrt(1.5) * (-H);
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int j = 0; j < 6; j++) D[6 * i + j] -= Dn[i] * Dn[j] / den;
        return 0;
    }
    // Drucker-Prager
    const double alpha = mp.p2, H = mp.p4;
    double V[6], Nu[6];
    const double j2 = am_J2(sig);
    if (j2 != 0.0) {
        double s[6];
        am_dev(sig, s);
        const double ns = am_norm(s);
#pragma unroll
        for (int i = 0; i < 6; i++) V[i] = alpha * (i < 3 ? 1.0 : 0.0) + (s[i] / ns) / sqrt(2.0);
        const double nv = am_norm(V);
#pragma unroll
        for (int i = 0; i < 6; i++) Nu[i] = V[i] / nv;
    } else {
#pragma unroll
        for (int i = 0; i < 6; i++) Nu[i] = V[i] = (i < 3 ? 1.0 / sqrt(3.0) : 0.0);
    }
    double DNu[6], VD[6];
    am_De_mul(mp.E, mp.nu, Nu, DNu);
    am_De_mul(mp.E, mp.nu, V, VD);
    double den = H;
#pragma unroll
    for (int i = 0; i < 6; i++) den += VD[i] * Nu[i];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) D[6 * i + j] -= DNu[i] * VD[j] / den;
    return 0;
}

// Stress update at one IP: state (sig, eps, epa, dlam) in/out, deps in, dsig out.  returns 0 or AMARU_FAIL_MATERIAL.
__device__ inline int am_update(const MatPar &mp, double *sig, double *eps, double &epa, double &dlam,
                                const double *deps, double *dsig) {
    double ds[6], sini[6], str[6];
    am_De_mul(mp.E, mp.nu, deps, ds);
#pragma unroll
    for (int i = 0; i < 6; i++) sini[i] = sig[i];
    if (mp.kind == AMARU_MAT_LINEAR_ELASTIC) {
#pragma unroll
        for (int i = 0; i < 6; i++) {
            eps[i] += deps[i];
            sig[i] += ds[i];
            dsig[i] = ds[i];
        }
        return 0;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) str[i] = sig[i] + ds[i];
    const double E = mp.E, nu = mp.nu;
    if (mp.kind == AMARU_MAT_VON_MISES) {
        const double fy = mp.p2, H = mp.p3;
        const double ftr = sqrt(3.0 * am_J2(str)) - fy - H * epa;
        if (ftr < 1e-8) {
            dlam = 0.0;
#pragma unroll
            for (int i = 0; i < 6; i++) sig[i] = str[i];
        } else {
            const double G = E / (2.0 * (1.0 + nu));
            const double j2tr = am_J2(str);
            dlam = ftr / (3.0 * G + sqrt(1.5) * H);
            if (!(sqrt(j2tr) - dlam * sqrt(3.0) * G >= 0.0)) return AMARU_FAIL_MATERIAL;
            double s[6];
            am_dev(str, s);
            const double f = 1.0 - sqrt(3.0) * G * dlam / sqrt(j2tr);
#pragma unroll
            for (int i = 0; i < 6; i++) s[i] *= f;
            const double ns = am_norm(s);
#pragma unroll
            for (int i = 0; i < 6; i++) sig[i] = str[i] - sqrt(6.0) * G * dlam * s[i] / ns;
            epa += dlam;
        }
    } else {
        const double alpha = mp.p2, kappa = mp.p3, H = mp.p4;
        const double ftr = alpha * am_tr(str) + sqrt(am_J2(str)) - kappa - H * epa;
        if (ftr < 1.e-8) {
            dlam = 0.0;
#pragma unroll
            for (int i = 0; i < 6; i++) sig[i] = str[i];
        } else {
            const double K = E / (3.0 * (1.0 - 2.0 * nu)), G = E / (2.0 * (1.0 + nu));
            const double n = 1.0 / sqrt(3.0 * alpha * alpha + 0.5);
            const double j1tr = am_tr(str), j2tr = am_J2(str);
            double s[6];
            am_dev(str, s);
            if (sqrt(j2tr) - dlam * n * G > 0.0) {  // previous Δγ (drucker-prager.jl:130)
                dlam = ftr / (9 * alpha * alpha * n * K + n * G + H);
                const double j1 = j1tr - 9 * dlam * alpha * n * K;
                const double mm = 1.0 - dlam * n * G / sqrt(j2tr);
#pragma unroll
                for (int i = 0; i < 6; i++) sig[i] = mm * s[i] + (i < 3 ? j1 / 3.0 : 0.0);
            } else {
                dlam = (alpha * j1tr - kappa - H * epa) / (3 * sqrt(3.0) * alpha * K + H);
                const double j1 = j1tr - 3 * sqrt(3.0) * dlam * K;
#pragma unroll
                for (int i = 0; i < 6; i++) sig[i] = (i < 3 ? j1 / 3.0 : 0.0);
            }
            epa += dlam;
        }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
        eps[i] += deps[i];
        dsig[i] = sig[i] - sini[i];
    }
    return 0;
}

// Jacobian of one IP: J = C'*dNdR, its inverse and determinant.  X: [NN][ND] (smem), dN: [NN][ND].
template <int NN, int ND>
__device__ __forceinline__ double am_jacobian(const double *X, const double *dN, double *Ji) {
    double J[ND * ND];
#pragma unroll
    for (int i = 0; i < ND * ND; i++) J[i] = 0.0;
    for (int a = 0; a < NN; a++) {
#pragma unroll
        for (int i = 0; i < ND; i++)
#pragma unroll
            for (int j = 0; j < ND; j++) J[ND * i + j] += X[a * ND + i] * dN[a * ND + j];
    }
    double det;
    if constexpr (ND == 2) {
        det = J[0] * J[3] - J[1] * J[2];
        Ji[0] = J[3] / det; Ji[1] = -J[1] / det; Ji[2] = -J[2] / det; Ji[3] = J[0] / det;
    } else {
        const double c0 = J[4] * J[8] - J[5] * J[7], c1 = J[5] * J[6] - J[3] * J[8], c2 = J[3] * J[7] - J[4] * J[6];
        det = J[0] * c0 + J[1] * c1 + J[2] * c2;
        Ji[0] = c0 / det; Ji[1] = (J[2] * J[7] - J[1] * J[8]) / det; Ji[2] = (J[1] * J[5] - J[2] * J[4]) / det;
        Ji[3] = c1 / det; Ji[4] = (J[0] * J[8] - J[2] * J[6]) / det; Ji[5] = (J[2] * J[3] - J[0] * J[5]) / det;
        Ji[6] = c2 / det; Ji[7] = (J[1] * J[6] - J[0] * J[7]) / det; Ji[8] = (J[0] * J[4] - J[1] * J[3]) / det;
    }
    return det;
}
