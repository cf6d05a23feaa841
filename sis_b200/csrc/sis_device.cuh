// sis_device.cuh -- device data layout and per-term math of the SIS hot path (fp64).
//
// Every term is written in GATHER form: the thread group that owns bead k computes the
// complete force on k from each term instance k takes part in, so the step needs no atomics,
// no force zeroing pass and no second kernel ("one launch per MD step").
// The arithmetic of each term follows the reference routine cited next to it
// (file:line relative to the reference tree).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#define SIS_HD __host__ __device__ __forceinline__
#define SIS_D __device__ __forceinline__

// ---- constants (src/const_phys.F90:5-18, src/const_idx.F90) ---------------------------
#define SIS_PI 3.14159265358979323846
constexpr double SIS_EPS0 = 8.8541878128e-12;
constexpr double SIS_ELE = 1.602176634e-19;
constexpr double SIS_BOLTZ_J = 1.380649e-23;
constexpr double SIS_N_AVO = 6.02214076e23;
constexpr double SIS_KCAL2JOUL = 4184.0;
constexpr double SIS_JOUL2KCAL_MOL = (1.0 / SIS_KCAL2JOUL) * SIS_N_AVO;
constexpr double SIS_BOLTZ_KCAL_MOL = SIS_BOLTZ_J * SIS_JOUL2KCAL_MOL;
constexpr double SIS_SMALL_VALUE = 2.220446049250313e-16;

enum { SEQ_A = 0, SEQ_U = 1, SEQ_G = 2, SEQ_C = 3, SEQ_D = 4 };
enum { MOLT_RNA = 1, MOLT_CIRC = 2 };
enum { ENE_TOTAL = 0, ENE_BOND, ENE_ANGL, ENE_DIHE, ENE_BP, ENE_EXV, ENE_ELE, ENE_STAGE, ENE_TWZ,
       ENE_RG, ENE_REST, ENE_N };

// bead meta flags (int4 meta: x = chain start, y = chain length, z = local index, w = flags)
constexpr int MF_SEQ_MASK = 7, MF_CHARGE = 8, MF_CIRC = 16, MF_FROZEN = 32;
// neighbour row entry: bit31 = in WCA list, bit30 = in DH list, bit29 = the pair may need the
// periodic image shift before the next rebuild (see nl_needs_image), low 29 bits = partner bead
constexpr uint32_t NB_WCA = 0x80000000u, NB_ELE = 0x40000000u, NB_IMG = 0x20000000u, NB_IDX = 0x1FFFFFFFu;

// src/var_potential.F90:63-82 (field order kept)
struct BpPara {
    double cutoff_ddist, U0, bond_k, bond_r;
    double angl_k1, angl_k2, angl_k3, angl_k4;
    double angl_theta1, angl_theta2, angl_theta3, angl_theta4;
    double dihd_k1, dihd_k2, dihd_phi1, dihd_phi2;
    double cos_phi1, sin_phi1, cos_phi2, sin_phi2;   // of dihd_phi1/2: cos(phi + phi0) without atan2 / sincos
};

// force-field + run constants shared by all replicas (kernel argument, by value)
struct FFParams {
    double bond_k, bond_r0, angl_k, angl_t0, angl_co0, dih_k, dih_w, dih_p0;
    double wca_sigma2, wca_coef, wca_eps, wca_sigma;
    double bp_cutoff_energy, bias_ss_force;
    double stage_sigma, stage_eps, bias_rg_k;
    double box[3], box_half[3];
    int flg_angl_ReB, flg_dih_cos, flg_dih_exp, flg_ele, flg_pbc, flg_stage, flg_bias_ss;
    int bp_triplet, nstep_bp_MC, temp_independent, integrator, flg_bias_rg, bias_rg_pott;
    int ntwz_DCF, ntwz_FR, nrest;
    int mixed;            // option "precision" = 1: WCA / Debye-Hueckel pair scalars in fp32 (vectors, sums, everything else fp64)
    BpPara bp[7];
};

// per-replica parameters (device array, one per local replica)
struct RepParams {
    double tempK;          // replica temperature (rep2val(TEMP) or tempK)
    double ele_coef;       // src/set_ele.F90:81-86
    double inv_lambdaD;    // 1 / lambdaD
    double lambdaD;        // kept beside its inverse: ele_cutoff = cutoff_inp * lambdaD must be bit-identical to the host's
    double ele_cutoff;     // refreshed in neighbor_list only for cutoff_type 2 (quirk 3)
    double diele_dTcoef;
    double sqrt_tempK;     // noise scaling: md_coef_rep(1) = noise0[imp] * sqrt(T)
    double bias_rg_0;
    double twz_force;      // magnitude multiplying twz_DCF_direction
    double rg_cm[3], rg_val;  // scratch written by the Rg pre-pass
    int stream;            // RNG stream = global replica id (grep - 1)
    int pad;
};

struct Vec3 {
    double x, y, z;
};
SIS_HD Vec3 mk(double x, double y, double z) { return Vec3{x, y, z}; }
SIS_HD Vec3 operator+(Vec3 a, Vec3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
SIS_HD Vec3 operator-(Vec3 a, Vec3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
SIS_HD Vec3 operator*(double s, Vec3 a) { return mk(s * a.x, s * a.y, s * a.z); }
SIS_HD Vec3 operator-(Vec3 a) { return mk(-a.x, -a.y, -a.z); }
SIS_HD void operator+=(Vec3 &a, Vec3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
SIS_HD void operator-=(Vec3 &a, Vec3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
SIS_HD double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SIS_HD Vec3 cross(Vec3 a, Vec3 b)
{
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
SIS_HD Vec3 v3(double4 p) { return mk(p.x, p.y, p.z); }
#ifdef __CUDACC__
// one 256-bit read-only load of a position (LDG.E.ENL2.256.CONSTANT on sm_100a); only for buffers
// the running kernel never writes (the `cur` position buffer of a step)
SIS_D double4 ldg_pos(const double4 *p)
{
    double4 v;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
#endif

// src/pbc.F90:140-165 pbc_vec_d: single image, strict inequalities
SIS_HD double minimg1(double v, double L, double Lh)
{
    if (v > Lh) v -= L;
    else if (v < -Lh) v += L;
    return v;
}
SIS_HD Vec3 pbc_vec_d(const FFParams &P, Vec3 a, Vec3 b)
{
    Vec3 v = a - b;
    if (P.flg_pbc) {
        v.x = minimg1(v.x, P.box[0], P.box_half[0]);
        v.y = minimg1(v.y, P.box[1], P.box_half[1]);
        v.z = minimg1(v.z, P.box[2], P.box_half[2]);
    }
    return v;
}
// src/pbc.F90:215-223 pbc_wrap: into [0, L]
SIS_HD double wrap1(double x, double L)
{
    if (x > L) x = x - L * ((int)((x - L) / L) + 1);
    else if (x < 0.0) x = x + L * ((int)((0.0 - x) / L) + 1);
    return x;
}

SIS_HD double powi(double x, int n)
{
    double r = 1.0;
    while (n) {
        if (n & 1) r *= x;
        x *= x;
        n >>= 1;
    }
    return r;
}

// ---- bonded terms --------------------------------------------------------------------------
// Each term has a "vectors" core that evaluates the term once (used by the once-per-term tables
// of k_md_fast) and a per-position form (gather form: force on the bead at position `pos`).
// src/force_bond.F90:34-56: returns f with F1 = -f, F2 = +f
SIS_HD Vec3 bond_force_vec(const FFParams &P, Vec3 x1, Vec3 x2)
{
    Vec3 v = pbc_vec_d(P, x1, x2);
    double d = sqrt(dot(v, v));
    return (P.bond_k * (d - P.bond_r0) / d) * v;
}
SIS_HD Vec3 bond_force(const FFParams &P, Vec3 x1, Vec3 x2, int pos)
{
    Vec3 f = bond_force_vec(P, x1, x2);
    return pos == 0 ? -f : f;
}
// angles: F1 = *fa, F3 = *fc, F2 = -(fa + fc)
// src/force_angl.F90:36-82
SIS_HD void angl_force_vecs(const FFParams &P, Vec3 x1, Vec3 x2, Vec3 x3, Vec3 *fa, Vec3 *fc)
{
    Vec3 v21 = pbc_vec_d(P, x2, x1), v32 = pbc_vec_d(P, x3, x2);
    double d21 = dot(v21, v21), d32 = dot(v32, v32), d2132 = dot(v32, v21);
    double cosine = -d2132 / sqrt(d21 * d32);
    if (fabs(cosine) > 1.0) cosine = copysign(1.0, cosine);
    double delta = acos(cosine) - P.angl_t0;
    double t = d21 * d32 - d2132 * d2132;
    if (t <= 1.0) t = 1.0;
    double f = P.angl_k * delta / sqrt(t);
    Vec3 f21 = f * ((d2132 / d21) * v21 - v32);
    Vec3 f32 = f * ((d2132 / d32) * v32 - v21);
    *fa = -f21;
    *fc = f32;
}
SIS_HD Vec3 angl_force(const FFParams &P, Vec3 x1, Vec3 x2, Vec3 x3, int pos)
{
    Vec3 fa, fc;
    angl_force_vecs(P, x1, x2, x3, &fa, &fc);
    if (pos == 0) return fa;
    if (pos == 1) return -(fa + fc);   // = f21 - f32
    return fc;
}
// src/force_angl_ReB.F90:36-73
SIS_HD void angl_ReB_force_vecs(const FFParams &P, Vec3 x1, Vec3 x2, Vec3 x3, Vec3 *fa, Vec3 *fc)
{
    Vec3 v12 = pbc_vec_d(P, x1, x2), v32 = pbc_vec_d(P, x3, x2);
    double d12 = dot(v12, v12), d32 = dot(v32, v32);
    double a1232 = sqrt(d12 * d32);
    double c = dot(v12, v32) / a1232;
    double om = 1.0 - c * c;
    double f = -P.angl_k * (c - P.angl_co0) * (1.0 - c * P.angl_co0) / (om * om);
    *fa = f * ((1.0 / a1232) * v32 - (c / d12) * v12);
    *fc = f * ((1.0 / a1232) * v12 - (c / d32) * v32);
}
SIS_HD Vec3 angl_ReB_force(const FFParams &P, Vec3 x1, Vec3 x2, Vec3 x3, int pos)
{
    Vec3 fa, fc;
    angl_ReB_force_vecs(P, x1, x2, x3, &fa, &fc);
    if (pos == 0) return fa;
    if (pos == 1) return -(fa + fc);
    return fc;
}
// src/force_dih_cos.F90:21-59, src/force_dih_exp.F90:36-94: forces on the four beads
SIS_HD void dih_force_vecs(const FFParams &P, Vec3 x1, Vec3 x2, Vec3 x3, Vec3 x4, Vec3 *f)
{
    Vec3 vij = pbc_vec_d(P, x1, x2), vkj = pbc_vec_d(P, x3, x2), vkl = pbc_vec_d(P, x3, x4);
    Vec3 m = cross(vij, vkj), n = cross(vkj, vkl);
    double akj2 = dot(vkj, vkj), akj = sqrt(akj2);
    double dih = atan2(akj * dot(vij, n), dot(m, n));
    double pre = 0.0;
    if (P.flg_dih_cos) pre += -P.dih_k * sin(dih + P.dih_p0) * akj;
    if (P.flg_dih_exp) {
        double dd = dih - P.dih_p0;
        pre += (P.dih_k * P.dih_w) * dd * exp((-0.5 * P.dih_w) * dd * dd) * akj;
    }
    Vec3 fi = (-pre / dot(m, m)) * m;
    Vec3 fl = (pre / dot(n, n)) * n;
    double a = dot(vij, vkj) / akj2, b = dot(vkl, vkj) / akj2;
    f[0] = fi;
    f[1] = (-1.0 + a) * fi + (-b) * fl;
    f[2] = (-a) * fi + (-1.0 + b) * fl;
    f[3] = fl;
}
SIS_HD Vec3 dih_force(const FFParams &P, Vec3 x1, Vec3 x2, Vec3 x3, Vec3 x4, int pos)
{
    Vec3 f[4];
    dih_force_vecs(P, x1, x2, x3, x4, f);
    return pos == 0 ? f[0] : pos == 1 ? f[1] : pos == 2 ? f[2] : f[3];
}

// energies: src/energy_bond.F90:17-28, energy_angl.F90:18-41, energy_angl_ReB.F90:18-29,
// energy_dih_cos.F90:18-39, energy_dih_exp.F90:20-41
SIS_HD double bond_energy(const FFParams &P, Vec3 x1, Vec3 x2)
{
    Vec3 v = pbc_vec_d(P, x1, x2);
    double d = sqrt(dot(v, v));
    return 0.5 * P.bond_k * (d - P.bond_r0) * (d - P.bond_r0);
}
SIS_HD double angl_energy(const FFParams &P, Vec3 x1, Vec3 x2, Vec3 x3)
{
    Vec3 v12 = pbc_vec_d(P, x1, x2), v32 = pbc_vec_d(P, x3, x2);
    double t = dot(v32, v12) / sqrt(dot(v12, v12) * dot(v32, v32));
    if (P.flg_angl_ReB) return 0.5 * P.angl_k * (t - P.angl_co0) * (t - P.angl_co0) / (1.0 - t * t);
    if (t > 1.0) t = 1.0;
    else if (t < -1.0) t = -1.0;
    t = acos(t);
    return 0.5 * P.angl_k * (t - P.angl_t0) * (t - P.angl_t0);
}
SIS_HD double dih_energy(const FFParams &P, Vec3 x1, Vec3 x2, Vec3 x3, Vec3 x4)
{
    Vec3 v12 = pbc_vec_d(P, x2, x1), v23 = pbc_vec_d(P, x3, x2), v34 = pbc_vec_d(P, x4, x3);
    Vec3 c1 = cross(v12, v23), c2 = cross(v23, v34);
    double dih = atan2(sqrt(dot(v23, v23)) * dot(v12, c2), dot(c1, c2));
    double e = 0.0;
    if (P.flg_dih_cos) e += P.dih_k * (1.0 + cos(dih + P.dih_p0));
    if (P.flg_dih_exp) e -= P.dih_k * exp((-0.5 * P.dih_w) * (dih - P.dih_p0) * (dih - P.dih_p0));
    return e;
}

// ---- pair terms: force on bead k from partner j, v = x_k - x_j (min image) ---------------
// src/force_wca.F90:34-66: F_k = coef * ((s2/d2)^7 - (s2/d2)^4) * v, d2 < sigma^2
SIS_HD double wca_scalar(const FFParams &P, double d2)
{
    double x = P.wca_sigma2 / d2;
    double x2 = x * x, x4 = x2 * x2;
    return P.wca_coef * (x4 * x2 * x - x4);
}
// src/force_ele_DH.F90:39-81: F_k = C/r^2 (1/r + 1/lambda) exp(-r/lambda) * v, d2 <= cutoff^2
SIS_HD double dh_scalar(double ele_coef, double inv_lambdaD, double d2)
{
#ifdef __CUDA_ARCH__
    const double rdist1 = rsqrt(d2);      // 1/r directly; r = d2 * (1/r)
#else
    const double rdist1 = 1.0 / sqrt(d2);
#endif
    const double dist1 = d2 * rdist1;
    return ele_coef * rdist1 * rdist1 * (rdist1 + inv_lambdaD) * exp(-dist1 * inv_lambdaD);
}
#ifdef __CUDACC__
// Straight-line (branch-free) versions for the pair loop of k_md_big, where four pairs are
// evaluated side by side and any branch would serialise them.
// 1 / sqrt(x) for normal x: MUFU seed + two Newton steps (error ~1 ulp of the result)
SIS_D double rsqrt_sl(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));      // MUFU.RSQ64H: relative error 2^-22
    const double h = 0.5 * x;
    double e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    return fma(y, e, y);
}
// exp(-x) for 0 <= x < 700: x = -(k ln2 + r), |r| <= ln2 / 2, degree-13 Taylor polynomial (truncation
// 4e-18), scaled by 2^k through the exponent field
SIS_D double exp_neg_sl(double x)
{
    const double t = -x;
    const double magic = 6755399441055744.0;                 // 1.5 * 2^52: round to nearest integer
    const double kz = fma(t, 1.4426950408889634, magic);
    const int k = __double2loint(kz);
    const double kf = kz - magic;
    double r = fma(kf, -6.93147180369123816490e-01, t);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;                       // 1 / 13!
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.388888888888889e-03);
    p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
#endif
// Mixed-precision mode (option "precision" = 1; tolerance 2e-5 of the largest force component, tests/test_gpu_mixed.py):
// the scalar of one listed pair, coef ((s2/d2)^7 - (s2/d2)^4) + C/r^2 (1/r + 1/lambda) exp(-r/lambda), in fp32 from the
// fp64 squared distance -- one MUFU.RSQ and one MUFU.EX2 instead of the fp64 Newton / Horner chains; the pair vector, the
// product and every accumulation stay fp64
SIS_HD double pair_scalar_f32(const FFParams &P, bool wca, bool ele, double d2, double ele_coef, double inv_lambdaD)
{
    const float d2f = (float)d2;
#ifdef __CUDA_ARCH__
    const float ri = rsqrtf(d2f);
#else
    const float ri = 1.0f / sqrtf(d2f);
#endif
    const float ri2 = ri * ri;
    float s = 0.0f;
    if (ele) {
        const float il = (float)inv_lambdaD;
        s = (float)ele_coef * ri2 * (ri + il) * expf(-(d2f * ri) * il);
    }
    if (wca) {
        const float x = (float)P.wca_sigma2 * ri2, x2 = x * x, x4 = x2 * x2;
        s += (float)P.wca_coef * (x4 * x2 * x - x4);
    }
    return (double)s;
}
// src/energy_wca.F90:19-35, src/energy_ele_DH.F90:20-40
SIS_HD double wca_energy(const FFParams &P, double d2)
{
    double d = sqrt(d2);
    if (d >= P.wca_sigma) return 0.0;
    d = P.wca_sigma / d;
    double d3 = d * d * d, d6 = d3 * d3;
    return P.wca_eps * (d6 * d6 - 2 * d6 + 1.0);
}
SIS_HD double dh_energy(const FFParams &P, const RepParams &R, double d2)
{
    double dist = sqrt(d2);
    if (dist > R.ele_cutoff) return 0.0;
    double rk = dist * R.inv_lambdaD;
    double e = R.ele_coef / dist * exp(-rk);
    if (P.temp_independent) e *= -(1.0 + 0.5 * rk) * R.diele_dTcoef;
    return e;
}

// ---- base pairs ---------------------------------------------------------------------------
// Force side of src/force_bp.F90:50-204 / force_bp_triplet.F90:62-240 / pass 1 of
// force_bp_limit_triplet.F90:85-257.  x[0..5] = beads i, j, i-1, j-1, i+1, j+1.
// Returns false when the distance gate |r12 - r0| > cutoff_ddist fails.  On success *u_out is
// the exponent and f[0..5] the (unscaled) d(u)/dr contributions on the six beads.
// acos arguments are NOT clamped here, exactly as in the reference's force routines.
SIS_HD bool bp_force_all(const FFParams &P, const BpPara &p, const Vec3 *x, double *u_out, Vec3 *f,
                         double *a12_out, Vec3 *v12_out)
{
    Vec3 v12 = pbc_vec_d(P, x[0], x[1]);
    double d1212 = dot(v12, v12), a12 = sqrt(d1212);
    double d = a12 - p.bond_r;
    *a12_out = a12;
    *v12_out = v12;
    if (fabs(d) > p.cutoff_ddist) return false;
    double u = p.bond_k * d * d;
#pragma unroll
    for (int s = 0; s < 6; s++) f[s] = mk(0.0, 0.0, 0.0);
    Vec3 fi = (2.0 * p.bond_k * d / a12) * v12, fk, fl;
    f[0] = fi;
    f[1] = -fi;
    Vec3 v13 = pbc_vec_d(P, x[0], x[2]), v15 = pbc_vec_d(P, x[0], x[4]);
    Vec3 v42 = pbc_vec_d(P, x[3], x[1]), v62 = pbc_vec_d(P, x[5], x[1]);
    double d1313 = dot(v13, v13), d4242 = dot(v42, v42), d1213 = dot(v12, v13);
    double d1242 = dot(v12, v42), d1215 = dot(v12, v15), d1515 = dot(v15, v15);
    double d1262 = dot(v12, v62), d6262 = dot(v62, v62);
    double r1213 = d1213 / d1212, r1242 = d1242 / d1212, r1215 = d1215 / d1212, r1262 = d1262 / d1212;
    double cosine, pre;
    // angle 3-1=2
    cosine = d1213 / (sqrt(d1313) * a12);
    d = acos(cosine) - p.angl_theta2;
    pre = 2.0 * p.angl_k2 * d / sqrt(d1313 * d1212 - d1213 * d1213);
    u += p.angl_k2 * d * d;
    fi = pre * (v12 - (d1213 / d1313) * v13);
    fk = pre * (v13 - r1213 * v12);
    f[0] += -fi - fk; f[1] += fk; f[2] = fi;
    // angle 1=2-4
    cosine = d1242 / (a12 * sqrt(d4242));
    d = acos(cosine) - p.angl_theta1;
    u += p.angl_k1 * d * d;
    pre = 2.0 * p.angl_k1 * d / sqrt(d1212 * d4242 - d1242 * d1242);
    fi = (-pre) * (v42 - r1242 * v12);
    fk = (-pre) * (v12 - (d1242 / d4242) * v42);
    f[0] += fi; f[1] += -fi - fk; f[3] = fk;
    // angle 5-1=2
    cosine = d1215 / (sqrt(d1515) * a12);
    d = acos(cosine) - p.angl_theta4;
    u += p.angl_k4 * d * d;
    pre = 2.0 * p.angl_k4 * d / sqrt(d1515 * d1212 - d1215 * d1215);
    fi = pre * (v12 - (d1215 / d1515) * v15);
    fk = pre * (v15 - r1215 * v12);
    f[0] += -fi - fk; f[1] += fk; f[4] = fi;
    // angle 1=2-6
    cosine = d1262 / (a12 * sqrt(d6262));
    d = acos(cosine) - p.angl_theta3;
    u += p.angl_k3 * d * d;
    pre = 2.0 * p.angl_k3 * d / sqrt(d1212 * d6262 - d1262 * d1262);
    fi = (-pre) * (v62 - r1262 * v12);
    fk = (-pre) * (v12 - (d1262 / d6262) * v62);
    f[0] += fi; f[1] += -fi - fk; f[5] = fk;
    // dihedral 4-2=1-3
    Vec3 m = cross(v42, v12), n = cross(v12, v13);
    double dih = atan2(dot(v42, n) * a12, dot(m, n));
    d = dih + p.dihd_phi1;
    u += p.dihd_k1 * (1.0 + cos(d));
    pre = -p.dihd_k1 * sin(d) * a12;
    fi = (pre / dot(m, m)) * m;
    fl = (-pre / dot(n, n)) * n;
    f[3] += fi;
    f[1] += (-1.0 + r1242) * fi - r1213 * fl;
    f[0] += (-1.0 + r1213) * fl - r1242 * fi;
    f[2] += fl;
    // dihedral 6-2=1-5
    m = cross(v62, v12);
    n = cross(v12, v15);
    dih = atan2(dot(v62, n) * a12, dot(m, n));
    d = dih + p.dihd_phi2;
    u += p.dihd_k2 * (1.0 + cos(d));
    pre = -p.dihd_k2 * sin(d) * a12;
    fi = (pre / dot(m, m)) * m;
    fl = (-pre / dot(n, n)) * n;
    f[5] += fi;
    f[1] += (-1.0 + r1262) * fi - r1215 * fl;
    f[0] += (-1.0 + r1215) * fl - r1262 * fi;
    f[4] += fl;
    *u_out = u;
    return true;
}

// The same force-side arithmetic split into seven independent ROLES so that the lanes of a
// group evaluate one base pair cooperatively (k_md_fused): role 0 = bond 1=2, roles 1..4 = the
// angles 3-1=2, 1=2-4, 5-1=2, 1=2-6, roles 5..6 = the dihedrals 4-2=1-3 and 6-2=1-5.
// A role touches beads 0 (i), 1 (j) and at most two more: bead bp_role_A(role) receives `ca`,
// bead bp_role_B(role) (dihedral roles only) receives `cb`.  Summing u and the contributions
// over the roles reproduces bp_force_all term by term (same expressions, other summation order).
struct BpSub {
    double u;
    Vec3 c0, c1, ca, cb;
};
SIS_HD int bp_role_A(int role) { return role == 1 ? 2 : role == 2 ? 3 : role == 3 ? 4 : role == 4 ? 5 : role == 5 ? 3 : role == 6 ? 5 : -1; }
SIS_HD int bp_role_B(int role) { return role == 5 ? 2 : role == 6 ? 4 : -1; }
// xa = position of bead bp_role_A(role), xb = position of bead bp_role_B(role) (dihedral roles)
SIS_HD BpSub bp_force_role(const FFParams &P, const BpPara &p, int role, Vec3 x0, Vec3 x1, Vec3 xa, Vec3 xb,
                           Vec3 v12, double d1212, double a12)
{
    BpSub o;
    o.u = 0.0;
    o.c0 = o.c1 = o.ca = o.cb = mk(0.0, 0.0, 0.0);
    if (role == 0) {
        const double d = a12 - p.bond_r;
        o.u = p.bond_k * d * d;
        const Vec3 fi = (2.0 * p.bond_k * d / a12) * v12;
        o.c0 = fi;
        o.c1 = -fi;
    } else if (role <= 4) {
        // angle between v12 and w; odd roles: w = x0 - xa (vertex i), even roles: w = xa - x1 (vertex j)
        const bool at_i = role & 1;
        const Vec3 w = at_i ? pbc_vec_d(P, x0, xa) : pbc_vec_d(P, xa, x1);
        const double k = role == 1 ? p.angl_k2 : role == 2 ? p.angl_k1 : role == 3 ? p.angl_k4 : p.angl_k3;
        const double th = role == 1 ? p.angl_theta2 : role == 2 ? p.angl_theta1 : role == 3 ? p.angl_theta4 : p.angl_theta3;
        const double dww = dot(w, w), d12w = dot(v12, w);
        const double cosine = at_i ? d12w / (sqrt(dww) * a12) : d12w / (a12 * sqrt(dww));
        const double d = acos(cosine) - th;
        o.u = k * d * d;
        const double pre = 2.0 * k * d / sqrt(at_i ? dww * d1212 - d12w * d12w : d1212 * dww - d12w * d12w);
        const double r12w = d12w / d1212;
        if (at_i) {
            const Vec3 fi = pre * (v12 - (d12w / dww) * w);
            const Vec3 fk = pre * (w - r12w * v12);
            o.c0 = -fi - fk; o.c1 = fk; o.ca = fi;
        } else {
            const Vec3 fi = (-pre) * (w - r12w * v12);
            const Vec3 fk = (-pre) * (v12 - (d12w / dww) * w);
            o.c0 = fi; o.c1 = -fi - fk; o.ca = fk;
        }
    } else if (role <= 6) {
        // dihedral a-2=1-b: a on the j side (bead 3 or 5), b on the i side (bead 2 or 4)
        const Vec3 va2 = pbc_vec_d(P, xa, x1), v1b = pbc_vec_d(P, x0, xb);
        const double ra = dot(v12, va2) / d1212, rb = dot(v12, v1b) / d1212;
        const Vec3 m = cross(va2, v12), n = cross(v12, v1b);
        // phi = atan2(Y, X): cos(phi + phi0) and sin(phi + phi0) from (X, Y) / |(X, Y)| and the
        // tabulated cos / sin of phi0 (src/force_bp.F90:143-160 evaluates atan2, then cos and sin)
        const double Y = dot(va2, n) * a12, X = dot(m, n);
        const double kk = role == 5 ? p.dihd_k1 : p.dihd_k2;
        const double c0 = role == 5 ? p.cos_phi1 : p.cos_phi2, s0 = role == 5 ? p.sin_phi1 : p.sin_phi2;
#ifdef __CUDA_ARCH__
        const double rinv = rsqrt(X * X + Y * Y);
#else
        const double rinv = 1.0 / sqrt(X * X + Y * Y);
#endif
        const double cosd = (X * c0 - Y * s0) * rinv, sind = (Y * c0 + X * s0) * rinv;
        o.u = kk * (1.0 + cosd);
        const double pre = -kk * sind * a12;
        const Vec3 fi = (pre / dot(m, m)) * m;
        const Vec3 fl = (-pre / dot(n, n)) * n;
        o.ca = fi;
        o.c1 = (-1.0 + ra) * fi - rb * fl;
        o.c0 = (-1.0 + rb) * fl - ra * fi;
        o.cb = fl;
    }
    return o;
}

// Role-major form of the same arithmetic (k_md_rep, k_md_fast): ONE role of one base pair per thread, so that a warp
// runs one code path; the results of the seven roles go to BP_ROLE doubles per pair:
//   angle role q = 1..4 at 7 (q - 1): {u, fi, fk};  dihedral role q = 5, 6 at 28 + 9 (q - 5): {u, fi, fl, ra, rb};
//   bond role 0 at 46: {u, fi}.
// xa / xb as in bp_force_role.  The angles go through one reciprocal square root each (cos = v12.w rsqrt(|v12|^2 |w|^2)).
constexpr int BP_ROLE = 50;
SIS_HD double bp_rsqrt(double x)
{
#ifdef __CUDA_ARCH__
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
SIS_HD void bp_role_eval(const FFParams &P, const BpPara &p, int role, Vec3 x0, Vec3 x1, Vec3 xa, Vec3 xb, double *out)
{
    const Vec3 v12 = pbc_vec_d(P, x0, x1);
    const double d1212 = dot(v12, v12);
    if (role == 0) {
        const double a12 = sqrt(d1212), d = a12 - p.bond_r;
        const Vec3 fi = (2.0 * p.bond_k * d / a12) * v12;
        out[46] = p.bond_k * d * d;
        out[47] = fi.x; out[48] = fi.y; out[49] = fi.z;
    } else if (role <= 4) {
        const bool at_i = role & 1;
        const Vec3 w = at_i ? pbc_vec_d(P, x0, xa) : pbc_vec_d(P, xa, x1);
        const double k = role == 1 ? p.angl_k2 : role == 2 ? p.angl_k1 : role == 3 ? p.angl_k4 : p.angl_k3;
        const double th = role == 1 ? p.angl_theta2 : role == 2 ? p.angl_theta1 : role == 3 ? p.angl_theta4 : p.angl_theta3;
        const double dww = dot(w, w), d12w = dot(v12, w), dd = dww * d1212;
        const double cosine = d12w * bp_rsqrt(dd);
        const double d = acos(cosine) - th;
        const double pre = 2.0 * k * d * bp_rsqrt(dd - d12w * d12w);
        const double r12w = d12w / d1212, rww = d12w / dww;
        Vec3 fi, fk;
        if (at_i) { fi = pre * (v12 - rww * w); fk = pre * (w - r12w * v12); }
        else { fi = (-pre) * (w - r12w * v12); fk = (-pre) * (v12 - rww * w); }
        out += 7 * (role - 1);
        out[0] = k * d * d;
        out[1] = fi.x; out[2] = fi.y; out[3] = fi.z; out[4] = fk.x; out[5] = fk.y; out[6] = fk.z;
    } else {
        const Vec3 va2 = pbc_vec_d(P, xa, x1), v1b = pbc_vec_d(P, x0, xb);
        const double ra = dot(v12, va2) / d1212, rb = dot(v12, v1b) / d1212;
        const Vec3 m = cross(va2, v12), n = cross(v12, v1b);
        const double a12 = sqrt(d1212);
        const double Y = dot(va2, n) * a12, X = dot(m, n);
        const double kk = role == 5 ? p.dihd_k1 : p.dihd_k2;
        const double c0 = role == 5 ? p.cos_phi1 : p.cos_phi2, s0 = role == 5 ? p.sin_phi1 : p.sin_phi2;
        const double rinv = bp_rsqrt(X * X + Y * Y);
        const double cosd = (X * c0 - Y * s0) * rinv, sind = (Y * c0 + X * s0) * rinv;
        const double pre = -kk * sind * a12;
        const Vec3 fi = (pre / dot(m, m)) * m;
        const Vec3 fl = (-pre / dot(n, n)) * n;
        out += 28 + 9 * (role - 5);
        out[0] = kk * (1.0 + cosd);
        out[1] = fi.x; out[2] = fi.y; out[3] = fi.z; out[4] = fl.x; out[5] = fl.y; out[6] = fl.z;
        out[7] = ra; out[8] = rb;
    }
}
// exponent of the pair from its roles
SIS_HD double bp_roles_u(const double *r) { return r[46] + ((r[0] + r[7]) + (r[14] + r[21])) + (r[28] + r[37]); }
// unscaled force on bead t (0..5 = i, j, i-1, j-1, i+1, j+1) from the roles:
// angle roles 1, 3 (vertex i): c0 = -fi - fk, c1 = fk, ca = fi; roles 2, 4 (vertex j): c0 = fi, c1 = -fi - fk, ca = fk;
// dihedral roles 5, 6: ca = fi, cb = fl, c1 = (ra - 1) fi - rb fl, c0 = (rb - 1) fl - ra fi; bond: +fi on i, -fi on j
SIS_HD Vec3 bp_roles_bead_force(const double *r, int t)
{
    auto V = [&](int o) { return mk(r[o], r[o + 1], r[o + 2]); };
    if (t < 2) {
        const Vec3 f0 = V(47);
        const Vec3 di5 = V(29), dl5 = V(32), di6 = V(38), dl6 = V(41);
        const double ra5 = r[35], rb5 = r[36], ra6 = r[44], rb6 = r[45];
        if (t == 0) return f0 - (V(1) + V(4)) + V(8) - (V(15) + V(18)) + V(22) + ((rb5 - 1.0) * dl5 - ra5 * di5) + ((rb6 - 1.0) * dl6 - ra6 * di6);
        return -f0 + V(4) - (V(8) + V(11)) + V(18) - (V(22) + V(25)) + ((ra5 - 1.0) * di5 - rb5 * dl5) + ((ra6 - 1.0) * di6 - rb6 * dl6);
    }
    if (t == 2) return V(1) + V(32);
    if (t == 3) return V(11) + V(29);
    if (t == 4) return V(15) + V(41);
    return V(25) + V(38);
}

// Energy side: src/energy_bp.F90:40-62,76-118 (clamped acos, mp_dihedral)
SIS_HD double mp_angle(const FFParams &P, Vec3 x1, Vec3 x2, Vec3 x3)
{
    Vec3 v12 = pbc_vec_d(P, x1, x2), v32 = pbc_vec_d(P, x3, x2);
    double co = dot(v32, v12) / sqrt(dot(v12, v12) * dot(v32, v32));
    if (co > 1.0) co = 1.0;
    else if (co < -1.0) co = -1.0;
    return acos(co);
}
SIS_HD double mp_dihedral(const FFParams &P, Vec3 x1, Vec3 x2, Vec3 x3, Vec3 x4)
{
    Vec3 v12 = pbc_vec_d(P, x1, x2), v32 = pbc_vec_d(P, x3, x2), v34 = pbc_vec_d(P, x3, x4);
    Vec3 m = cross(v12, v32), n = cross(v32, v34);
    return atan2(dot(v12, n) * sqrt(dot(v32, v32)), dot(m, n));
}
SIS_HD bool bp_energy_u(const FFParams &P, const BpPara &p, const Vec3 *x, double *u_out)
{
    Vec3 v = pbc_vec_d(P, x[0], x[1]);
    double d = sqrt(dot(v, v)) - p.bond_r;
    if (fabs(d) > p.cutoff_ddist) return false;
    double u = p.bond_k * d * d, t;
    t = mp_angle(P, x[0], x[1], x[3]); u += p.angl_k1 * (t - p.angl_theta1) * (t - p.angl_theta1);
    t = mp_angle(P, x[2], x[0], x[1]); u += p.angl_k2 * (t - p.angl_theta2) * (t - p.angl_theta2);
    t = mp_angle(P, x[0], x[1], x[5]); u += p.angl_k3 * (t - p.angl_theta3) * (t - p.angl_theta3);
    t = mp_angle(P, x[4], x[0], x[1]); u += p.angl_k4 * (t - p.angl_theta4) * (t - p.angl_theta4);
    t = mp_dihedral(P, x[2], x[0], x[1], x[3]); u += p.dihd_k1 * (1.0 + cos(t + p.dihd_phi1));
    t = mp_dihedral(P, x[4], x[0], x[1], x[5]); u += p.dihd_k2 * (1.0 + cos(t + p.dihd_phi2));
    *u_out = u;
    return true;
}

// ---- sequence rules (src/const_idx.F90:343-369, src/init_bp.F90:184-313,352-378) ----------
SIS_HD bool is_complement(int s1, int s2)
{
    if (s1 == SEQ_A) return s2 == SEQ_U;
    if (s1 == SEQ_U) return s2 == SEQ_A || s2 == SEQ_G;
    if (s1 == SEQ_G) return s2 == SEQ_C || s2 == SEQ_U;
    if (s1 == SEQ_C) return s2 == SEQ_G;
    return false;
}
SIS_HD int bp3_hash_key(int w, int x, int u, int z, int y, int v)
{
    int a, b, c, d, e, f;
    if (x < y) { a = w; b = x; c = u; d = z; e = y; f = v; }
    else { f = w; e = x; d = u; c = z; b = y; a = v; }
    return a + b * 4 + c * 16 + d * 64 + e * 256 + f * 1024;  // 0-based key
}
SIS_HD int bp_type_of(int a, int b)
{
    if (a == SEQ_G && b == SEQ_C) return 1;
    if (a == SEQ_A && b == SEQ_U) return 2;
    if (a == SEQ_G && b == SEQ_U) return 3;
    if (a == SEQ_C && b == SEQ_G) return 4;
    if (a == SEQ_U && b == SEQ_A) return 5;
    if (a == SEQ_U && b == SEQ_G) return 6;
    return 0;
}

// chain neighbour of a bead described by meta (start, len, local, flags): bead at local+off,
// wrapping for circRNA; -1 if it does not exist on a linear chain
SIS_HD int chain_nb(int4 m, int off)
{
    int l = m.z + off;
    if (m.w & MF_CIRC) {
        l %= m.y;
        if (l < 0) l += m.y;
        return m.x + l;
    }
    return (l >= 0 && l < m.y) ? m.x + l : -1;
}

// bp_map / bp3_map of src/init_bp.F90:184-313 evaluated on the fly for beads imp < jmp.
// seq = per-bead sequence codes.  Returns bp type (1..6) in *bptype and the 1-based bp3 index
// (model 5) or the type (other models); 0 = pair not allowed.
// secondary-structure restriction (src/read_ss.F90, src/init_bp.F90:178-209): sorted keys
// imp * nmp + jmp (0-based, imp < jmp) of the pairs read_ss leaves in bp_map; n_ss < 0: all pairwise
struct BpRestrict {
    const unsigned long long *keys;
    int n_ss, nmp;
};
SIS_HD bool bp_restrict_has(const BpRestrict &Rs, int imp, int jmp)
{
    const unsigned long long key = (unsigned long long)imp * (unsigned long long)Rs.nmp + (unsigned long long)jmp;
    int lo = 0, hi = Rs.n_ss;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (Rs.keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo < Rs.n_ss && Rs.keys[lo] == key;
}
SIS_HD int bp_lookup(int model, int min_loop, const int *__restrict__ hash, const int *__restrict__ seq,
                     int imp, int4 mi, int jmp, int4 mj, int *bptype);
// bp_lookup under a restriction: only listed pairs; for bp_model /= 5, bp3_map = bp_map
// (src/init_bp.F90:315-317) keeps every listed pair, whatever the rules of :211-266 say
SIS_HD int bp_lookup_r(const BpRestrict &Rs, int model, int min_loop, const int *__restrict__ hash,
                       const int *__restrict__ seq, int imp, int4 mi, int jmp, int4 mj, int *bptype)
{
    if (Rs.n_ss >= 0) {
        *bptype = 0;
        if (!bp_restrict_has(Rs, imp, jmp)) return 0;
        if (model != 5) {
            const int t = bp_type_of(seq[mi.x + mi.z], seq[mj.x + mj.z]);
            *bptype = t;
            return t;
        }
    }
    return bp_lookup(model, min_loop, hash, seq, imp, mi, jmp, mj, bptype);
}
SIS_HD int bp_lookup(int model, int min_loop, const int *__restrict__ hash, const int *__restrict__ seq,
                     int imp, int4 mi, int jmp, int4 mj, int *bptype)
{
    *bptype = 0;
    int i = mi.z + 1, j = mj.z + 1, ni = mi.y, nj = mj.y;  // 1-based local as in the reference
    bool i_circ = mi.w & MF_CIRC, j_circ = mj.w & MF_CIRC;
    int i_pre = i - 1, i_nxt = i + 1, j_pre = j - 1, j_nxt = j + 1;
    if (i_circ) {
        if (i == 1) i_pre = ni;
        if (i == ni) i_nxt = 1;
    } else if (i == 1 || i == ni) return 0;
    if (j_circ) {
        if (j == 1) j_pre = nj;
        if (j == nj) j_nxt = 1;
    } else if (j == 1 || j == nj) return 0;
    const int *si = seq + mi.x - 1, *sj = seq + mj.x - 1;  // si[1..ni]
    bool p_ipre_jnxt = is_complement(si[i_pre], sj[j_nxt]);
    bool p_inxt_jpre = is_complement(si[i_nxt], sj[j_pre]);
    if (!i_circ && i_pre == 1) p_ipre_jnxt = false;
    if (!j_circ && j_nxt == nj) p_ipre_jnxt = false;
    if (!i_circ && i_nxt == ni) p_inxt_jpre = false;
    if (!j_circ && j_pre == 1) p_inxt_jpre = false;
    if (mi.x == mj.x) {  // same chain
        if (i + min_loop >= j) return 0;
        if (i + 1 + min_loop >= j - 1) p_inxt_jpre = false;
        if (i_circ) {
            if (j + min_loop >= i + ni) return 0;
            if (j + 1 + min_loop >= i - 1 + ni) p_ipre_jnxt = false;
        }
    }
    if (model == 3 || model == 4 || model == 5)
        if (!p_ipre_jnxt && !p_inxt_jpre) return 0;
    int a = si[i], b = sj[j];
    int t = bp_type_of(a, b);
    *bptype = t;
    if (model != 5) return t;
    if (a > 3 || b > 3) return 0;
    int key;
    if (!p_ipre_jnxt) key = bp3_hash_key(SEQ_U, a, si[i_nxt], SEQ_U, b, sj[j_pre]);
    else if (!p_inxt_jpre) key = bp3_hash_key(si[i_pre], a, SEQ_U, sj[j_nxt], b, SEQ_U);
    else key = bp3_hash_key(si[i_pre], a, si[i_nxt], sj[j_nxt], b, sj[j_pre]);
    (void)imp; (void)jmp;
    return (key >= 0 && key < 4096) ? hash[key] : 0;
}

// ---- counter-based Gaussian generator (replaces mt_stream + rnd_boxmuller) -----------------
// Philox4x32-10, key = (seed_lo, seed_hi + stream), counter = (step_lo, step_hi, bead, draw).
SIS_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
SIS_HD double u53(uint32_t lo, uint32_t hi)
{
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
// component c (0..2) of the three normal deviates of (stream, step, bead)
SIS_HD double gauss_component(uint64_t seed, uint32_t stream, uint64_t step, uint32_t bead, int c)
{
    uint32_t o4[4];
    philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), bead, c == 2 ? 1u : 0u, (uint32_t)seed,
                  (uint32_t)(seed >> 32) + stream, o4);
    double u1 = u53(o4[0], o4[1]), u2 = u53(o4[2], o4[3]);
    double rad = sqrt(-2.0 * log(u1));
    double ang = 2.0 * SIS_PI * u2;
    return rad * (c == 1 ? sin(ang) : cos(ang));
}
SIS_HD double uniform_counter(uint64_t seed, uint32_t stream, uint64_t counter)
{
    uint32_t o4[4];
    philox4x32_10((uint32_t)counter, (uint32_t)(counter >> 32), 0u, 2u, (uint32_t)seed,
                  (uint32_t)(seed >> 32) + stream, o4);
    return u53(o4[0], o4[1]);
}
