// sis_fast.cuh -- k_md_fast: the MD loop body (src/job_md.F90:374-460) for systems with at most
// one bead per 8-lane group of the grid (R * nmp <= 64 * #SM: BASELINE configs C1..C4).
//
// Same warp-specialised persistent structure as k_md_fused (sis_fused.cuh) with everything that
// is constant between two list rebuilds held on chip for the whole launch:
//   MAIN  (16 warps, 64 groups x 8 lanes): group g owns one bead.  Its neighbour row lives in
//         shared memory, its position / velocity / beta / displacement / coefficients in
//         registers.  Per step: pair loop (src/force_wca.F90, src/force_ele_DH.F90), then the
//         integrator (src/job_md.F90:409-456) as soon as the other force parts are complete.
//   BP    (2 warps): candidate data (six bead ids, type, dG) cached in shared memory; one
//         thread per candidate applies the distance gate, 8-lane groups evaluate the survivors
//         (one role per lane, bp_force_role) and scatter with fp64 atomics.
//   LOC   (6 warps): every bonded term the CTA's beads take part in is evaluated ONCE
//         (src/force_bond.F90, force_angl*.F90, force_dih_*.F90) into shared tables the MAIN
//         groups read; then the Gaussians of the NEXT step (Philox + Box-Muller) into a
//         double-buffered shared array; one LOC warp also evaluates the rebuild test.
// Grid barriers: bar[0] "bp scatter done" (arrive: BP warps, wait: MAIN after the pair loop) and
// bar[32] "positions written" (the one exposed barrier per step).  The rebuild test of
// src/job_md.F90:376-388 for R == 1 goes through per-CTA (max, 2nd max) slots read after the
// position barrier, off the critical path; with R > 1 through the atomic top-2 of k_md_fused.
#pragma once

#include "sis_fused.cuh"

constexpr int FAST_MAIN_WARPS = 16, FAST_BP_WARPS = 2, FAST_LOC_WARPS = 6;
constexpr int FAST_MAIN = 32 * FAST_MAIN_WARPS, FAST_BP = 32 * FAST_BP_WARPS, FAST_LOC = 32 * FAST_LOC_WARPS;
constexpr int FAST_THREADS = FAST_MAIN + FAST_BP + FAST_LOC;
constexpr int FAST_GROUPS = FAST_MAIN / 8;        // beads per CTA
constexpr int FAST_BPGROUPS = FAST_BP / 8;
constexpr int FAST_NT = FAST_GROUPS + 3;          // term-table slots: starts at item0-3 .. item0+63
constexpr int FAST_BPC = 256;                     // cached bp candidates per CTA
constexpr int FAST_MAX_CTA = 1024;                // capacity of the top-2 slot array

struct FastBpCand {
    int b[6];       // beads i, j, i-1, j-1, i+1, j+1
    int ibp, type, rep, last;   // last = bp_status last written (-1 unknown)
    double dG;
};

struct FastShared {
    Vec3 bondf[FAST_NT];        // bond starting at slot: F(first) = -f, F(second) = +f
    Vec3 angf[FAST_NT][2];      // angle starting at slot: F1, F3; F2 = -(F1 + F3)
    Vec3 dihf[FAST_NT][4];      // dihedral starting at slot: F1..F4
    double gauss[2][FAST_GROUPS][3];
    double d2[FAST_GROUPS];
    FastBpCand cand[FAST_BPC];
    int active[FAST_BPC];
    int nactive, ncand;
    int blocked[2];             // rebuild verdict, indexed by iteration parity
    double top[2];
};

SIS_D void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// three normal deviates of (stream, step, bead): same values as gauss_component (one Philox block
// gives x and y, a second one z)
SIS_D void gauss_pair(uint64_t seed, uint32_t stream, uint64_t step, uint32_t bead, uint32_t draw, double *g0, double *g1)
{
    uint32_t o4[4];
    philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), bead, draw, (uint32_t)seed, (uint32_t)(seed >> 32) + stream, o4);
    const double u1 = u53(o4[0], o4[1]), u2 = u53(o4[2], o4[3]);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincos(2.0 * SIS_PI * u2, &sn, &cs);
    *g0 = rad * cs;
    *g1 = rad * sn;
}

// pair loop of one bead over its row in shared memory: U gathers in flight per lane
template <int U>
SIS_D Vec3 fast_pair(const FFParams &P, const uint32_t *row, int nn, int lane, Vec3 xk, const double4 *__restrict__ pos,
                     double ele_coef, double inv_lambdaD, double cut2)
{
    Vec3 F = mk(0.0, 0.0, 0.0);
    for (int e0 = lane; e0 < nn; e0 += U * 8) {
        uint32_t ent[U];
        Vec3 xj[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int e = e0 + u * 8;
            ent[u] = e < nn ? row[e] : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (ent[u]) xj[u] = v3(pos[ent[u] & NB_IDX]);
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!ent[u]) continue;
            Vec3 v = xk - xj[u];
            if (ent[u] & NB_IMG) {
                v.x = minimg1(v.x, P.box[0], P.box_half[0]);
                v.y = minimg1(v.y, P.box[1], P.box_half[1]);
                v.z = minimg1(v.z, P.box[2], P.box_half[2]);
            }
            const double d2 = dot(v, v);
            double s = 0.0;
            if ((ent[u] & NB_WCA) && d2 < P.wca_sigma2) s += wca_scalar(P, d2);
            if ((ent[u] & NB_ELE) && d2 <= cut2) s += dh_scalar(ele_coef, inv_lambdaD, d2);
            F += s * v;
        }
    }
    return F;
}

// (rep, bead) of a flattened item
SIS_D void fast_item(int item, int nmp, int *rep, int *k) { *rep = item / nmp; *k = item - *rep * nmp; }

// closure terms of a circular chain (src/list_local.F90:59-104) that wrap past the chain end are
// not in the CTA's term tables (their start bead is not item - p): gather form, rare
SIS_D double fast_closure_terms(const DevSys &S, const FFParams &P, const double4 *pos, int4 m, int k, int c)
{
    double f = 0.0;
    const Vec3 xk = v3(pos[k]);
    const bool dih_on = P.flg_dih_cos || P.flg_dih_exp;
    for (int kind = 0; kind < 3; kind++) {
        if (kind == 2 && !dih_on) break;
        const int span = kind + 1;
        if (m.y <= span) continue;
        for (int p = 0; p <= span; p++) {
            if (m.z - p >= 0) continue;    // regular: from the tables
            Vec3 x[4];
            for (int q = 0; q <= span; q++) x[q] = (q == p) ? xk : v3(pos[chain_nb(m, q - p)]);
            Vec3 t;
            if (kind == 0) t = bond_force(P, x[0], x[1], p);
            else if (kind == 1) t = P.flg_angl_ReB ? angl_ReB_force(P, x[0], x[1], x[2], p) : angl_force(P, x[0], x[1], x[2], p);
            else t = dih_force(P, x[0], x[1], x[2], x[3], p);
            f += c == 0 ? t.x : c == 1 ? t.y : t.z;
        }
    }
    return f;
}

// common per-launch context of the three roles
struct FastCtx {
    int R, nmp, K, nitems, ncta, item0, nsteps;
    unsigned long long step0;
    int cur0;
    unsigned *bar_bp, *bar_pos;
    bool dih_on;
};
SIS_D void cta_sync() { named_bar_sync(0, FAST_THREADS); }

// Loop protocol shared by the roles (iteration n, parity q = n & 1):
//   LOC   : [n > 0: verdict -> sh.blocked[q]]  tables  arrive(3)  gauss  cta_sync  read blocked[q]
//   MAIN  : pair  sync(3)  read blocked[q] (leave)  wait bar_bp  integrate  arrive+wait bar_pos  cta_sync
//   BP    : gate  eval  arrive bar_bp  cta_sync  read blocked[q]
// A MAIN thread that leaves at barrier 3 performs the iteration's cta_sync before returning, so
// every role executes the same number of cta_syncs.  Returns the number of completed steps.

SIS_D int fast_main_role(const DevSys &S, const FFParams &P, const StepArgs &A, const FastCtx &C, FastShared &sh,
                         uint32_t *s_rows)
{
    (void)A;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 7, grp = tid >> 3;
    const int nmp = C.nmp, K = C.K, R = C.R;
    unsigned long long step = C.step0;
    int cur = C.cur0;
    const bool prof = S.prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;

    // ---- state of the owned bead (registers for the whole launch) -----------------------------
    int rep = 0, k = 0, nn = 0;
    const int item = C.item0 + grp;
    const bool valid = item < C.nitems;
    int4 meta = make_int4(0, 1, 0, 0);
    Vec3 xk = mk(0.0, 0.0, 0.0);
    double xc = 0.0, vc = 0.0, ac = 0.0, mc = 0.0;     // component `lane` (< 3) of x, v, beta, xyz_move
    double ca = 0.0, cb = 0.0, cc = 0.0;               // a, sqrt(b)h/m, sqrt(b)h of set_md_coef
    double ele_coef = 0.0, inv_lambdaD = 0.0, cut2 = -1.0, noise = 0.0;
    size_t o3 = 0;
    if (valid) {
        fast_item(item, nmp, &rep, &k);
        meta = S.meta[k];
        nn = S.nnbr[(size_t)rep * nmp + k];
        const uint32_t *row = S.nbr + ((size_t)rep * nmp + k) * K;
        for (int e = lane; e < nn; e += 8) s_rows[grp * K + e] = row[e];
        const double4 p = S.pos[cur][(size_t)rep * nmp + k];
        xk = v3(p);
        const RepParams &Rp = S.rp[rep];
        ele_coef = Rp.ele_coef; inv_lambdaD = Rp.inv_lambdaD;
        if (P.flg_ele) cut2 = Rp.ele_cutoff * Rp.ele_cutoff;
        const double4 coef = S.coef[k];
        ca = coef.x; cb = coef.y; cc = coef.z;
        noise = coef.w * Rp.sqrt_tempK;
        if (lane < 3) {
            o3 = ((size_t)rep * nmp + k) * 3 + lane;
            xc = lane == 0 ? p.x : lane == 1 ? p.y : p.z;
            vc = S.vel[o3]; ac = S.acc[o3]; mc = S.move[o3];
        }
    }
    cta_sync();                                        // prologue of all roles complete
    const bool overflow = sh.ncand > FAST_BPC;
    int n = 0;
    if (!sh.blocked[0] && !overflow)
    for (; n < C.nsteps; n++) {
        if (prof) c0 = clock64();
        const double4 *pos = S.pos[cur] + (size_t)rep * nmp;
        Vec3 F = fast_pair<2>(P, s_rows + grp * K, nn, lane, xk, pos, ele_coef, inv_lambdaD, cut2);
        F.x = group_sum<8>(F.x); F.y = group_sum<8>(F.y); F.z = group_sum<8>(F.z);
        double f = lane == 0 ? F.x : lane == 1 ? F.y : F.z;
        if (prof) c1 = clock64();
        named_bar_sync(3, FAST_MAIN + FAST_LOC);       // bonded tables + rebuild verdict
        if (sh.blocked[n & 1]) { cta_sync(); break; }
        if (valid && lane < 3) {
            const int l = meta.z, len = meta.y, s0 = grp + 3;
            const bool circ = meta.w & MF_CIRC;
            if (circ ? len > 1 : l + 1 < len) f -= (&sh.bondf[s0].x)[lane];
            if (l >= 1) f += (&sh.bondf[s0 - 1].x)[lane];
            if (circ ? len > 2 : l + 2 < len) f += (&sh.angf[s0][0].x)[lane];
            if (l >= 1 && (circ ? len > 2 : l + 1 < len)) f -= (&sh.angf[s0 - 1][0].x)[lane] + (&sh.angf[s0 - 1][1].x)[lane];
            if (l >= 2) f += (&sh.angf[s0 - 2][1].x)[lane];
            if (C.dih_on) {
                if (circ ? len > 3 : l + 3 < len) f += (&sh.dihf[s0][0].x)[lane];
                if (l >= 1 && (circ ? len > 3 : l + 2 < len)) f += (&sh.dihf[s0 - 1][1].x)[lane];
                if (l >= 2 && (circ ? len > 3 : l + 1 < len)) f += (&sh.dihf[s0 - 2][2].x)[lane];
                if (l >= 3) f += (&sh.dihf[s0 - 3][3].x)[lane];
            }
            if (circ && l < 3) f += fast_closure_terms(S, P, pos, meta, k, lane);
        }
        if (tid == 0) {                                // base-pair scatter of every CTA complete?
            const unsigned target = (unsigned)(n + 1) * (unsigned)C.ncta;
            while (ld_acquire_u32(C.bar_bp) < target) { }
        }
        named_bar_sync(2, FAST_MAIN);
        if (prof) c2 = clock64();
        // ---- integrate (src/job_md.F90:415-456) ----------------------------------------------
        double xnew = 0.0;
        if (valid && lane < 3) {
            f += __ldcg(S.f_bp + o3);
            S.f_bp[o3] = 0.0;
            const double g = sh.gauss[(step + 1ull) & 1][grp][lane];
            double dx;
            if (P.integrator == 1) {
                const double accels_pre = noise * g;
                vc = ca * vc + cb * f + (ac + accels_pre);
                ac = accels_pre;
                dx = cc * vc;
            } else {
                dx = ca * f + noise * g;
            }
            xc += dx;
            mc += dx;
            xnew = xc;
            if (f > 100.0) S.misc->max_force_flag |= 1;
        }
        xk.x = __shfl_sync(0xffffffffu, xnew, 0, 8);
        xk.y = __shfl_sync(0xffffffffu, xnew, 1, 8);
        xk.z = __shfl_sync(0xffffffffu, xnew, 2, 8);
        const double s0 = __shfl_sync(0xffffffffu, mc, 0, 8);
        const double s1 = __shfl_sync(0xffffffffu, mc, 1, 8);
        const double s2 = __shfl_sync(0xffffffffu, mc, 2, 8);
        const double d2 = valid ? s0 * s0 + s1 * s1 + s2 * s2 : 0.0;
        if (lane == 0) {
            if (valid) S.pos[cur ^ 1][(size_t)rep * nmp + k] = make_double4(xk.x, xk.y, xk.z, 0.0);
            if (R == 1) sh.d2[grp] = d2;
            else if (d2 > 0.0) top2_push(S.top2 + ((size_t)rep * 2 + ((step + 1ull) & 1ull)) * 2, d2, 0.0);
        }
        named_bar_sync(2, FAST_MAIN);
        if (warp == 0) {
            if (R == 1) {                              // CTA top-2 of |xyz_move|^2 -> this CTA's slot
                double t1 = sh.d2[tid], t2 = sh.d2[tid + 32];
                if (t2 > t1) { const double t = t1; t1 = t2; t2 = t; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double b1 = __shfl_xor_sync(0xffffffffu, t1, o), b2 = __shfl_xor_sync(0xffffffffu, t2, o);
                    top2_merge(t1, t2, b1, b2);
                }
                if (tid == 0) S.top2_slots[((step + 1ull) & 1ull) * FAST_MAX_CTA + blockIdx.x] = make_double2(t1, t2);
            } else if (tid == 0 && blockIdx.x == 0) {
                for (int r = 0; r < R; r++) {          // protocol of k_md_fused: clear the slots of step + 2
                    unsigned long long *t = S.top2 + ((size_t)r * 2 + ((step + 2ull) & 1ull)) * 2;
                    t[0] = 0ull; t[1] = 0ull;
                }
            }
            if (tid == 0) {
                if (prof) c3 = clock64();
                red_release_add(C.bar_pos, 1u);
                const unsigned target = (unsigned)(n + 1) * (unsigned)C.ncta;
                while (ld_acquire_u32(C.bar_pos) < target) { }
                if (prof) {
                    c4 = clock64();
                    S.prof[0] += (unsigned long long)(c1 - c0); S.prof[1] += (unsigned long long)(c2 - c1);
                    S.prof[2] += (unsigned long long)(c3 - c2); S.prof[3] += (unsigned long long)(c4 - c3);
                    S.prof[4] += 1ull;
                }
            }
        }
        step++;
        cur ^= 1;
        cta_sync();
    }
    cta_sync();                                        // everybody is out of the loop
    if (valid && lane < 3) { S.vel[o3] = vc; S.acc[o3] = ac; S.move[o3] = mc; }
    return n;
}

SIS_D void fast_bp_role(const DevSys &S, const FFParams &P, const FastCtx &C, FastShared &sh)
{
    const int tid = threadIdx.x, bt = tid - FAST_MAIN;
    const int nmp = C.nmp, R = C.R, KB = S.KB;
    int cur = C.cur0;
    // ---- candidate cache: the CTA's share of the flattened (replica, half-list index) space -----
    if (bt < 32) {
        int total = 0;
        for (int r0 = 0; r0 < R; r0 += 32) total += (r0 + bt < R) ? S.nbp_half[r0 + bt] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        const int lo = (int)((long long)total * blockIdx.x / C.ncta);
        const int hi = (int)((long long)total * (blockIdx.x + 1) / C.ncta);
        if (bt == 0) { sh.ncand = 0; sh.nactive = 0; }
        __syncwarp();
        int base = 0;
        for (int r = 0; r < R; r++) {
            const int nb = S.nbp_half[r];
            const int a = max(lo, base), b = min(hi, base + nb);
            for (int g = a + bt; g < b; g += 32) {
                const int ibp = S.bp_half[(size_t)r * S.bp_half_cap + (g - base)];
                const size_t rbase = (size_t)r * nmp, pbase = rbase * KB;
                const int ki = ibp / KB, e = ibp - ki * KB;
                const int j = S.bprow[(rbase + ki) * KB + e].x;
                double dG = 1.0;
                if (P.bp_triplet) {
                    const double2 cf = S.bp_coef[pbase + ibp];
                    dG = cf.x - S.rp[r].tempK * cf.y;           // src/force_bp_triplet.F90:77
                    if (dG >= 0.0) {                            // never active at this temperature
                        S.bp_status[pbase + ibp] = 0;
                        S.ene_bp[pbase + ibp] = 0.0;
                        continue;
                    }
                }
                const int slot = atomicAdd(&sh.ncand, 1);
                if (slot < FAST_BPC) {
                    FastBpCand &cd = sh.cand[slot];
                    const int4 mi = S.meta[ki], mj = S.meta[j];
                    cd.b[0] = ki; cd.b[1] = j;
                    cd.b[2] = chain_nb(mi, -1); cd.b[3] = chain_nb(mj, -1);
                    cd.b[4] = chain_nb(mi, +1); cd.b[5] = chain_nb(mj, +1);
                    cd.ibp = ibp; cd.type = S.bp_type[pbase + ibp]; cd.rep = r; cd.last = -1; cd.dG = dG;
                }
            }
            base += nb;
        }
    }
    cta_sync();
    const int ncand = min(sh.ncand, FAST_BPC);
    if (sh.ncand > FAST_BPC && bt == 0) S.misc->max_force_flag |= 4;   // the host sizes the grid so that this cannot happen
    const int bl = bt & 7, bg = bt >> 3;
    const unsigned gmask = 0xFFu << ((tid & 31) & ~7);
    if (!sh.blocked[0] && sh.ncand <= FAST_BPC)
    for (int n = 0; n < C.nsteps; n++) {
        for (int i = bt; i < ncand; i += FAST_BP) {            // distance gate, one thread per candidate
            FastBpCand &cd = sh.cand[i];
            const double4 *pos = S.pos[cur] + (size_t)cd.rep * nmp;
            const Vec3 v12 = pbc_vec_d(P, v3(pos[cd.b[0]]), v3(pos[cd.b[1]]));
            const double a12 = sqrt(dot(v12, v12));
            const int type = cd.type;
            if (fabs(a12 - P.bp[type].bond_r) <= P.bp[type].cutoff_ddist) {
                sh.active[atomicAdd(&sh.nactive, 1)] = i;
            } else {
                if (P.bp_triplet && P.flg_bias_ss) {           // src/force_bp_triplet.F90:85-93
                    const double s = P.bias_ss_force / a12;
                    double *fi = S.f_bp + ((size_t)cd.rep * nmp + cd.b[0]) * 3, *fj = S.f_bp + ((size_t)cd.rep * nmp + cd.b[1]) * 3;
                    atomicAdd(fi, -s * v12.x); atomicAdd(fi + 1, -s * v12.y); atomicAdd(fi + 2, -s * v12.z);
                    atomicAdd(fj, s * v12.x); atomicAdd(fj + 1, s * v12.y); atomicAdd(fj + 2, s * v12.z);
                }
                if (cd.last != 0) {
                    const size_t pb = (size_t)cd.rep * nmp * KB + cd.ibp;
                    S.bp_status[pb] = 0; S.ene_bp[pb] = 0.0;
                    cd.last = 0;
                }
            }
        }
        named_bar_sync(1, FAST_BP);
        const int nact = sh.nactive;
        for (int a = bg; a < nact; a += FAST_BPGROUPS) {       // survivors: one role per lane
            FastBpCand &cd = sh.cand[sh.active[a]];
            const double4 *pos = S.pos[cur] + (size_t)cd.rep * nmp;
            const int type = cd.type;
            const Vec3 x0 = v3(pos[cd.b[0]]), x1 = v3(pos[cd.b[1]]);
            const int tA = bp_role_A(bl), tB = bp_role_B(bl);
            Vec3 xa = x0, xb = x0;
            if (tA >= 0) xa = v3(pos[cd.b[tA]]);
            if (tB >= 0) xb = v3(pos[cd.b[tB]]);
            const Vec3 v12 = pbc_vec_d(P, x0, x1);
            const double d1212 = dot(v12, v12), a12 = sqrt(d1212);
            const BpSub o = bp_force_role(P, P.bp[type], bl, x0, x1, xa, xb, v12, d1212, a12);
            const double u = group_sum_masked(o.u, gmask);
            const double ene = P.bp_triplet ? P.bp[type].U0 * cd.dG * exp(-u) : P.bp[type].U0 * exp(-u);
            Vec3 F0, F1;
            F0.x = group_sum_masked(o.c0.x, gmask); F0.y = group_sum_masked(o.c0.y, gmask); F0.z = group_sum_masked(o.c0.z, gmask);
            F1.x = group_sum_masked(o.c1.x, gmask); F1.y = group_sum_masked(o.c1.y, gmask); F1.z = group_sum_masked(o.c1.z, gmask);
            // bead t = lane (2..5): ca of its angle role + (ca | cb) of its dihedral role
            const int src_ang = (bl >= 2 && bl <= 5) ? bl - 1 : 0;
            const int src_dih = (bl == 2 || bl == 3) ? 5 : 6;
            const bool dih_b = (bl == 2 || bl == 4);
            Vec3 Fa, Fda, Fdb;
            Fa.x = __shfl_sync(gmask, o.ca.x, src_ang, 8); Fa.y = __shfl_sync(gmask, o.ca.y, src_ang, 8); Fa.z = __shfl_sync(gmask, o.ca.z, src_ang, 8);
            Fda.x = __shfl_sync(gmask, o.ca.x, src_dih, 8); Fda.y = __shfl_sync(gmask, o.ca.y, src_dih, 8); Fda.z = __shfl_sync(gmask, o.ca.z, src_dih, 8);
            Fdb.x = __shfl_sync(gmask, o.cb.x, src_dih, 8); Fdb.y = __shfl_sync(gmask, o.cb.y, src_dih, 8); Fdb.z = __shfl_sync(gmask, o.cb.z, src_dih, 8);
            const bool on = ene <= P.bp_cutoff_energy;
            if (on && bl < 6) {
                const Vec3 Ft = bl == 0 ? F0 : bl == 1 ? F1 : Fa + (dih_b ? Fdb : Fda);
                double *fb = S.f_bp + ((size_t)cd.rep * nmp + cd.b[bl]) * 3;
                atomicAdd(fb, ene * Ft.x); atomicAdd(fb + 1, ene * Ft.y); atomicAdd(fb + 2, ene * Ft.z);
            }
            if (bl == 0) {
                const size_t pb = (size_t)cd.rep * nmp * KB + cd.ibp;
                S.bp_status[pb] = on ? 1 : 0;
                S.ene_bp[pb] = on ? ene : 0.0;
                cd.last = on ? 1 : 0;
            }
        }
        named_bar_sync(1, FAST_BP);
        if (bt == 0) { sh.nactive = 0; red_release_add(C.bar_bp, 1u); }
        cur ^= 1;
        cta_sync();
        if (sh.blocked[n & 1]) break;
    }
    cta_sync();
}

SIS_D void fast_loc_role(const DevSys &S, const FFParams &P, const StepArgs &A, const FastCtx &C, FastShared &sh)
{
    const int tid = threadIdx.x, lt = tid - FAST_MAIN - FAST_BP;
    const int nmp = C.nmp, R = C.R;
    unsigned long long step = C.step0;
    int cur = C.cur0;
    const bool wwarp = lt >= FAST_LOC - 32;            // the warp that evaluates the rebuild test
    const int ll = lt - (FAST_LOC - 32);
    // Gaussians of the first step
    const int gg = lt >> 1, gdraw = lt & 1, gitem = C.item0 + gg;
    int grep = 0, gk = 0;
    uint32_t gstream = 0;
    const bool gon = lt < 2 * FAST_GROUPS && gitem < C.nitems;
    if (gon) {
        fast_item(gitem, nmp, &grep, &gk);
        gstream = (uint32_t)S.rp[grep].stream;
        double g0, g1;
        gauss_pair(A.seed, gstream, step + 1ull, (uint32_t)gk, (uint32_t)gdraw, &g0, &g1);
        if (gdraw == 0) { sh.gauss[(step + 1ull) & 1][gg][0] = g0; sh.gauss[(step + 1ull) & 1][gg][1] = g1; }
        else sh.gauss[(step + 1ull) & 1][gg][2] = g0;
    }
    if (wwarp) {                                       // rebuild test on the state left by the previous launch
        bool b = false;
        for (int r = ll; r < R; r += 32) {
            const unsigned long long *t = S.top2 + ((size_t)r * 2 + (step & 1ull)) * 2;
            const double m1 = __longlong_as_double((long long)__ldcg(t));
            const double m2 = __longlong_as_double((long long)__ldcg(t + 1));
            b |= (sqrt(m1) + sqrt(m2) > S.nl_margin);
        }
        b = __any_sync(0xffffffffu, b);
        if (ll == 0) { sh.blocked[0] = b; sh.blocked[1] = 0; }
    }
    // bonded-term slot of this thread: [0, NT) angle + bond, [NT, 2 NT) dihedral
    const int kind = lt / FAST_NT, slot = lt - kind * FAST_NT;
    const int titem = C.item0 - 3 + slot;
    bool ton = lt < 2 * FAST_NT && titem >= 0 && titem < C.nitems && (kind == 0 || C.dih_on);
    int trep = 0, tk = 0;
    int4 tm = make_int4(0, 1, 0, 0);
    if (ton) { fast_item(titem, nmp, &trep, &tk); tm = S.meta[tk]; }
    const bool tcirc = tm.w & MF_CIRC;
    cta_sync();
    if (!sh.blocked[0] && sh.ncand <= FAST_BPC)
    for (int n = 0; n < C.nsteps; n++) {
        // rebuild test on the displacements left by the previous step
        if (n > 0 && wwarp) {
            bool b = false;
            if (R == 1) {
                double t1 = 0.0, t2 = 0.0;
                for (int c = ll; c < C.ncta; c += 32) {
                    const double2 s = __ldcg(S.top2_slots + (step & 1ull) * FAST_MAX_CTA + c);
                    top2_merge(t1, t2, s.x, s.y);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double b1 = __shfl_xor_sync(0xffffffffu, t1, o), b2 = __shfl_xor_sync(0xffffffffu, t2, o);
                    top2_merge(t1, t2, b1, b2);
                }
                b = sqrt(t1) + sqrt(t2) > S.nl_margin;
            } else {
                for (int r = ll; r < R; r += 32) {
                    const unsigned long long *t = S.top2 + ((size_t)r * 2 + (step & 1ull)) * 2;
                    const double m1 = __longlong_as_double((long long)__ldcg(t));
                    const double m2 = __longlong_as_double((long long)__ldcg(t + 1));
                    b |= (sqrt(m1) + sqrt(m2) > S.nl_margin);
                }
                b = __any_sync(0xffffffffu, b);
            }
            if (ll == 0) sh.blocked[n & 1] = b;
        }
        // bonded terms, each evaluated once (src/force_bond.F90, force_angl*.F90, force_dih_*.F90)
        if (ton) {
            const double4 *pos = S.pos[cur] + (size_t)trep * nmp;
            if (kind == 0) {
                if (tcirc ? tm.y > 1 : tm.z + 1 < tm.y) {
                    const Vec3 x0 = v3(pos[tk]), x1 = v3(pos[chain_nb(tm, 1)]);
                    sh.bondf[slot] = bond_force_vec(P, x0, x1);
                    if (tcirc ? tm.y > 2 : tm.z + 2 < tm.y) {
                        const Vec3 x2 = v3(pos[chain_nb(tm, 2)]);
                        Vec3 fa, fc;
                        if (P.flg_angl_ReB) angl_ReB_force_vecs(P, x0, x1, x2, &fa, &fc);
                        else angl_force_vecs(P, x0, x1, x2, &fa, &fc);
                        sh.angf[slot][0] = fa; sh.angf[slot][1] = fc;
                    }
                }
            } else if (tcirc ? tm.y > 3 : tm.z + 3 < tm.y) {
                Vec3 f0, f1, f2, f3;
                {
                    Vec3 f4[4];
                    dih_force_vecs(P, v3(pos[tk]), v3(pos[chain_nb(tm, 1)]), v3(pos[chain_nb(tm, 2)]), v3(pos[chain_nb(tm, 3)]), f4);
                    f0 = f4[0]; f1 = f4[1]; f2 = f4[2]; f3 = f4[3];
                }
                sh.dihf[slot][0] = f0; sh.dihf[slot][1] = f1; sh.dihf[slot][2] = f2; sh.dihf[slot][3] = f3;
            }
        }
        __threadfence_block();
        __syncwarp();
        bar_arrive(3, FAST_MAIN + FAST_LOC);
        // Gaussians of the next step (buffer parity of step + 2)
        if (gon && n + 1 < C.nsteps) {
            double g0, g1;
            gauss_pair(A.seed, gstream, step + 2ull, (uint32_t)gk, (uint32_t)gdraw, &g0, &g1);
            if (gdraw == 0) { sh.gauss[step & 1][gg][0] = g0; sh.gauss[step & 1][gg][1] = g1; }
            else sh.gauss[step & 1][gg][2] = g0;
        }
        step++;
        cur ^= 1;
        cta_sync();
        if (sh.blocked[n & 1]) break;
    }
    cta_sync();
}

__global__ void __launch_bounds__(FAST_THREADS, 1) k_md_fast(DevSys S, FFParams P, StepArgs A, int nsteps)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastShared &sh = *reinterpret_cast<FastShared *>(smem_raw);
    uint32_t *s_rows = reinterpret_cast<uint32_t *>(smem_raw + ((sizeof(FastShared) + 15) & ~size_t(15)));
    const DevCtl ctl = S.ctl[0];
    FastCtx C;
    C.R = S.nrep; C.nmp = S.nmp; C.K = S.K; C.nitems = S.nrep * S.nmp; C.ncta = gridDim.x;
    C.item0 = blockIdx.x * FAST_GROUPS; C.nsteps = nsteps;
    C.step0 = ctl.step; C.cur0 = ctl.cur;
    C.bar_bp = S.bar; C.bar_pos = S.bar + 32;
    C.dih_on = P.flg_dih_cos || P.flg_dih_exp;
    const int tid = threadIdx.x;
    if (tid >= FAST_MAIN + FAST_BP) { fast_loc_role(S, P, A, C, sh); return; }
    if (tid >= FAST_MAIN) { fast_bp_role(S, P, C, sh); return; }
    const int n = fast_main_role(S, P, A, C, sh, s_rows);
    // ---- bookkeeping for the host and the other kernels (block 0, warp 0) ---------------------------
    if (blockIdx.x == 0 && tid < 32) {
        const unsigned long long step = C.step0 + (unsigned long long)n;
        const int cur = C.cur0 ^ (n & 1);
        if (C.R == 1 && n > 0) {
            double t1 = 0.0, t2 = 0.0;
            for (int c = tid; c < C.ncta; c += 32) {
                const double2 s = __ldcg(S.top2_slots + (step & 1ull) * FAST_MAX_CTA + c);
                top2_merge(t1, t2, s.x, s.y);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double b1 = __shfl_xor_sync(0xffffffffu, t1, o), b2 = __shfl_xor_sync(0xffffffffu, t2, o);
                top2_merge(t1, t2, b1, b2);
            }
            if (tid == 0) {
                unsigned long long *t = S.top2 + (step & 1ull) * 2;
                t[0] = (unsigned long long)__double_as_longlong(t1); t[1] = (unsigned long long)__double_as_longlong(t2);
                t = S.top2 + ((step + 1ull) & 1ull) * 2;
                t[0] = 0ull; t[1] = 0ull;
            }
        }
        if (tid == 0) {
            DevCtl a; a.step = step; a.cur = cur; a.blocked = 0;
            S.ctl[0] = a;
        }
    }
}
