// sis_big.cuh -- k_md_big: the MD loop body (src/job_md.F90:374-460) for LARGE systems
// (more beads than k_md_fast can hold on chip: BASELINE config C5, 144 384 beads), ONE launch per
// MD step, replayed as CUDA-graph chunks with the device-driven list-rebuild chain in front.
//
// Why not a persistent kernel here: at this size a step is tens of microseconds of fp64 work, so a
// launch boundary (~2 us inside a graph) costs what a grid barrier costs, while ordinary launches
// let several CTAs share an SM and the hardware scheduler balance 2 000+ CTAs of unequal pair work
// over the 148 SMs.  (k_md_fused, the persistent general kernel of round 1, ran 211 us per C5 step
// at one 640-thread CTA per SM; the per-phase graph mode 145 us.)
//
// Everything is in GATHER form, so a step needs no atomics on forces, no force zeroing and no
// second kernel: a CTA owns `bpc` consecutive beads and computes their complete force:
//   pair terms  (src/force_wca.F90:34-66, src/force_ele_DH.F90:39-81): the CTA's rows are cut
//     into QUADS of 4 entries (one 16-byte load); the warps draw chunks of 32 quads from a shared
//     counter (a warp delayed by its other tasks simply takes fewer); a lane evaluates the four
//     pairs of its quad SIDE BY SIDE with branch-free arithmetic (independent fp64 dependency
//     chains: rsqrt by Newton steps, exp(-x) by a Taylor polynomial) and writes the quad's force to
//     a fixed shared-memory slot; the bead's owner adds its quads in order (deterministic);
//   bonded terms (src/force_bond.F90, force_angl*.F90, force_dih_*.F90): every term the CTA's beads
//     take part in is evaluated ONCE into shared tables (terms straddling two CTAs twice);
//   base pairs  (src/force_bp.F90:50-204, force_bp_triplet.F90:62-240): one thread per bead of the
//     range (+ its chain neighbours) walks the bead's candidate row and applies the distance gate
//     in the prologue; the survivors are evaluated by 8-lane groups (one role per lane,
//     bp_force_role) before those warps join the pair loop; the parts acting on the CTA's own beads
//     are added in shared memory.  A pair is evaluated by the 2-3 CTAs that own its six beads; the
//     CTA that owns bead i keeps bp_status / ene_bp (src/force_bp.F90:191-193) up to date;
//   integrator  (src/job_md.F90:409-456) with the counter-based Gaussians, xyz_move and the
//     displacement top-2 of the rebuild test (:376-388), 4 lanes per bead.
// Two block-wide barriers per step (after the prologue + gate, after the forces).  A variant with
// fully independent warps (16 beads each, no barrier) was measured slower: 58 M instead of 42 M warp
// instructions per C5 step (bonded tables and Gaussians at low lane occupancy) and 24 % of the
// samples without an instruction to issue (profiles/r02_k_md_big_variants.txt).
// Step bookkeeping without a host: launch `idx` of a chunk reads S.ctl[idx & 1] and the LAST CTA to
// finish publishes S.ctl[(idx + 1) & 1]; every CTA evaluates the rebuild test itself from the
// complete top-2 of the previous step, and a kernel that finds the step blocked (or the stop step
// reached, or the lists invalid) only passes the bookkeeping on.
#pragma once

#include "sis_fast.cuh"

constexpr int BIG_THREADS = 256;
constexpr int BIG_BPC = 64;                 // beads per CTA at most (4 integrate lanes per bead)
constexpr int BIG_NT = BIG_BPC + 3;         // bonded-term slots: terms starting at item0 - 3 .. item0 + bpc - 1
constexpr int BIG_QMAX = 1024;              // quad force slots per pass (24 KB)
constexpr int BIG_QMAP = BIG_BPC * 64;      // quads a CTA can map (rows of up to 256 entries per bead on average)
constexpr int BIG_ACT = 512;                // survivors of the base-pair distance gate held per CTA (4 bytes each)
constexpr int BIG_NEXTRA = 24;              // chain neighbours of the CTA's beads outside its range

struct BigBead {
    double x[3];
    double ele_coef, inv_lambdaD, cut2, noise;
    int4 meta;
    int rep, k, nn, qstart;
};
struct BigShared {
    BigBead bead[BIG_BPC];
    Vec3 quadF[BIG_QMAX];
    Vec3 bondf[BIG_NT], angf[BIG_NT][2], dihf[BIG_NT][4];
    double gauss[BIG_BPC][3];
    double fbp[BIG_BPC][3];
    unsigned act[BIG_ACT];                  // survivor = (index in E) << 8 | row slot of that bead's candidate row
    int extra[BIG_NEXTRA];
    unsigned short qmap[BIG_QMAP];          // quad -> owner slot
    double d2[BIG_BPC];                     // |xyz_move|^2 of the CTA's beads after the step, and their replica
    int d2rep[BIG_BPC];
    int nq, nact, nextra, qnext;
};

// one survivor of the distance gate on the 8 lanes of a group (role = lane): forces on the CTA's
// own beads into shared memory, bp_status / ene_bp by the CTA that owns bead i
SIS_D void big_bp_eval_one(const DevSys &S, const FFParams &P, BigShared &sh, const double4 *__restrict__ posall,
                           int b_item, int e, int lo_item, int hi_item)
{
    const int bl = threadIdx.x & 7;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
    const int nmp = S.nmp, KB = S.KB;
    int rep, k;
    fast_item(b_item, nmp, &rep, &k);
    const size_t rbase = (size_t)rep * nmp, pbase = rbase * KB;
    const double4 *pos = posall + rbase;
    const int2 ent = S.bprow[(rbase + k) * KB + e];
    const int ki = min(k, ent.x), kj = max(k, ent.x), ibp = ent.y;
    const int type = S.bp_type[pbase + ibp];
    double dG = 1.0;
    if (P.bp_triplet) {
        const double2 cf = S.bp_coef[pbase + ibp];
        dG = cf.x - S.rp[rep].tempK * cf.y;                 // src/force_bp_triplet.F90:77
    }
    const int4 mi = S.meta[ki], mj = S.meta[kj];
    // beads 0..5 = i, j, i-1, j-1, i+1, j+1
    auto bead_of = [&](int t) -> int {
        return t == 0 ? ki : t == 1 ? kj : t == 2 ? chain_nb(mi, -1) : t == 3 ? chain_nb(mj, -1)
             : t == 4 ? chain_nb(mi, +1) : chain_nb(mj, +1);
    };
    const Vec3 x0 = v3(ldg_pos(pos + ki)), x1 = v3(ldg_pos(pos + kj));
    const int tA = bp_role_A(bl), tB = bp_role_B(bl);
    Vec3 xa = x0, xb = x0;
    if (tA >= 0) xa = v3(ldg_pos(pos + bead_of(tA)));
    if (tB >= 0) xb = v3(ldg_pos(pos + bead_of(tB)));
    const Vec3 v12 = pbc_vec_d(P, x0, x1);
    const double d1212 = dot(v12, v12), a12 = sqrt(d1212);
    const BpSub o = bp_force_role(P, P.bp[type], bl, x0, x1, xa, xb, v12, d1212, a12);
    const double u = group_sum_masked(o.u, gmask);
    const double ene = P.bp_triplet ? P.bp[type].U0 * dG * exp(-u) : P.bp[type].U0 * exp(-u);
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
        const int it = rep * nmp + bead_of(bl);
        if (it >= lo_item && it < hi_item) {
            const Vec3 Ft = bl == 0 ? F0 : bl == 1 ? F1 : Fa + (dih_b ? Fdb : Fda);
            double *fb = sh.fbp[it - lo_item];
            atomicAdd(fb, ene * Ft.x); atomicAdd(fb + 1, ene * Ft.y); atomicAdd(fb + 2, ene * Ft.z);
        }
    }
    const int ii = rep * nmp + ki;
    if (bl == 0 && ii >= lo_item && ii < hi_item) {
        S.bp_status[pbase + ibp] = on ? 1 : 0;
        S.ene_bp[pbase + ibp] = on ? ene : 0.0;
    }
}

// (replica, bead) of a flattened item without the division when there is one replica
SIS_D void big_item(int item, int nmp, int R, int *rep, int *k)
{
    if (R == 1) { *rep = 0; *k = item; }
    else fast_item(item, nmp, rep, k);
}

// W pairs of a quad side by side, branch-free (src/force_wca.F90:34-66, src/force_ele_DH.F90:39-81).
// v = x_k - x_j with the single-image shift of pbc_vec_d written as v - L * rint(v / L): the same
// subtraction for every pair that can be in range (|v| < 1.5 L).  Every intermediate is an array over
// the W pairs and every step a loop over them, so the W fp64 dependency chains interleave (W = 2 is
// what 80 registers per thread hold without spilling).
template <int W>
SIS_D void big_pairs_force(const FFParams &P, const uint32_t *ent, const double4 *pj, Vec3 xk, Vec3 inv_box,
                           double ec, double il, double c2, Vec3 &F)
{
    if (P.mixed) {        // fp32 pair scalars (pair_scalar_f32), W pairs side by side
#pragma unroll
        for (int u = 0; u < W; u++) {
            double vx = xk.x - pj[u].x, vy = xk.y - pj[u].y, vz = xk.z - pj[u].z;
            if (P.flg_pbc) {
                vx = fma(-P.box[0], rint(vx * inv_box.x), vx);
                vy = fma(-P.box[1], rint(vy * inv_box.y), vy);
                vz = fma(-P.box[2], rint(vz * inv_box.z), vz);
            }
            const double d2 = vx * vx + vy * vy + vz * vz;
            const bool e_on = (ent[u] & NB_ELE) && d2 <= c2, w_on = (ent[u] & NB_WCA) && d2 < P.wca_sigma2;
            const double sc = pair_scalar_f32(P, w_on, e_on, (e_on || w_on) ? d2 : 1.0, ec, il);
            F.x = fma(sc, vx, F.x); F.y = fma(sc, vy, F.y); F.z = fma(sc, vz, F.z);
        }
        return;
    }
    double dd[W], y[W], h[W], e[W];
    bool ie[W], iw[W];
#pragma unroll
    for (int u = 0; u < W; u++) {
        double vx = xk.x - pj[u].x, vy = xk.y - pj[u].y, vz = xk.z - pj[u].z;
        if (P.flg_pbc) {
            vx = fma(-P.box[0], rint(vx * inv_box.x), vx);
            vy = fma(-P.box[1], rint(vy * inv_box.y), vy);
            vz = fma(-P.box[2], rint(vz * inv_box.z), vz);
        }
        const double d2 = vx * vx + vy * vy + vz * vz;
        ie[u] = (ent[u] & NB_ELE) && d2 <= c2;
        iw[u] = (ent[u] & NB_WCA) && d2 < P.wca_sigma2;
        dd[u] = (ie[u] || iw[u]) ? d2 : 1.0;
    }
    // y = 1 / sqrt(dd): MUFU.RSQ64H seed (2^-22) + two Newton steps
#pragma unroll
    for (int u = 0; u < W; u++) {
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[u]) : "d"(dd[u]));
        h[u] = 0.5 * dd[u];
    }
#pragma unroll
    for (int u = 0; u < W; u++) e[u] = fma(-h[u] * y[u], y[u], 0.5);
#pragma unroll
    for (int u = 0; u < W; u++) y[u] = fma(y[u], e[u], y[u]);
#pragma unroll
    for (int u = 0; u < W; u++) e[u] = fma(-h[u] * y[u], y[u], 0.5);
#pragma unroll
    for (int u = 0; u < W; u++) y[u] = fma(y[u], e[u], y[u]);
    // ex = exp(-r / lambda_D): r / lambda = k ln2 + rr, degree-13 Taylor polynomial, scaled through the exponent field
    double rr[W], p[W];
    int kk[W];
    const double magic = 6755399441055744.0;
#pragma unroll
    for (int u = 0; u < W; u++) {
        const double t = -(dd[u] * y[u]) * il;
        const double kz = fma(t, 1.4426950408889634, magic);
        kk[u] = __double2loint(kz);
        const double kf = kz - magic;
        rr[u] = fma(kf, -1.90821492927058770002e-10, fma(kf, -6.93147180369123816490e-01, t));
        p[u] = 1.6059043836821613e-10;
    }
#define BIG_HORNER(c) _Pragma("unroll") for (int u = 0; u < W; u++) p[u] = fma(p[u], rr[u], c)
    BIG_HORNER(2.08767569878681e-09); BIG_HORNER(2.505210838544172e-08); BIG_HORNER(2.755731922398589e-07);
    BIG_HORNER(2.7557319223985893e-06); BIG_HORNER(2.48015873015873e-05); BIG_HORNER(1.984126984126984e-04);
    BIG_HORNER(1.388888888888889e-03); BIG_HORNER(8.333333333333333e-03); BIG_HORNER(4.1666666666666664e-02);
    BIG_HORNER(1.6666666666666666e-01); BIG_HORNER(0.5); BIG_HORNER(1.0); BIG_HORNER(1.0);
#undef BIG_HORNER
#pragma unroll
    for (int u = 0; u < W; u++) {
        const double ex = __hiloint2double(__double2hiint(p[u]) + (kk[u] << 20), __double2loint(p[u]));
        const double ri2 = y[u] * y[u];
        double sc = ie[u] ? ec * ri2 * (y[u] + il) * ex : 0.0;
        const double x = P.wca_sigma2 * ri2, x2 = x * x, x4 = x2 * x2;
        sc += iw[u] ? P.wca_coef * (x4 * x2 * x - x4) : 0.0;
        double vx = xk.x - pj[u].x, vy = xk.y - pj[u].y, vz = xk.z - pj[u].z;    // recomputed: cheaper than keeping them live
        if (P.flg_pbc) {
            vx = fma(-P.box[0], rint(vx * inv_box.x), vx);
            vy = fma(-P.box[1], rint(vy * inv_box.y), vy);
            vz = fma(-P.box[2], rint(vz * inv_box.z), vz);
        }
        F.x = fma(sc, vx, F.x); F.y = fma(sc, vy, F.y); F.z = fma(sc, vz, F.z);
    }
}
#ifndef BIG_W
#define BIG_W 2
#endif
SIS_D Vec3 big_quad_force(const FFParams &P, const uint32_t ent[4], const double4 pj[4], Vec3 xk, Vec3 inv_box,
                          double ec, double il, double c2)
{
    Vec3 F = mk(0.0, 0.0, 0.0);
#pragma unroll
    for (int u0 = 0; u0 < 4; u0 += BIG_W) big_pairs_force<BIG_W>(P, ent + u0, pj + u0, xk, inv_box, ec, il, c2, F);
    return F;
}

#ifndef BIG_MINB
#define BIG_MINB 3
#endif
__global__ void __launch_bounds__(BIG_THREADS, BIG_MINB) k_md_big(DevSys S, FFParams P, StepArgs A, int idx, int bpc)
{
    extern __shared__ __align__(16) unsigned char big_smem[];
    BigShared &sh = *reinterpret_cast<BigShared *>(big_smem);
    const int tid = threadIdx.x, lane = tid & 31;
    const DevCtl ctl = S.ctl[idx & 1];
    const int R = S.nrep, nmp = S.nmp, K = S.K, KB = S.KB;
    const int nitems = R * nmp;
    const unsigned long long step = ctl.step;
    const int cur = ctl.cur;
    // ---- run or pass the bookkeeping on: stop step, invalid lists, rebuild test (src/job_md.F90:376-388) ----
    {
        bool blocked = (long long)step >= S.misc->stop_step || S.misc->lists_invalid != 0;
        for (int r = 0; r < R && !blocked; r++) {
            const unsigned long long *t = S.top2 + ((size_t)r * 2 + (step & 1ull)) * 2;
            const double m1 = __longlong_as_double((long long)__ldcg(t));
            const double m2 = __longlong_as_double((long long)__ldcg(t + 1));
            blocked = sqrt(m1) + sqrt(m2) > S.nl_margin;
        }
        if (blocked) {
            if (blockIdx.x == 0 && tid == 0) S.ctl[(idx + 1) & 1] = ctl;
            return;
        }
    }
    const int item0 = blockIdx.x * bpc;
    const int nb = min(bpc, nitems - item0);
    const int lo_item = item0, hi_item = item0 + nb;
    const double4 *__restrict__ posall = S.pos[cur];
    const bool dih_on = P.flg_dih_cos || P.flg_dih_exp;
    const unsigned long long istep = step + 1ull;
    const Vec3 inv_box = P.flg_pbc ? mk(1.0 / P.box[0], 1.0 / P.box[1], 1.0 / P.box[2]) : mk(0.0, 0.0, 0.0);

    // ---- prologue: the CTA's beads -> shared memory -----------------------------------------------
    if (tid < BIG_BPC) {
        BigBead &B = sh.bead[tid];
        int nn = 0;
        if (tid < nb) {
            int rep, k;
            big_item(item0 + tid, nmp, R, &rep, &k);
            nn = S.nnbr[(size_t)rep * nmp + k];
            const double4 p = ldg_pos(posall + (size_t)rep * nmp + k);
            const RepParams &Rp = S.rp[rep];
            B.x[0] = p.x; B.x[1] = p.y; B.x[2] = p.z;
            B.ele_coef = Rp.ele_coef; B.inv_lambdaD = Rp.inv_lambdaD;
            B.cut2 = P.flg_ele ? Rp.ele_cutoff * Rp.ele_cutoff : -1.0;
            B.noise = S.coef[k].w * Rp.sqrt_tempK;
            B.meta = S.meta[k];
            B.rep = rep; B.k = k;
        }
        B.nn = nn;
        sh.fbp[tid][0] = 0.0; sh.fbp[tid][1] = 0.0; sh.fbp[tid][2] = 0.0;
    }
    if (tid == 0) { sh.nact = 0; sh.nextra = 0; sh.qnext = 0; }
    __syncthreads();
    if (tid < 32) {                                    // qstart = exclusive scan of ceil(nn / 4)
        const int a0 = (sh.bead[2 * tid].nn + 3) >> 2, a1 = (sh.bead[2 * tid + 1].nn + 3) >> 2;
        int incl = a0 + a1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += t;
        }
        const int run = incl - (a0 + a1);
        sh.bead[2 * tid].qstart = run;
        sh.bead[2 * tid + 1].qstart = run + a0;
        if (tid == 31) sh.nq = incl;
    }
    // chain neighbours of the CTA's beads outside its range (base-pair candidates touch i+-1, j+-1)
    if (tid >= 64 && tid < 64 + nb) {
        const BigBead &B = sh.bead[tid - 64];
        for (int d = -1; d <= 1; d += 2) {
            const int kb = chain_nb(B.meta, d);
            if (kb < 0) continue;
            const int it = B.rep * nmp + kb;
            if (it >= lo_item && it < hi_item) continue;
            const int s = atomicAdd(&sh.nextra, 1);
            if (s < BIG_NEXTRA) sh.extra[s] = it;
        }
    }
    __syncthreads();
    // outside the design envelope (the host sizes bpc so that this cannot happen): flag it, the host stops with an error
    if ((sh.nq > BIG_QMAP || sh.nextra > BIG_NEXTRA) && tid == 0) S.misc->max_force_flag |= 8;
    const int nextra = min(sh.nextra, BIG_NEXTRA);
    const int nq = min(sh.nq, BIG_QMAP);
    {   // quad -> owner slot
        const int slot = tid >> 2;
        if (slot < nb) {
            const int q0 = sh.bead[slot].qstart, n = (sh.bead[slot].nn + 3) >> 2;
            for (int o = tid & 3; o < n; o += 4)
                if (q0 + o < BIG_QMAP) sh.qmap[q0 + o] = (unsigned short)slot;
        }
    }
    // ---- base-pair candidates of one bead of E = range + chain neighbours: distance gate ---------------
    // (the threads at the top: their warps take the evaluation of the survivors below)
    if ((A.term_mask & (1u << ENE_BP)) && BIG_THREADS - 1 - tid < nb + nextra) {
        const int ei = BIG_THREADS - 1 - tid;
        const int b_item = ei < nb ? lo_item + ei : sh.extra[ei - nb];
        bool dup = false;                              // duplicates in the extra list: first occurrence works
        for (int j = 0; j < ei - nb; j++) dup |= sh.extra[j] == b_item;
        if (!dup) {
            int rep, k;
            big_item(b_item, nmp, R, &rep, &k);
            const size_t rbase = (size_t)rep * nmp, pbase = rbase * KB;
            const double4 *pos = posall + rbase;
            const int nrow = S.nbp_row[rbase + k];
            auto in_E = [&](int it) {
                if (it >= lo_item && it < hi_item) return true;
                for (int j = 0; j < nextra; j++)
                    if (sh.extra[j] == it) return true;
                return false;
            };
            for (int e = 0; e < nrow; e++) {
                const int2 ent = S.bprow[(rbase + k) * KB + e];
                const int p_item = rep * nmp + ent.x;
                if (p_item < b_item && in_E(p_item)) continue;          // enumerated from the partner
                const int ki = min(k, ent.x), kj = max(k, ent.x), ibp = ent.y;
                const bool own = rep * nmp + ki >= lo_item && rep * nmp + ki < hi_item;
                double dG = 1.0;
                bool alive = true;
                if (P.bp_triplet) {
                    const double2 cf = S.bp_coef[pbase + ibp];
                    dG = cf.x - S.rp[rep].tempK * cf.y;                 // src/force_bp_triplet.F90:77
                    alive = dG < 0.0;
                }
                const int type = S.bp_type[pbase + ibp];
                bool pass = false;
                Vec3 v12 = mk(0.0, 0.0, 0.0);
                double a12 = 1.0;
                if (alive) {
                    v12 = pbc_vec_d(P, v3(ldg_pos(pos + ki)), v3(ldg_pos(pos + kj)));
                    a12 = sqrt(dot(v12, v12));
                    pass = fabs(a12 - P.bp[type].bond_r) <= P.bp[type].cutoff_ddist;
                }
                if (pass) {
                    const int sl = atomicAdd(&sh.nact, 1);
                    if (sl < BIG_ACT) sh.act[sl] = ((unsigned)ei << 8) | (unsigned)e;
                } else {
                    if (alive && P.bp_triplet && P.flg_bias_ss) {       // src/force_bp_triplet.F90:85-93
                        const double sc = P.bias_ss_force / a12;
                        const int ii = rep * nmp + ki, jj = rep * nmp + kj;
                        if (ii >= lo_item && ii < hi_item) {
                            double *fi = sh.fbp[ii - lo_item];
                            atomicAdd(fi, -sc * v12.x); atomicAdd(fi + 1, -sc * v12.y); atomicAdd(fi + 2, -sc * v12.z);
                        }
                        if (jj >= lo_item && jj < hi_item) {
                            double *fj = sh.fbp[jj - lo_item];
                            atomicAdd(fj, sc * v12.x); atomicAdd(fj + 1, sc * v12.y); atomicAdd(fj + 2, sc * v12.z);
                        }
                    }
                    if (own) { S.bp_status[pbase + ibp] = 0; S.ene_bp[pbase + ibp] = 0.0; }
                }
            }
        }
    }
    __syncthreads();                                   // bead table, quad map, survivors of the gate
    // ---- survivors of the distance gate: 8-lane groups (the groups of the top warps first) ----------------
    {
        const int nact = min(sh.nact, BIG_ACT);
        if (sh.nact > BIG_ACT && tid == 0) S.misc->max_force_flag |= 8;
        for (int a = (BIG_THREADS - 1 - tid) >> 3; a < nact; a += BIG_THREADS / 8) {
            const unsigned rec = sh.act[a];
            const int ei = (int)(rec >> 8), e = (int)(rec & 0xFFu);
            big_bp_eval_one(S, P, sh, posall, ei < nb ? lo_item + ei : sh.extra[ei - nb], e, lo_item, hi_item);
        }
    }
    // ---- once-per-CTA tasks (bonded tables, Gaussians), dealt to the low threads --------------------------
    {
        const int nslot = nb + 3;                              // bonded-term slots in use
        const int n_bond = nslot, n_dih = dih_on ? nslot : 0, n_gauss = 2 * nb;
        const int ntask = n_bond + n_dih + n_gauss;
        for (int task = tid; task < ntask; task += BIG_THREADS) {
            if (task < n_bond + n_dih) {
                // bonded terms starting at item0 - 3 + slot (src/force_bond.F90, force_angl*.F90, force_dih_*.F90)
                const int kind = task >= n_bond, slot = task - kind * n_bond;
                const int titem = item0 - 3 + slot;
                if (titem < 0 || titem >= nitems) continue;
                int trep, tk;
                big_item(titem, nmp, R, &trep, &tk);
                const int4 tm = S.meta[tk];
                const bool tcirc = tm.w & MF_CIRC;
                const double4 *pos = posall + (size_t)trep * nmp;
                if (kind == 0) {
                    if (tcirc ? tm.y > 1 : tm.z + 1 < tm.y) {
                        const Vec3 x0 = v3(ldg_pos(pos + tk)), x1 = v3(ldg_pos(pos + chain_nb(tm, 1)));
                        sh.bondf[slot] = bond_force_vec(P, x0, x1);
                        if (tcirc ? tm.y > 2 : tm.z + 2 < tm.y) {
                            const Vec3 x2 = v3(ldg_pos(pos + chain_nb(tm, 2)));
                            Vec3 fa, fc;
                            if (P.flg_angl_ReB) angl_ReB_force_vecs(P, x0, x1, x2, &fa, &fc);
                            else angl_force_vecs(P, x0, x1, x2, &fa, &fc);
                            sh.angf[slot][0] = fa; sh.angf[slot][1] = fc;
                        }
                    }
                } else if (tcirc ? tm.y > 3 : tm.z + 3 < tm.y) {
                    Vec3 f4[4];
                    dih_force_vecs(P, v3(ldg_pos(pos + tk)), v3(ldg_pos(pos + chain_nb(tm, 1))), v3(ldg_pos(pos + chain_nb(tm, 2))),
                                   v3(ldg_pos(pos + chain_nb(tm, 3))), f4);
                    sh.dihf[slot][0] = f4[0]; sh.dihf[slot][1] = f4[1]; sh.dihf[slot][2] = f4[2]; sh.dihf[slot][3] = f4[3];
                }
            } else {
                // Gaussians of this step: draw 0 -> (x, y), draw 1 -> z of bead g
                const int t = task - n_bond - n_dih, g = t >> 1, draw = t & 1;
                const BigBead &B = sh.bead[g];
                double g0, g1;
                gauss_pair(A.seed, (uint32_t)S.rp[B.rep].stream, istep, (uint32_t)B.k, (uint32_t)draw, &g0, &g1);
                if (draw == 0) { sh.gauss[g][0] = g0; sh.gauss[g][1] = g1; }
                else sh.gauss[g][2] = g0;
            }
        }
    }
    // ---- pair quads: the warps draw chunks of 32 quads; pass p covers quads [p * QMAX, (p + 1) * QMAX) ---------
    const int oslot = tid >> 2, comp = tid & 3;
    const bool owner = oslot < nb && comp < 3;
    double f = 0.0;
    for (int q0 = 0; q0 < nq || q0 == 0; q0 += BIG_QMAX) {
        const int q1 = min(nq, q0 + BIG_QMAX);
        for (;;) {
            int qb = 0;
            if (lane == 0) qb = atomicAdd(&sh.qnext, 32);
            qb = __shfl_sync(0xffffffffu, qb, 0);
            if (qb >= q1) break;
            const int q = qb + lane;
            if (q < q1) {
                const int slot = sh.qmap[q];
                const BigBead &B = sh.bead[slot];
                const int e0 = 4 * (q - B.qstart);
                const uint32_t *row = S.nbr + ((size_t)B.rep * nmp + B.k) * K + e0;
                const uint4 e4 = __ldg(reinterpret_cast<const uint4 *>(row));
                const int nv = B.nn - e0;                   // >= 1 valid entries in this quad
                const uint32_t ent[4] = {e4.x, nv > 1 ? e4.y : 0u, nv > 2 ? e4.z : 0u, nv > 3 ? e4.w : 0u};
                const double4 *pr = posall + (size_t)B.rep * nmp;
                double4 pj[4];
#pragma unroll
                for (int u = 0; u < 4; u++) pj[u] = ldg_pos(pr + (ent[u] & NB_IDX));
                sh.quadF[q - q0] = big_quad_force(P, ent, pj, mk(B.x[0], B.x[1], B.x[2]), inv_box, B.ele_coef, B.inv_lambdaD, B.cut2);
            }
        }
        __syncthreads();
        if (owner) {
            const BigBead &B = sh.bead[oslot];
            const int a = max(B.qstart, q0) - q0, b = min(B.qstart + ((B.nn + 3) >> 2), q1) - q0;
            for (int q = a; q < b; q++) f += (&sh.quadF[q].x)[comp];
        }
        if (q1 < nq) {
            // the chunk that ran past q1 is handed back: the counter restarts at the pass boundary
            __syncthreads();
            if (tid == 0) sh.qnext = q1;
            __syncthreads();
        }
    }
    // ---- total force on the CTA's beads, integrate (src/job_md.F90:415-456) ------------------------------
    double d2 = 0.0, mv = 0.0;
    int myrep = -1;
    if (oslot < nb) {
        BigBead &B = sh.bead[oslot];
        myrep = B.rep;
        if (comp < 3) {
            const int4 meta = B.meta;
            const int l = meta.z, len = meta.y, s0 = oslot + 3;
            const bool circ = meta.w & MF_CIRC;
            if (A.term_mask & (1u << ENE_BOND)) {
                if (circ ? len > 1 : l + 1 < len) f -= (&sh.bondf[s0].x)[comp];
                if (l >= 1) f += (&sh.bondf[s0 - 1].x)[comp];
            }
            if (A.term_mask & (1u << ENE_ANGL)) {
                if (circ ? len > 2 : l + 2 < len) f += (&sh.angf[s0][0].x)[comp];
                if (l >= 1 && (circ ? len > 2 : l + 1 < len)) f -= (&sh.angf[s0 - 1][0].x)[comp] + (&sh.angf[s0 - 1][1].x)[comp];
                if (l >= 2) f += (&sh.angf[s0 - 2][1].x)[comp];
            }
            if (dih_on && (A.term_mask & (1u << ENE_DIHE))) {
                if (circ ? len > 3 : l + 3 < len) f += (&sh.dihf[s0][0].x)[comp];
                if (l >= 1 && (circ ? len > 3 : l + 2 < len)) f += (&sh.dihf[s0 - 1][1].x)[comp];
                if (l >= 2 && (circ ? len > 3 : l + 1 < len)) f += (&sh.dihf[s0 - 2][2].x)[comp];
                if (l >= 3) f += (&sh.dihf[s0 - 3][3].x)[comp];
            }
            const double4 *pos = posall + (size_t)B.rep * nmp;
            if (circ && l < 3) f += fast_closure_terms(S, P, pos, meta, B.k, comp);
            if (P.flg_stage || P.ntwz_DCF || P.ntwz_FR || P.nrest) {
                const Vec3 fb = bias_force(S, P, S.rp[B.rep], pos, B.k, mk(B.x[0], B.x[1], B.x[2]), A.term_mask, istep);
                f += comp == 0 ? fb.x : comp == 1 ? fb.y : fb.z;
            }
            f += sh.fbp[oslot][comp];
            const size_t o3 = ((size_t)B.rep * nmp + B.k) * 3 + comp;
            const double4 c = S.coef[B.k];
            const double g = sh.gauss[oslot][comp];
            double dx;
            if (P.integrator == 1) {                   // GJF-2GJ
                const double accels_pre = B.noise * g;
                const double v = c.x * S.vel[o3] + c.y * f + (S.acc[o3] + accels_pre);
                S.vel[o3] = v;
                S.acc[o3] = accels_pre;
                dx = c.z * v;
            } else {                                   // Ermak-McCammon
                dx = c.x * f + B.noise * g;
            }
            B.x[comp] += dx;
            mv = S.move[o3] + dx;
            S.move[o3] = mv;
            if (f > 100.0) S.misc->max_force_flag |= 1;
        }
    }
    __syncwarp();
    {
        const double s0 = __shfl_sync(0xffffffffu, mv, 0, 4);
        const double s1 = __shfl_sync(0xffffffffu, mv, 1, 4);
        const double s2 = __shfl_sync(0xffffffffu, mv, 2, 4);
        if (oslot < nb && comp == 0) {
            const BigBead &B = sh.bead[oslot];
            d2 = s0 * s0 + s1 * s1 + s2 * s2;
            S.pos[cur ^ 1][(size_t)B.rep * nmp + B.k] = make_double4(B.x[0], B.x[1], B.x[2], 0.0);
        }
    }
    // ---- displacement top-2 of the CTA per replica it holds (src/job_md.F90:376-388): warp 0 reduces ----
    if (comp == 0) { sh.d2[oslot] = oslot < nb ? d2 : 0.0; sh.d2rep[oslot] = myrep; }
    __syncthreads();
    if (tid < 32) {
        const int rep_lo = R == 1 ? 0 : item0 / nmp, rep_hi = R == 1 ? 0 : (hi_item - 1) / nmp;
        for (int r = rep_lo; r <= rep_hi; r++) {
            double t1 = sh.d2rep[tid] == r ? sh.d2[tid] : 0.0, t2 = sh.d2rep[tid + 32] == r ? sh.d2[tid + 32] : 0.0;
            if (t2 > t1) { const double t = t1; t1 = t2; t2 = t; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double b1 = __shfl_xor_sync(0xffffffffu, t1, o), b2 = __shfl_xor_sync(0xffffffffu, t2, o);
                top2_merge(t1, t2, b1, b2);
            }
            if (tid == 0 && t1 > 0.0) top2_push(S.top2 + ((size_t)r * 2 + (istep & 1ull)) * 2, t1, t2);
        }
    }
    // ---- the last CTA to finish publishes the step and clears the top-2 slots step + 2 accumulates into ----
    if (tid == 0) {
        __threadfence();
        const unsigned done = atomicAdd(S.bar, 1u);
        if (done == gridDim.x - 1) {
            __threadfence();
            for (int r = 0; r < R; r++) {
                unsigned long long *t = S.top2 + ((size_t)r * 2 + (step & 1ull)) * 2;
                t[0] = 0ull; t[1] = 0ull;
            }
            DevCtl a; a.step = istep; a.cur = cur ^ 1; a.blocked = 0;
            S.ctl[(idx + 1) & 1] = a;
            *S.bar = 0u;
        }
    }
}
