// sis_fast.cuh -- k_md_fast: the MD loop body (src/job_md.F90:374-460) for systems with at most
// ~96 beads per SM on average (BASELINE configs C1..C4).
//
// Persistent, warp-specialised, one CTA per SM; everything that is constant between two list
// rebuilds is held on chip for the whole launch:
//   MAIN  (16 warps = 64 groups x 8 lanes).  The rows of the CTA's beads are packed octet-aligned
//         in shared memory; a UNIT is one octet of one row.  Units are dealt round-robin to the
//         groups (src/force_wca.F90, src/force_ele_DH.F90), reduced by shuffles into fixed slots.
//         Each 4-lane half of a group owns one bead (128 owner slots per CTA): position /
//         velocity / beta / displacement / coefficients in shared memory, integrated
//         (src/job_md.F90:409-456) as soon as the other force parts are complete.  Bead ranges are
//         cut by pair work (k_nl_partition), not by bead count.
//   BP    (4 warps): every base-pair candidate that touches one of the CTA's beads is cached in
//         shared memory (six bead ids, type, dG); one thread per candidate applies the distance
//         gate, 8-lane groups evaluate the survivors (one role per lane, bp_force_role) drawn from a
//         shared counter and add the parts acting on the CTA's own beads to a shared-memory force
//         array: pairs are evaluated redundantly by the 2-3 CTAs owning their beads, so the
//         base-pair term needs no global atomics and no grid barrier.
//   LOC   (4 warps): every bonded term the CTA's beads take part in is evaluated ONCE
//         (src/force_bond.F90, force_angl*.F90, force_dih_*.F90) into shared tables the owners
//         read; then they join the base-pair evaluation (a CTA with more formed base pairs than BP
//         groups would otherwise need a second round of the ~2 us evaluation chain), as the MAIN
//         groups do when their units are done; then the Gaussians of the NEXT step (Philox +
//         Box-Muller) into a double-buffered shared array; one LOC warp also evaluates the rebuild test.
// One grid barrier per step: bar[32] "positions written".  The rebuild test of
// src/job_md.F90:376-388 for R == 1 goes through per-CTA (max, 2nd max) slots read after the
// position barrier, off the critical path; with R > 1 through the atomic top-2 of k_md_fused.
// Replica-local mode (template LOCAL): one CTA owns one whole replica -- no grid barrier, the lists
// are rebuilt inside the kernel (fast_local_rebuild) and the Monte-Carlo limit of base pairs per
// nucleotide (src/force_bp_limit_triplet.F90, nstep_bp_MC > 0) runs in the BP role (fast_bp_mc_step).
#pragma once

#include "sis_fused.cuh"

// -DSIS_FAST_PROF: the phase clocks of option "profile" (bench.py --profile) compiled in; without it the step loops
// carry no clock reads (the kernel is large enough for its instruction footprint to matter)
#ifdef SIS_FAST_PROF
constexpr bool FAST_PROF = true;
#else
constexpr bool FAST_PROF = false;
#endif
constexpr int FAST_MAIN_WARPS = 16, FAST_BP_WARPS = 4, FAST_LOC_WARPS = 4;
constexpr int FAST_MAIN = 32 * FAST_MAIN_WARPS, FAST_BP = 32 * FAST_BP_WARPS, FAST_LOC = 32 * FAST_LOC_WARPS;
constexpr int FAST_THREADS = FAST_MAIN + FAST_BP + FAST_LOC;
constexpr int FAST_GROUPS = FAST_MAIN / 8;        // 8-lane groups of the pair loop
constexpr int FAST_OWN = 2 * FAST_GROUPS;         // owner slots = beads per CTA at most: the two 4-lane halves of a
                                                  // group each own one bead (x, y, z on three lanes), so the bead
                                                  // ranges can be cut by pair work instead of by bead count
constexpr int FAST_EMAX = 8192;                   // row entries (padded to octets) of one CTA held in shared memory
constexpr int FAST_BPGROUPS = FAST_BP / 8;
constexpr int FAST_NT = FAST_OWN + 3;             // term-table slots: starts at item0-3 .. item0+127
constexpr int FAST_BPC = 256;                     // cached bp candidates per CTA
constexpr int FAST_MAX_CTA = 1024;                // capacity of the top-2 slot array

struct FastBpCand {
    int b[6];       // flattened (replica, bead) items of beads i, j, i-1, j-1, i+1, j+1
    int type, last; // last = bp_status last written (-1 unknown)
    int own;        // this CTA owns bead i: it keeps bp_status / ene_bp of the pair up to date
    int pad;
    size_t pb;      // index of the pair in bp_status / ene_bp
    double dG;
};
constexpr int REMD_LOCAL_MAX = 1024;          // replicas (labels) the in-kernel exchange of the replica-local mode handles
constexpr int FAST_NEXTRA = 160;   // chain neighbours of the CTA's beads outside its range

// state of one owned bead, resident in shared memory for the whole launch
struct FastBead {
    double x[3], v[3], a[3], m[3];     // position, velocity, beta (accels), xyz_move
    double ca, cb, cc, noise;          // a, sqrt(b)h/m, sqrt(b)h of set_md_coef; noise coefficient * sqrt(T)
    double ele_coef, inv_lambdaD, cut2;
    int4 meta;
    int rep, k, nn, valid;
};

// exchange-step bookkeeping of a launch (identical in every CTA: each replays every decision)
struct FastRemdState {
    int idx_pair, attempt;
    unsigned long long counter;
    long long natt, nacc;
};

struct FastShared {
    Vec3 bondf[FAST_NT];        // bond starting at slot: F(first) = -f, F(second) = +f
    Vec3 angf[FAST_NT][2];      // angle starting at slot: F1, F3; F2 = -(F1 + F3)
    Vec3 dihf[FAST_NT][4];      // dihedral starting at slot: F1..F4
    double gauss[2][FAST_OWN][3];
    double d2[FAST_OWN];
    int nn[FAST_OWN], unit_start[FAST_OWN + 1];   // row lengths; first pair unit (octet) of each bead
    FastBead bead[FAST_OWN];
    FastBpCand cand[FAST_BPC];
    int active[FAST_BPC];
    int nactive[2], bp_next[2];      // survivors of the distance gate / next one to evaluate, by step parity
    int roles_done[2];               // role chunks of the step's survivors that are complete (fast_bp_eval_roles), by step parity
    int ncand, nextra, overflow, overflow_rows;
    int gate_done[2];                // step index + 1 of the last completed distance-gate pass, by step parity
    int extra[FAST_NEXTRA];
    double fbp[FAST_OWN][3];         // base-pair forces on the CTA's beads (this step)
    int4 tmeta[FAST_NT];             // bonded-term slots: meta of the first bead
    int tbead[FAST_NT];              // ... its flattened item, -1 = no such bead
    int blocked[2];             // rebuild verdict, indexed by iteration parity
    double top[2];
    // exchange step inside the kernel (fast_remd_exchange): energy reduction scratch and this CTA's private
    // copy of the label tables (every CTA replays every decision, so the copies stay identical)
    double xred[16][2 * ENE_N];
    double xE[4];
    int x_newlab, x_natt, x_nacc;
    double x_vscale;
    FastRemdState xs;
    int x_r2l[REMD_LOCAL_MAX], x_l2r[REMD_LOCAL_MAX];
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

// one listed pair: v = x_k - x_j (+ image shift when flagged), WCA and Debye-Hueckel scalars
SIS_D void fast_pair_one(const FFParams &P, uint32_t ent, Vec3 xk, Vec3 xj, double ele_coef, double inv_lambdaD,
                         double cut2, Vec3 &F)
{
    Vec3 v = xk - xj;
    if (ent & NB_IMG) {
        v.x = minimg1(v.x, P.box[0], P.box_half[0]);
        v.y = minimg1(v.y, P.box[1], P.box_half[1]);
        v.z = minimg1(v.z, P.box[2], P.box_half[2]);
    }
    const double d2 = dot(v, v);
    double s = 0.0;
    if (P.mixed) {
        const bool w_on = (ent & NB_WCA) && d2 < P.wca_sigma2, e_on = (ent & NB_ELE) && d2 <= cut2;
        if (w_on || e_on) s = pair_scalar_f32(P, w_on, e_on, d2, ele_coef, inv_lambdaD);
        F += s * v;
        return;
    }
    if ((ent & NB_WCA) && d2 < P.wca_sigma2) s += wca_scalar(P, d2);
#ifdef FAST_PAIR_LIBM
    if ((ent & NB_ELE) && d2 <= cut2) s += dh_scalar(ele_coef, inv_lambdaD, d2);
#else
    if ((ent & NB_ELE) && d2 <= cut2) {
        // the shorter dependency chain of the branch-free forms (Newton rsqrt on the MUFU seed, Taylor exp): the pair
        // loop of this kernel is bound by the latency of one pair's chain, not by instruction count
        const double ri = rsqrt_sl(d2);
        s += ele_coef * ri * ri * (ri + inv_lambdaD) * exp_neg_sl((d2 * ri) * inv_lambdaD);
    }
#endif
    F += s * v;
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
    int R, nmp, K, nitems, ncta, item0, nb, nsteps;   // item0 / nb: the CTA's bead range
    unsigned long long step0;
    int cur0;
    unsigned *bar_pos;
    bool dih_on;
    bool local;      // replica-local mode: the CTA owns every bead of its replica, no grid barrier
    bool no_loc_help;     // option: the LOC warps do not join the base-pair evaluation
    bool after_rebuild;   // this segment follows an in-kernel list rebuild: its first step is a Monte-Carlo step
    unsigned long long seed;
    int max_bp_per_nt;
    bool xch;             // exchange steps inside the kernel (replica-local mode, S.remd set)
};
SIS_D void fast_remd_hook(const DevSys &S, const FFParams &P, const FastCtx &C, FastShared &sh, unsigned long long step, int cur);
SIS_D void cta_sync() { named_bar_sync(0, FAST_THREADS); }
SIS_D void fast_bp_eval(const DevSys &S, const FFParams &P, FastShared &sh, const double4 *pos, int lo_item, int hi_item,
                        int q, int nact);

// Loop protocol shared by the roles (iteration n, parity q = n & 1):
//   LOC   : [n > 0: verdict -> sh.blocked[q]]  tables  arrive(3)  gauss  cta_sync  read blocked[q]
//   MAIN  : pair  sync(3)  read blocked[q] (leave)  integrate  arrive+wait bar_pos  cta_sync
//   BP    : gate  eval  arrive(3)  cta_sync  read blocked[q]
// A MAIN thread that leaves at barrier 3 performs the iteration's cta_sync before returning, so
// every role executes the same number of cta_syncs.  Returns the number of completed steps.

// MAIN role.  The rows of the CTA's beads are packed into shared memory octet-aligned: row of owner
// slot s starts at entry 8 * unit_start[s], so UNIT u is simply entries [8u, 8u + 8).  Units are
// dealt round-robin to the 64 groups (one entry per lane), so every group runs the same number of
// iterations whatever the row lengths; each unit's force is written to a fixed slot and the bead's
// owner adds its units in order (deterministic, no atomics).  Owner slot s < nb is held by the
// 4-lane half (s / 64) of group s % 64.
template <int U, bool XCH>
SIS_D int fast_main_role(const DevSys &S, const FFParams &P, const StepArgs &A, const FastCtx &C, FastShared &sh,
                         uint32_t *s_rows)
{
    (void)A;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 7, grp = tid >> 3;
    const int nmp = C.nmp, R = C.R;
    unsigned long long step = C.step0;
    int cur = C.cur0;
    const bool prof = FAST_PROF && S.prof != nullptr && tid == 0;   // every CTA's thread 0: per-CTA phase clocks
    long long c0 = 0, c1 = 0, c1b = 0, c2 = 0, c3 = 0, c4 = 0;
    // shared-memory regions behind the rows: prologue scratch, later the unit tables
    uint32_t *tmp_all = s_rows + FAST_EMAX;
    unsigned char *cls_all = reinterpret_cast<unsigned char *>(s_rows + 2 * FAST_EMAX);
    Vec3 *unitF = reinterpret_cast<Vec3 *>(tmp_all);                 // [nunit] (24 bytes per octet of 32-byte scratch)
    unsigned char *unit_q = cls_all;                                 // [nunit] owner slot of the unit

    // ---- prologue: the beads of the CTA's range -> shared memory ----------------------------------
    if (tid < FAST_OWN) {
        int nn = 0;
        if (tid < C.nb) {
            int rep, k;
            fast_item(C.item0 + tid, nmp, &rep, &k);
            nn = S.nnbr[(size_t)rep * nmp + k];
        }
        sh.nn[tid] = nn;
    }
    named_bar_sync(2, FAST_MAIN);
    if (warp == 0) {                                   // unit_start = exclusive scan of ceil(nn / 8)
        int a[4], tot = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { a[j] = (sh.nn[4 * tid + j] + 7) >> 3; tot += a[j]; }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += t;
        }
        int run = incl - tot;
#pragma unroll
        for (int j = 0; j < 4; j++) { sh.unit_start[4 * tid + j] = run; run += a[j]; }
        if (tid == 31) { sh.unit_start[FAST_OWN] = incl; sh.overflow_rows = incl * 8 > FAST_EMAX; }
    }
    named_bar_sync(2, FAST_MAIN);
    const bool rows_fit = !sh.overflow_rows;
    for (int slot = grp; slot < FAST_OWN; slot += FAST_GROUPS) {
        FastBead &B = sh.bead[slot];
        if (slot < C.nb && rows_fit) {
            int rep, k;
            fast_item(C.item0 + slot, nmp, &rep, &k);
            const int nn = sh.nn[slot], e0 = 8 * sh.unit_start[slot];
            const double4 p = S.pos[cur][(size_t)rep * nmp + k];
            const Vec3 xk0 = v3(p);
            const RepParams &Rp = S.rp[rep];
            const double cut2 = P.flg_ele ? Rp.ele_cutoff * Rp.ele_cutoff : -1.0;
            // row -> shared memory as a stable partition: pairs inside the Debye-Hueckel / WCA range
            // now first, the margin shell after them (the lanes of a warp then leave the expensive
            // branch together); bead-id order, i.e. coalesced gathers, is kept inside each class
            const uint32_t *row = S.nbr + ((size_t)rep * nmp + k) * C.K;
            uint32_t *tmp = tmp_all + e0;
            unsigned char *cls = cls_all + e0;
            const double4 *pos0 = S.pos[cur] + (size_t)rep * nmp;
            // pass 1: class of every entry (bit c of `mine` = chunk c of this lane is in range), near count
            const unsigned gm = 0xFFu << ((tid & 31) & ~7);
            const int gshift = (tid & 31) & ~7;
            unsigned long long mine = 0ull;   // one bit per octet of the row: up to 512 entries
            int nnear = 0;
            const int nchunk = (nn + 7) >> 3;
            for (int c = 0; c < nchunk; c++) {
                const int e = 8 * c + lane;
                bool near = false;
                if (e < nn) {
                    const uint32_t ent = row[e];
                    const Vec3 v = pbc_vec_d(P, xk0, v3(pos0[ent & NB_IDX]));
                    const double d2 = dot(v, v);
                    near = ((ent & NB_ELE) && d2 <= cut2) || ((ent & NB_WCA) && d2 < P.wca_sigma2);
                    tmp[e] = ent;
                }
                const unsigned m = (__ballot_sync(gm, near) >> gshift) & 0xFFu;
                if (near) mine |= 1ull << c;
                nnear += __popc(m);
            }
            // pass 2: stable placement with running counters (near class first, then the margin shell)
            int pn = 0, pf = nnear;
            for (int c = 0; c < nchunk; c++) {
                const int e = 8 * c + lane;
                const bool in = e < nn, near = (mine >> c) & 1ull;
                const unsigned mn = (__ballot_sync(gm, in && near) >> gshift) & 0xFFu;
                const unsigned mf = (__ballot_sync(gm, in && !near) >> gshift) & 0xFFu;
                const unsigned below = (1u << lane) - 1u;
                if (in) s_rows[e0 + (near ? pn + __popc(mn & below) : pf + __popc(mf & below))] = tmp[e];
                pn += __popc(mn); pf += __popc(mf);
            }
            (void)cls;
            if (lane == 0) {
                B.meta = S.meta[k];
                B.rep = rep; B.k = k; B.nn = nn; B.valid = 1;
                const double4 coef = S.coef[k];
                B.ca = coef.x; B.cb = coef.y; B.cc = coef.z;
                B.noise = coef.w * Rp.sqrt_tempK;
                B.ele_coef = Rp.ele_coef; B.inv_lambdaD = Rp.inv_lambdaD; B.cut2 = cut2;
            }
            if (lane < 3) {
                const size_t o3 = ((size_t)rep * nmp + k) * 3 + lane;
                B.x[lane] = lane == 0 ? p.x : lane == 1 ? p.y : p.z;
                B.v[lane] = S.vel[o3]; B.a[lane] = S.acc[o3]; B.m[lane] = S.move[o3];
            }
        } else if (lane == 0) {
            B.valid = 0; B.nn = 0; B.rep = 0; B.k = 0;
        }
    }
    named_bar_sync(2, FAST_MAIN);                      // rows sorted, scratch free
    if (rows_fit)
        for (int slot = grp; slot < C.nb; slot += FAST_GROUPS) {
            const int u0 = sh.unit_start[slot], nu = sh.unit_start[slot + 1] - u0;
            for (int o = lane; o < nu; o += 8) unit_q[u0 + o] = (unsigned char)slot;
        }
    if (tid == 0 && !rows_fit) S.misc->max_force_flag |= 4;   // the host checks the flag after the launch
    cta_sync();                                        // prologue of all roles complete
    const int nunit = sh.unit_start[FAST_OWN];
    const bool overflow = sh.ncand > FAST_BPC || sh.overflow || sh.overflow_rows;
    const unsigned gmask = 0xFFu << ((tid & 31) & ~7);
    // owner slot of this lane: half `lane >> 2` of the group, component `lane & 3` (3 idles)
    const int sub = lane >> 2, comp = lane & 3, slot = grp + FAST_GROUPS * sub;
    FastBead &B = sh.bead[slot];
    int n = 0;
    if (!sh.blocked[0] && !overflow)
    for (; n < C.nsteps; n++) {
        if (prof) c0 = clock64();
        // ---- pair units: software pipeline, the next unit's entry and position are fetched before
        // the current pair is evaluated --------------------------------------------------------------
        {
            const double4 *posb = S.pos[cur];
            int u = grp;
            uint32_t ent = 0;
            Vec3 xj = mk(0.0, 0.0, 0.0);
            int q = 0;
            if (u < nunit) {
                q = unit_q[u];
                const int e = 8 * u + lane;
                ent = e - 8 * sh.unit_start[q] < sh.bead[q].nn ? s_rows[e] : 0u;
                if (ent) xj = v3(posb[(size_t)sh.bead[q].rep * nmp + (ent & NB_IDX)]);
            }
            while (u < nunit) {
                const int un = u + FAST_GROUPS;
                uint32_t ent_n = 0;
                Vec3 xn = mk(0.0, 0.0, 0.0);
                int qn = 0;
                if (un < nunit) {
                    qn = unit_q[un];
                    const int e = 8 * un + lane;
                    ent_n = e - 8 * sh.unit_start[qn] < sh.bead[qn].nn ? s_rows[e] : 0u;
                    if (ent_n) xn = v3(posb[(size_t)sh.bead[qn].rep * nmp + (ent_n & NB_IDX)]);
                }
                Vec3 F = mk(0.0, 0.0, 0.0);
                if (ent) {
                    const FastBead &Q = sh.bead[q];
                    fast_pair_one(P, ent, mk(Q.x[0], Q.x[1], Q.x[2]), xj, Q.ele_coef, Q.inv_lambdaD, Q.cut2, F);
                }
                F.x = group_sum_masked(F.x, gmask); F.y = group_sum_masked(F.y, gmask); F.z = group_sum_masked(F.z, gmask);
                if (lane == 0) unitF[u] = F;
                u = un; ent = ent_n; xj = xn; q = qn;
            }
        }
        if (prof) c1 = clock64();
        // help the BP warps with the survivors of the distance gate
        named_bar_sync(4, FAST_MAIN + FAST_BP);
        fast_bp_eval(S, P, sh, S.pos[cur], C.item0, C.item0 + C.nb, n & 1, sh.nactive[n & 1]);
        __threadfence_block();
        named_bar_sync(3, FAST_MAIN + FAST_LOC + FAST_BP);   // all units, bonded tables, base-pair forces, rebuild verdict
        if (prof) c1b = clock64();
        if (sh.blocked[n & 1]) { cta_sync(); break; }
        const bool valid = B.valid != 0 && comp < 3;
        double f = 0.0;
        if (valid) {
            for (int u = sh.unit_start[slot]; u < sh.unit_start[slot + 1]; u++) f += (&unitF[u].x)[comp];
            const int4 meta = B.meta;
            const int l = meta.z, len = meta.y, s0 = slot + 3;
            const bool circ = meta.w & MF_CIRC;
            if (circ ? len > 1 : l + 1 < len) f -= (&sh.bondf[s0].x)[comp];
            if (l >= 1) f += (&sh.bondf[s0 - 1].x)[comp];
            if (circ ? len > 2 : l + 2 < len) f += (&sh.angf[s0][0].x)[comp];
            if (l >= 1 && (circ ? len > 2 : l + 1 < len)) f -= (&sh.angf[s0 - 1][0].x)[comp] + (&sh.angf[s0 - 1][1].x)[comp];
            if (l >= 2) f += (&sh.angf[s0 - 2][1].x)[comp];
            if (C.dih_on) {
                if (circ ? len > 3 : l + 3 < len) f += (&sh.dihf[s0][0].x)[comp];
                if (l >= 1 && (circ ? len > 3 : l + 2 < len)) f += (&sh.dihf[s0 - 1][1].x)[comp];
                if (l >= 2 && (circ ? len > 3 : l + 1 < len)) f += (&sh.dihf[s0 - 2][2].x)[comp];
                if (l >= 3) f += (&sh.dihf[s0 - 3][3].x)[comp];
            }
            if (circ && l < 3) f += fast_closure_terms(S, P, S.pos[cur] + (size_t)B.rep * nmp, meta, B.k, comp);
            // dual constant-force tweezers (src/force_tweezers.F90:27-35): +f on the first bead of a pair, -f on the second
            for (int t = 0; t < P.ntwz_DCF; t++) {
                const double ft = S.rp[B.rep].twz_force * S.twz_dir[3 * t + comp];
                if (S.twz_pairs[2 * t] == B.k) f += ft;
                if (S.twz_pairs[2 * t + 1] == B.k) f -= ft;
            }
        }
        if (prof) c2 = clock64();
        // ---- integrate (src/job_md.F90:415-456) ----------------------------------------------
        double mv = 0.0;
        if (valid) {
            f += sh.fbp[slot][comp];
            sh.fbp[slot][comp] = 0.0;
            const double g = sh.gauss[(step + 1ull) & 1][slot][comp];
            double dx;
            if (P.integrator == 1) {
                const double accels_pre = B.noise * g;
                const double v = B.ca * B.v[comp] + B.cb * f + (B.a[comp] + accels_pre);
                B.v[comp] = v;
                B.a[comp] = accels_pre;
                dx = B.cc * v;
            } else {
                dx = B.ca * f + B.noise * g;
            }
            B.x[comp] += dx;
            mv = B.m[comp] + dx;
            B.m[comp] = mv;
            if (f > 100.0) S.misc->max_force_flag |= 1;
        }
        __syncwarp();
        const double s0 = __shfl_sync(0xffffffffu, mv, 0, 4);
        const double s1 = __shfl_sync(0xffffffffu, mv, 1, 4);
        const double s2 = __shfl_sync(0xffffffffu, mv, 2, 4);
        const bool owner = B.valid != 0;
        const double d2 = owner ? s0 * s0 + s1 * s1 + s2 * s2 : 0.0;
        if (comp == 0) {
            if (owner) S.pos[cur ^ 1][(size_t)B.rep * nmp + B.k] = make_double4(B.x[0], B.x[1], B.x[2], 0.0);
            if (R == 1 || C.local) sh.d2[slot] = d2;
            else if (d2 > 0.0) top2_push(S.top2 + ((size_t)B.rep * 2 + ((step + 1ull) & 1ull)) * 2, d2, 0.0);
        }
        named_bar_sync(2, FAST_MAIN);
        if (warp == 0) {
            if (R == 1 || C.local) {                   // CTA top-2 of |xyz_move|^2 -> this CTA's slot
                double t1 = sh.d2[tid], t2 = sh.d2[tid + 32];
                if (t2 > t1) { const double t = t1; t1 = t2; t2 = t; }
                double t3 = sh.d2[tid + 64], t4 = sh.d2[tid + 96];
                if (t4 > t3) { const double t = t3; t3 = t4; t4 = t; }
                top2_merge(t1, t2, t3, t4);
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
                if (C.local) {
                    // no other CTA reads these positions; the acquire drops the SM's stale L1 lines of
                    // the buffer written two steps ago (positions are gathered through read-only loads)
                    __threadfence();
                    (void)ld_acquire_u32(C.bar_pos);
                } else {
                    // (a two-level arrival -- groups of 6 / 12 / 24 CTAs on their own line, the last of a group on the
                    // root -- was measured: 10.86 / 10.95 / 10.91 us per C4 step against 10.81 flat, so the flat one stays)
                    red_release_add(C.bar_pos, 1u);
                    const unsigned target = (unsigned)(n + 1) * (unsigned)C.ncta;
                    while (ld_acquire_u32(C.bar_pos) < target) { }
                }
                if (prof) {
                    c4 = clock64();
                    unsigned long long *pc = S.prof + 16 + 8 * blockIdx.x;
                    pc[0] += (unsigned long long)(c1 - c0); pc[1] += (unsigned long long)(c2 - c1);
                    pc[2] += (unsigned long long)(c3 - c2); pc[3] += (unsigned long long)(c4 - c3);
                    pc[5] += (unsigned long long)(c1b - c1);
                    pc[7] += (unsigned long long)nunit * 1000ull + (unsigned long long)C.nb;   // units, beads of the CTA
                    if (blockIdx.x == 0) {
                        S.prof[0] += (unsigned long long)(c1 - c0); S.prof[1] += (unsigned long long)(c2 - c1);
                        S.prof[2] += (unsigned long long)(c3 - c2); S.prof[3] += (unsigned long long)(c4 - c3);
                        S.prof[4] += 1ull;
                    }
                }
            }
        }
        step++;
        cur ^= 1;
        cta_sync();
        if (XCH) fast_remd_hook(S, P, C, sh, step, cur);
    }
    cta_sync();                                        // everybody is out of the loop
    if (B.valid && comp < 3 && rows_fit) {
        const size_t o3 = ((size_t)B.rep * nmp + B.k) * 3 + comp;
        S.vel[o3] = B.v[comp]; S.acc[o3] = B.a[comp]; S.move[o3] = B.m[comp];
    }
    return n;
}

// Evaluation of the base-pair candidates that passed the distance gate: 8-lane groups (one role per
// lane, bp_force_role) draw survivors from a shared counter.  The BP groups start right after the
// gate; the MAIN groups join when their pair units are done, so a CTA with many formed base pairs
// is not held up by its four BP warps.
// (kept out of line: it runs only with more survivors than role slots, or behind option "bp_groups")
__device__ __noinline__ void fast_bp_eval_groups(const DevSys &S, const FFParams &P, FastShared &sh, const double4 *pos, int lo_item, int hi_item,
                               int q, int nact)
{
    const int bl = threadIdx.x & 7;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
    auto in_range = [&](int it) { return it >= lo_item && it < hi_item; };
    for (;;) {
        int a = 0;
        if (bl == 0) a = atomicAdd(&sh.bp_next[q], 1);
        a = __shfl_sync(gmask, a, 0, 8);
        if (a >= nact) break;
        FastBpCand &cd = sh.cand[sh.active[a]];
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
        if (on && bl < 6 && in_range(cd.b[bl])) {
            const Vec3 Ft = bl == 0 ? F0 : bl == 1 ? F1 : Fa + (dih_b ? Fdb : Fda);
            double *fb = sh.fbp[cd.b[bl] - lo_item];
            atomicAdd(fb, ene * Ft.x); atomicAdd(fb + 1, ene * Ft.y); atomicAdd(fb + 2, ene * Ft.z);
        }
        if (bl == 0 && cd.own) {
            S.bp_status[cd.pb] = on ? 1 : 0;
            S.ene_bp[cd.pb] = on ? ene : 0.0;
            cd.last = on ? 1 : 0;
        }
    }
}

// Role-major evaluation of the survivors (as k_md_rep): whole warps draw chunks of 32 items from the shared counter.
// First the role items -- one angle (4 per survivor) or one dihedral / the bond (3 per survivor) per lane, so a warp runs
// ONE code path instead of the bond / angle / dihedral lanes of an 8-lane group one after the other; then, once every
// role chunk is complete, one lane per (survivor, bead): pair energy from the seven exponent parts, the bead's force,
// added to the CTA's shared force array if the bead is one of its own.  The role results (BP_ROLE doubles per survivor)
// live in the shared-memory scratch behind the unit tables, free once the prologue is over.
constexpr int FAST_ACT = 48;          // survivors of one CTA the role slots hold (more: the 8-lane groups)
SIS_D double *fast_bp_role_slots(FastShared &sh)
{
    return reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(&sh) + ((sizeof(FastShared) + 15) & ~size_t(15)) +
                                      2 * FAST_EMAX * sizeof(uint32_t) + 4096);
}
SIS_D void fast_bp_eval_roles(const DevSys &S, const FFParams &P, FastShared &sh, const double4 *pos, int lo_item, int hi_item,
                              int q, int nact)
{
    const int lane = threadIdx.x & 31;
    double *slots = fast_bp_role_slots(sh);
    const int cang = (4 * nact + 31) >> 5, cdih = (3 * nact + 31) >> 5, nrole = cang + cdih, cbead = (6 * nact + 31) >> 5;
    for (;;) {
        int ch = 0;
        if (lane == 0) ch = atomicAdd(&sh.bp_next[q], 1);
        ch = __shfl_sync(0xffffffffu, ch, 0);
        if (ch >= nrole + cbead) break;
        if (ch < nrole) {
            const bool ang = ch < cang;
            const int it = 32 * (ang ? ch : ch - cang) + lane;
            int a = -1, role = 0;
            if (ang) { if (it < 4 * nact) { a = it >> 2; role = 1 + (it & 3); } }
            else if (it < 3 * nact) { a = it / 3; const int r3 = it - 3 * a; role = r3 == 2 ? 0 : 5 + r3; }
            if (a >= 0) {
                const FastBpCand &cd = sh.cand[sh.active[a]];
                const Vec3 x0 = v3(pos[cd.b[0]]), x1 = v3(pos[cd.b[1]]);
                const int tA = bp_role_A(role), tB = bp_role_B(role);
                const Vec3 xa = tA >= 0 ? v3(pos[cd.b[tA]]) : x0, xb = tB >= 0 ? v3(pos[cd.b[tB]]) : x0;
                bp_role_eval(P, P.bp[cd.type], role, x0, x1, xa, xb, slots + BP_ROLE * a);
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) atomicAdd(&sh.roles_done[q], 1);
        } else {
            while (*(volatile int *)&sh.roles_done[q] < nrole) __nanosleep(32);
            __threadfence_block();
            const int it = 32 * (ch - nrole) + lane;
            if (it < 6 * nact) {
                const int a = it / 6, t = it - 6 * a;
                FastBpCand &cd = sh.cand[sh.active[a]];
                const double *r = slots + BP_ROLE * a;
                const double u = bp_roles_u(r);
                const double ene = P.bp_triplet ? P.bp[cd.type].U0 * cd.dG * exp(-u) : P.bp[cd.type].U0 * exp(-u);
                const bool on = ene <= P.bp_cutoff_energy;
                const int it_bead = cd.b[t];
                if (on && it_bead >= lo_item && it_bead < hi_item) {
                    const Vec3 Ft = bp_roles_bead_force(r, t);
                    double *fb = sh.fbp[it_bead - lo_item];
                    atomicAdd(fb, ene * Ft.x); atomicAdd(fb + 1, ene * Ft.y); atomicAdd(fb + 2, ene * Ft.z);
                }
                if (t == 0 && cd.own) {
                    S.bp_status[cd.pb] = on ? 1 : 0;
                    S.ene_bp[cd.pb] = on ? ene : 0.0;
                    cd.last = on ? 1 : 0;
                }
            }
        }
    }
}
SIS_D void fast_bp_eval(const DevSys &S, const FFParams &P, FastShared &sh, const double4 *pos, int lo_item, int hi_item,
                        int q, int nact)
{
    if (nact <= FAST_ACT && !S.dbg_bp_groups) fast_bp_eval_roles(S, P, sh, pos, lo_item, hi_item, q, nact);
    else fast_bp_eval_groups(S, P, sh, pos, lo_item, hi_item, q, nact);
}

// Monte-Carlo step of the BP role (replica-local mode): passes 1-2 of src/force_bp_limit_triplet.F90.
// Only instantiated in the replica-local kernel (fast_bp_role<true>).
SIS_D void fast_bp_mc_step(const DevSys &S, const FFParams &P, const FastCtx &C, FastShared &sh,
                                             const double4 *pos, int ncand)
{
    const int bt = threadIdx.x - FAST_MAIN, nmp = C.nmp;
    const int rep = C.item0 / nmp;
    const size_t rbase = (size_t)rep * nmp;
    for (int i = bt; i < nmp; i += FAST_BP) S.nt_excess[rbase + i] = -C.max_bp_per_nt;
    named_bar_sync(1, FAST_BP);
    for (int i = bt; i < ncand; i += FAST_BP) {    // pass 1 of force_bp_limit_triplet: k_bp_eval<0>
        FastBpCand &cd = sh.cand[i];
        Vec3 xs[6], fr[6], v12;
#pragma unroll
        for (int t = 0; t < 6; t++) xs[t] = v3(pos[cd.b[t]]);
        double u, a12;
        int st = 0;
        double en = 0.0;
        if (bp_force_all(P, P.bp[cd.type], xs, &u, fr, &a12, &v12)) {
            const double ene = P.bp_triplet ? P.bp[cd.type].U0 * cd.dG * exp(-u) : P.bp[cd.type].U0 * exp(-u);
            if (ene <= P.bp_cutoff_energy) { st = 1; en = ene; }
        }
        S.bp_status[cd.pb] = st; S.ene_bp[cd.pb] = en;
        cd.last = -1;
        if (st) { atomicAdd(&S.nt_excess[cd.b[0]], 1); atomicAdd(&S.nt_excess[cd.b[1]], 1); }
    }
    __threadfence_block();
    named_bar_sync(1, FAST_BP);
    if (bt < 32) bp_prune_warp(S, P, C.seed, rep, 0.0, 0, 0, 0);
    __threadfence_block();
    named_bar_sync(1, FAST_BP);
    for (int i = bt; i < ncand; i += FAST_BP) S.bp_status_MC[sh.cand[i].pb] = S.bp_status[sh.cand[i].pb];
    __threadfence_block();
    named_bar_sync(1, FAST_BP);
}

// BP role.  Every CTA evaluates every base-pair candidate that touches one of its beads (as i, j
// or one of the four chain neighbours) and keeps only the force parts that act on its own beads,
// in shared memory: no cross-CTA dependency, no global atomics, no grid barrier for the base
// pairs.  A pair is evaluated by the two or three CTAs that own its six beads; the CTA that owns
// bead i also maintains bp_status / ene_bp (src/force_bp.F90:191-193).
template <bool LOCAL, bool XCH>
SIS_D int fast_bp_role(const DevSys &S, const FFParams &P, const FastCtx &C, FastShared &sh)
{
    const int tid = threadIdx.x, bt = tid - FAST_MAIN;
    const int nmp = C.nmp, KB = S.KB;
    const int lo_item = C.item0, hi_item = C.item0 + C.nb;
    int cur = C.cur0;
    auto in_range = [&](int it) { return it >= lo_item && it < hi_item; };
    // ---- E = the CTA's beads + their chain neighbours outside the range ---------------------------
    if (bt == 0) { sh.ncand = 0; sh.nactive[0] = sh.nactive[1] = 0; sh.bp_next[0] = sh.bp_next[1] = 0; sh.roles_done[0] = sh.roles_done[1] = 0; sh.nextra = 0; sh.overflow = 0; sh.gate_done[0] = sh.gate_done[1] = 0; }
    for (int i = bt; i < FAST_OWN * 3; i += FAST_BP) (&sh.fbp[0][0])[i] = 0.0;
    named_bar_sync(1, FAST_BP);
    for (int q = bt; q < C.nb; q += FAST_BP) {
        int rep, k;
        fast_item(lo_item + q, nmp, &rep, &k);
        const int4 m = S.meta[k];
        for (int d = -1; d <= 1; d += 2) {
            const int kb = chain_nb(m, d);
            if (kb < 0) continue;
            const int it = rep * nmp + kb;
            if (in_range(it)) continue;
            const int slot = atomicAdd(&sh.nextra, 1);
            if (slot < FAST_NEXTRA) sh.extra[slot] = it;
        }
    }
    named_bar_sync(1, FAST_BP);
    if (bt == 0) {                                     // remove duplicates (a handful of entries)
        int n = min(sh.nextra, FAST_NEXTRA), m = 0;
        for (int i = 0; i < n; i++) {
            const int v = sh.extra[i];
            bool dup = false;
            for (int j = 0; j < m; j++) dup |= sh.extra[j] == v;
            if (!dup) sh.extra[m++] = v;
        }
        if (sh.nextra > FAST_NEXTRA) { S.misc->max_force_flag |= 4; sh.overflow = 1; }
        sh.nextra = m;
    }
    named_bar_sync(1, FAST_BP);
    const int nextra = sh.nextra;
    auto in_E = [&](int it) {
        if (in_range(it)) return true;
        for (int j = 0; j < nextra; j++)
            if (sh.extra[j] == it) return true;
        return false;
    };
    // ---- candidates: every list pair with a bead in E, taken from the lower E member --------------
    for (int ei = bt; ei < C.nb + nextra; ei += FAST_BP) {
        const int b_item = ei < C.nb ? lo_item + ei : sh.extra[ei - C.nb];
        int rep, k;
        fast_item(b_item, nmp, &rep, &k);
        const size_t rbase = (size_t)rep * nmp, pbase = rbase * KB;
        const int nrow = S.nbp_row[rbase + k];
        for (int e = 0; e < nrow; e++) {
            const int2 ent = S.bprow[(rbase + k) * KB + e];
            const int p_item = rep * nmp + ent.x;
            if (in_E(p_item) && p_item < b_item) continue;      // enumerated from the partner
            const int ki = min(k, ent.x), kj = max(k, ent.x), ibp = ent.y;
            const bool own = in_range(rep * nmp + ki);
            double dG = 1.0;
            if (P.bp_triplet) {
                const double2 cf = S.bp_coef[pbase + ibp];
                dG = cf.x - S.rp[rep].tempK * cf.y;             // src/force_bp_triplet.F90:77
                if (dG >= 0.0 && !C.xch) {                      // never active at this temperature (an exchange can change that)
                    if (own) { S.bp_status[pbase + ibp] = 0; S.ene_bp[pbase + ibp] = 0.0; }
                    continue;
                }
            }
            const int slot = atomicAdd(&sh.ncand, 1);
            if (slot < FAST_BPC) {
                FastBpCand &cd = sh.cand[slot];
                const int4 mi = S.meta[ki], mj = S.meta[kj];
                const int base = rep * nmp;
                cd.b[0] = base + ki; cd.b[1] = base + kj;
                cd.b[2] = base + chain_nb(mi, -1); cd.b[3] = base + chain_nb(mj, -1);
                cd.b[4] = base + chain_nb(mi, +1); cd.b[5] = base + chain_nb(mj, +1);
                cd.type = S.bp_type[pbase + ibp]; cd.last = -1; cd.own = own; cd.pb = pbase + ibp; cd.dG = dG;
            }
        }
    }
    cta_sync();
    const int ncand = min(sh.ncand, FAST_BPC);
    if (sh.ncand > FAST_BPC && bt == 0) S.misc->max_force_flag |= 4;   // the host checks the flag after the launch
    const bool prof = FAST_PROF && S.prof != nullptr && bt == 0;
    long long b0 = 0, b1 = 0, b2 = 0;
    int n = 0;
    if (!sh.blocked[0] && sh.ncand <= FAST_BPC && !sh.overflow)
    for (; n < C.nsteps; n++) {
        if (prof) b0 = clock64();
        const double4 *pos = S.pos[cur];
        // Monte-Carlo limit of base pairs per nucleotide (src/force_bp_limit_triplet.F90:85-328), replica-local
        // mode only: every nstep_bp_MC steps and right after a rebuild (src/job_md.F90:396-401) all candidates
        // are evaluated, the per-nucleotide excess is counted, one warp runs the sequential delete loop
        // (bp_prune_warp, the same code and uniform stream as k_bp_prune) and the survivors become
        // bp_status_MC; between those steps only pairs with bp_status_MC are evaluated
        const bool use_mc = LOCAL && P.nstep_bp_MC > 0;
        if (use_mc) {
            const unsigned long long istep = C.step0 + (unsigned long long)n + 1ull;
            const bool mc = (istep % (unsigned long long)P.nstep_bp_MC == 0ull) || (n == 0 && C.after_rebuild);
            if (LOCAL && mc) fast_bp_mc_step(S, P, C, sh, pos, ncand);
        }
        for (int i = bt; i < ncand; i += FAST_BP) {            // distance gate, one thread per candidate
            FastBpCand &cd = sh.cand[i];
            const Vec3 v12 = pbc_vec_d(P, v3(pos[cd.b[0]]), v3(pos[cd.b[1]]));
            const double a12 = sqrt(dot(v12, v12));
            const int type = cd.type;
            const bool masked_in = (!use_mc || S.bp_status_MC[cd.pb] != 0) && (!P.bp_triplet || cd.dG < 0.0);   // use_mask = 1 of the stream path; dG >= 0: cycle
            if (masked_in && fabs(a12 - P.bp[type].bond_r) <= P.bp[type].cutoff_ddist) {
                sh.active[atomicAdd(&sh.nactive[n & 1], 1)] = i;
            } else {
                if (P.bp_triplet && P.flg_bias_ss && cd.dG < 0.0) {   // src/force_bp_triplet.F90:85-93
                    const double sc = P.bias_ss_force / a12;
                    if (in_range(cd.b[0])) {
                        double *fi = sh.fbp[cd.b[0] - lo_item];
                        atomicAdd(fi, -sc * v12.x); atomicAdd(fi + 1, -sc * v12.y); atomicAdd(fi + 2, -sc * v12.z);
                    }
                    if (in_range(cd.b[1])) {
                        double *fj = sh.fbp[cd.b[1] - lo_item];
                        atomicAdd(fj, sc * v12.x); atomicAdd(fj + 1, sc * v12.y); atomicAdd(fj + 2, sc * v12.z);
                    }
                }
                if (cd.own && cd.last != 0) {
                    S.bp_status[cd.pb] = 0; S.ene_bp[cd.pb] = 0.0;
                    cd.last = 0;
                }
            }
        }
        named_bar_sync(1, FAST_BP);
        if (prof) b1 = clock64();
        const int nact = sh.nactive[n & 1];
        if (bt == 0) { __threadfence_block(); *(volatile int *)&sh.gate_done[n & 1] = n + 1; }   // the LOC groups may join
        bar_arrive(4, FAST_MAIN + FAST_BP);                    // gate complete: the MAIN groups may join the evaluation
        if (bt == 0) { sh.nactive[(n + 1) & 1] = 0; sh.bp_next[(n + 1) & 1] = 0; sh.roles_done[(n + 1) & 1] = 0; }
        fast_bp_eval(S, P, sh, pos, lo_item, hi_item, n & 1, nact);
        if (prof) b2 = clock64();
        __threadfence_block();
        __syncwarp();
        bar_arrive(3, FAST_MAIN + FAST_LOC + FAST_BP);         // base-pair forces of this CTA's beads are in sh.fbp
        if (prof) {
            const long long b3 = clock64();
            S.prof[16 + 8 * blockIdx.x + 4] += (unsigned long long)(b3 - b0);
            S.prof[16 + 8 * blockIdx.x + 6] += (unsigned long long)nact;
            if (blockIdx.x == 0) {
                S.prof[5] += (unsigned long long)(b1 - b0); S.prof[6] += (unsigned long long)(b2 - b1);
                S.prof[7] += (unsigned long long)(b3 - b2);
            }
        }
        cur ^= 1;
        cta_sync();
        if (sh.blocked[n & 1]) break;
        if (XCH) fast_remd_hook(S, P, C, sh, C.step0 + (unsigned long long)n + 1ull, cur);
    }
    cta_sync();
    return n;
}

template <bool XCH>
SIS_D int fast_loc_role(const DevSys &S, const FFParams &P, const StepArgs &A, const FastCtx &C, FastShared &sh)
{
    const int tid = threadIdx.x, lt = tid - FAST_MAIN - FAST_BP;
    const int nmp = C.nmp, R = C.R;
    unsigned long long step = C.step0;
    int cur = C.cur0;
    const bool wwarp = lt >= FAST_LOC - 32;            // the warp that evaluates the rebuild test
    const int ll = lt - (FAST_LOC - 32);
    // Gaussians of the first step: thread lt draws (x, y) or z of beads lt / 2 and lt / 2 + 64
    const int gdraw = lt & 1;
    auto gauss_fill = [&](unsigned long long gstep) {
        for (int gg = lt >> 1; gg < C.nb; gg += FAST_LOC / 2) {
            int grep, gk;
            fast_item(C.item0 + gg, nmp, &grep, &gk);
            double g0, g1;
            gauss_pair(A.seed, (uint32_t)S.rp[grep].stream, gstep, (uint32_t)gk, (uint32_t)gdraw, &g0, &g1);
            if (gdraw == 0) { sh.gauss[gstep & 1][gg][0] = g0; sh.gauss[gstep & 1][gg][1] = g1; }
            else sh.gauss[gstep & 1][gg][2] = g0;
        }
    };
    gauss_fill(step + 1ull);
    if (wwarp) {                                       // rebuild test on the state left by the previous launch
        bool b = false;
        const int rlo = C.local ? C.item0 / nmp : 0, rhi = C.local ? rlo + 1 : R;
        for (int r = rlo + ll; r < rhi; r += 32) {
            const unsigned long long *t = S.top2 + ((size_t)r * 2 + (step & 1ull)) * 2;
            const double m1 = __longlong_as_double((long long)__ldcg(t));
            const double m2 = __longlong_as_double((long long)__ldcg(t + 1));
            b |= (sqrt(m1) + sqrt(m2) > S.nl_margin);
        }
        b = __any_sync(0xffffffffu, b);
        if (ll == 0) { sh.blocked[0] = b; sh.blocked[1] = 0; }
    }
    // bonded-term slots: slot s <-> the terms starting at item0 - 3 + s
    for (int slot = lt; slot < FAST_NT; slot += FAST_LOC) {
        const int titem = C.item0 - 3 + slot;
        int4 tm = make_int4(0, 1, 0, 0);
        int tb = -1;
        if (titem >= 0 && slot < C.nb + 3) {
            int trep, tk;
            fast_item(titem, nmp, &trep, &tk);
            tm = S.meta[tk];
            tb = titem;
        }
        sh.tmeta[slot] = tm;
        sh.tbead[slot] = tb;
    }
    cta_sync();
    const bool prof = FAST_PROF && S.prof != nullptr && blockIdx.x == 0 && lt == 0;
    long long l0 = 0, l1 = 0;
    int n = 0;
    if (!sh.blocked[0] && sh.ncand <= FAST_BPC && !sh.overflow)
    for (; n < C.nsteps; n++) {
        if (prof) l0 = clock64();
        // rebuild test on the displacements left by the previous step
        if (n > 0 && wwarp) {
            bool b = false;
            if (R == 1 || C.local) {
                double t1 = 0.0, t2 = 0.0;
                for (int c = C.local ? blockIdx.x + ll : ll; c < (C.local ? blockIdx.x + 1 : C.ncta); c += 32) {
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
        // bonded terms, each evaluated once (src/force_bond.F90, force_angl*.F90, force_dih_*.F90):
        // tasks [0, NT) = bond + angle of a slot (shared loads), [NT, 2 NT) = its dihedral
        const int nslot = C.nb + 3;                    // slots in use: terms starting at item0 - 3 .. item0 + nb - 1
        for (int task = lt; task < (C.dih_on ? 2 * nslot : nslot); task += FAST_LOC) {
            const int kind = task >= nslot, slot = task - kind * nslot;
            const int tb = sh.tbead[slot];
            if (tb < 0) continue;
            const int4 tm = sh.tmeta[slot];
            const bool tcirc = tm.w & MF_CIRC;
            const double4 *pos = S.pos[cur] + (tb - (tm.x + tm.z));     // first bead of the replica
            const int tk = tm.x + tm.z;
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
        // the bonded tables are done: help the BP warps with the survivors of the distance gate before the
        // Gaussians (a CTA with more formed base pairs than BP groups would otherwise need a second round)
        if (!C.no_loc_help) {
            while (*(volatile int *)&sh.gate_done[n & 1] != n + 1) { }
            __threadfence_block();
            fast_bp_eval(S, P, sh, S.pos[cur], C.item0, C.item0 + C.nb, n & 1, sh.nactive[n & 1]);
        }
        __threadfence_block();
        __syncwarp();
        bar_arrive(3, FAST_MAIN + FAST_LOC + FAST_BP);
        if (prof) l1 = clock64();
        // Gaussians of the next step (buffer parity of step + 2)
        if (n + 1 < C.nsteps) gauss_fill(step + 2ull);
        if (prof) { S.prof[8] += (unsigned long long)(l1 - l0); S.prof[9] += (unsigned long long)(clock64() - l1); }
        step++;
        cur ^= 1;
        cta_sync();
        if (sh.blocked[n & 1]) break;
        if (XCH) fast_remd_hook(S, P, C, sh, step, cur);
    }
    cta_sync();
    return n;
}

// Replica-local mode: the MAIN warps rebuild the lists of the CTA's replica in place -- pbc_wrap and
// the predicates of src/neighbor_list.F90:28-229 over all bead pairs (k_nl_rows' all-pairs sweep),
// the ele_cutoff refresh of :80-82, pair ids, xyz_move = 0 (src/job_md.F90:390) -- and the kernel
// carries on; nothing leaves the SM.
SIS_D void fast_local_rebuild(const DevSys &S, const FFParams &P, NlArgs NA, int rep, int cur)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nmp = S.nmp;
    const size_t rbase = (size_t)rep * nmp;
    double4 *pos = S.pos[cur] + rbase;
    if (P.flg_pbc)
        for (int i = tid; i < nmp; i += FAST_MAIN) {
            double4 p = pos[i];
            p.x = wrap1(p.x, P.box[0]); p.y = wrap1(p.y, P.box[1]); p.z = wrap1(p.z, P.box[2]);
            pos[i] = p;
        }
    if (tid == 0) {
        S.nbp_half[rep] = 0;
        atomicAdd(&S.misc->nrebuild, 1);
        if (NA.ele_cutoff_type == 2) S.rp[rep].ele_cutoff = NA.ele_cutoff_inp * S.rp[rep].lambdaD;
    }
    __threadfence_block();
    named_bar_sync(2, FAST_MAIN);
    NA.all_pairs = 2;
    for (int i = warp; i < nmp; i += FAST_MAIN_WARPS) nl_rows_bead_at(S, P, NA, rep, i, lane, pos, nullptr, nullptr);
    __threadfence_block();
    named_bar_sync(2, FAST_MAIN);
    for (int idx = tid; idx < nmp * S.KB; idx += FAST_MAIN) nl_bp_slot(S, rep, idx);
    if (tid < 4) S.top2[(size_t)rep * 4 + tid] = 0ull;
    if (tid < 2) S.top2_slots[tid * FAST_MAX_CTA + blockIdx.x] = make_double2(0.0, 0.0);
    __threadfence_block();
    if (tid == 0) (void)ld_acquire_u32(S.bar + 32);   // drop the SM's L1 copies of the old lists
    named_bar_sync(2, FAST_MAIN);
}


// ---- exchange step inside the replica-local kernel (src/job_md.F90:474-481,514-570) -----------------------
// energy_replica (src/energy_replica.F90:45-163) for the CTA's replica: total energy with its own
// parameters (bp energies as cached by the last force evaluation: flg_bp_energy) and with the
// partner label's parameters (bp term recomputed on the energy side, :72, :121) -- same thread
// decomposition and summation order as k_energy<8> / k_bp_eval<1>, so both paths give the same
// bits; nothing is written to bp_status / ene_bp (the reference restores them, :64-70,157-162).
SIS_D void fast_remd_energies(const DevSys &S, const FFParams &P, FastShared &sh, int rep, int cur, const RepParams &Rown,
                              const RepParams &Ralt, bool has_opp)
{
    const int tid = threadIdx.x, lane = tid & 7, k = tid >> 3, nmp = S.nmp, KB = S.KB;
    const size_t rbase = (size_t)rep * nmp, pbase = rbase * KB;
    const double4 *pos = S.pos[cur] + rbase;
    double e[ENE_N], ea[ENE_N];
#pragma unroll
    for (int t = 0; t < ENE_N; t++) { e[t] = 0.0; ea[t] = 0.0; }
    if (k < nmp) {
        const int4 mk4 = S.meta[k];
        const Vec3 xk = v3(pos[k]);
        const uint32_t *row = S.nbr + (rbase + k) * S.K;
        const int nn = S.nnbr[rbase + k];
        for (int q = lane; q < nn; q += 8) {
            const uint32_t ent = row[q];
            const int j = ent & NB_IDX;
            if (j < k) continue;
            const Vec3 v = pbc_vec_d(P, xk, v3(pos[j]));
            const double d2 = dot(v, v);
            if (ent & NB_WCA) { const double w = wca_energy(P, d2); e[ENE_EXV] += w; ea[ENE_EXV] += w; }
            if ((ent & NB_ELE) && P.flg_ele) { e[ENE_ELE] += dh_energy(P, Rown, d2); if (has_opp) ea[ENE_ELE] += dh_energy(P, Ralt, d2); }
        }
        if (lane < 3) {
            const int span = lane + 1;
            const bool exists = (mk4.w & MF_CIRC) ? (mk4.y > span) : (mk4.z + span < mk4.y);
            if (exists) {
                const Vec3 x1 = v3(pos[chain_nb(mk4, 1)]);
                double t = 0.0;
                if (lane == 0) { t = bond_energy(P, xk, x1); e[ENE_BOND] += t; ea[ENE_BOND] += t; }
                else if (lane == 1) { t = angl_energy(P, xk, x1, v3(pos[chain_nb(mk4, 2)])); e[ENE_ANGL] += t; ea[ENE_ANGL] += t; }
                else if (P.flg_dih_cos || P.flg_dih_exp) {
                    t = dih_energy(P, xk, x1, v3(pos[chain_nb(mk4, 2)]), v3(pos[chain_nb(mk4, 3)]));
                    e[ENE_DIHE] += t; ea[ENE_DIHE] += t;
                }
            }
        }
        const int nrow = S.nbp_row[rbase + k];
        for (int q = lane; q < nrow; q += 8) {
            const int2 ent = S.bprow[(rbase + k) * KB + q];
            if (ent.x <= k) continue;                   // owner = lower bead
            if (S.bp_status[pbase + ent.y]) e[ENE_BP] += S.ene_bp[pbase + ent.y];
            if (has_opp) {                              // k_bp_eval<1> with the partner temperature
                const int type = S.bp_type[pbase + ent.y];
                double dG = 1.0;
                bool active = true;
                if (P.bp_triplet) {
                    const double tK = P.temp_independent ? 0.0 : Ralt.tempK;   // src/energy_bp_triplet.F90:27-31
                    const double2 cf = S.bp_coef[pbase + ent.y];
                    dG = cf.x - tK * cf.y;
                    if (dG >= 0.0) active = false;
                }
                if (active) {
                    const int j = ent.x;
                    const int4 mj = S.meta[j];
                    Vec3 xs[6];
                    xs[0] = xk; xs[1] = v3(pos[j]);
                    xs[2] = v3(pos[chain_nb(mk4, -1)]); xs[3] = v3(pos[chain_nb(mj, -1)]);
                    xs[4] = v3(pos[chain_nb(mk4, +1)]); xs[5] = v3(pos[chain_nb(mj, +1)]);
                    double u;
                    if (bp_energy_u(P, P.bp[type], xs, &u)) {
                        const double ene = P.bp_triplet ? P.bp[type].U0 * dG * exp(-u) : P.bp[type].U0 * exp(-u);
                        if (ene <= P.bp_cutoff_energy) ea[ENE_BP] += ene;
                    }
                }
            }
        }
    }
    // reduction in the order of k_energy<8>: warp tree, the 8 warps of a 32-bead block in order, the blocks in order
    const int warp = tid >> 5;
    if (warp < 16) {
#pragma unroll
        for (int t = 1; t <= ENE_ELE; t++) {
            double v = e[t], va = ea[t];
            for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); va += __shfl_xor_sync(0xffffffffu, va, o); }
            if ((tid & 31) == 0) { sh.xred[warp][t] = v; sh.xred[warp][ENE_N + t] = va; }
        }
    }
    cta_sync();
    if (tid == 0) {
        const int nblk = (nmp + 31) / 32;
        double tot[2] = {0.0, 0.0};
        for (int side = 0; side < 2; side++) {
            double et[ENE_N];
            for (int t = 1; t <= ENE_ELE; t++) {
                double v = 0.0;
                for (int b = 0; b < nblk; b++) {
                    double vb = 0.0;
                    for (int w = 0; w < 8; w++) vb += sh.xred[8 * b + w][side * ENE_N + t];
                    v += vb;
                }
                et[t] = v;
            }
            // the only bias the replica-local mode admits: dual constant-force tweezers (src/energy_tweezers.F90:21-31)
            double etwz = 0.0;
            for (int t = 0; t < P.ntwz_DCF; t++) {
                const Vec3 v = pbc_vec_d(P, v3(pos[S.twz_pairs[2 * t]]), v3(pos[S.twz_pairs[2 * t + 1]]));
                etwz -= (side == 0 ? Rown.twz_force : Ralt.twz_force) *
                        (v.x * S.twz_dir[3 * t] + v.y * S.twz_dir[3 * t + 1] + v.z * S.twz_dir[3 * t + 2]);
            }
            double sum = 0.0;
            for (int t = 1; t < ENE_N; t++) sum += (t <= ENE_ELE ? et[t] : t == ENE_TWZ ? etwz : 0.0);
            tot[side] = sum;
        }
        sh.xE[0] = tot[0];
        sh.xE[1] = has_opp ? tot[1] : 0.0;
    }
    cta_sync();
}

// One exchange attempt after the step `step` (a multiple of nstep_rep_exchange): energies of the CTA's replica,
// a flag barrier over the CTAs of the launch (they are co-resident: cooperative launch), then EVERY CTA replays
// replica_exchange (src/replica_exchange.F90:44-154) on its private label tables from the energies of all
// replicas and the shared counter-based uniform stream -- no second exchange of data -- refreshes the parameters
// of its replica (src/job_md.F90:538-567) and rescales its velocities (scale_velo, :175-213).
SIS_D void fast_remd_exchange(const DevSys &S, const FFParams &P, FastShared &sh, int rep, int cur)
{
    const FastRemdState X = sh.xs;
    RemdDev &D = *S.remd;
    const int tid = threadIdx.x, n = D.n, R = D.R, nmp = S.nmp;
    const int grep = D.rank * R + rep;
    const int label_own = sh.x_r2l[grep];
    const int label_opp = D.pair_tb[(size_t)(X.idx_pair - 1) * n + label_own - 1];
    const bool has_opp = label_opp != 0 && label_opp <= n;
    const RepParams Rown = S.rp[rep];
    RepParams Ralt = Rown;
    if (has_opp) remd_apply_label(D, Ralt, D.lab_rp[label_opp - 1]);
    fast_remd_energies(S, P, sh, rep, cur, Rown, Ralt, has_opp);
    const unsigned long long round = D.xch_round0 + (unsigned long long)X.attempt;
    const size_t eoff = (size_t)(round & 1ull) * 2 * n;
    const double *E = D.xE_peer[D.rank] + eoff;
    if (tid == 0) {
        // the two energies into every rank's copy (peer stores over NVLink), then one arrival on every rank's counter
        for (int r = 0; r < D.nranks_x; r++) {
            double *Er = D.xE_peer[r] + eoff;
            Er[2 * grep] = sh.xE[0];
            Er[2 * grep + 1] = sh.xE[1];
        }
        __threadfence_system();
        for (int r = 0; r < D.nranks_x; r++) atomicAdd_system(D.xbar_peer[r], 1ull);
        const unsigned long long target = (round + 1ull) * (unsigned long long)n;     // one CTA per replica on the whole node
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(D.xbar_peer[D.rank]) < target)
            if (clock64() - t0 > 8000000000ll) { S.misc->max_force_flag |= 16; break; }   // ~4 s: a CTA / a rank is missing
        sh.x_newlab = label_own; sh.x_vscale = 1.0; sh.x_natt = 0; sh.x_nacc = 0;
    }
    cta_sync();
    // ---- replay of the decisions: warp 0, one lane per label; pair k in label order uses uniform k ----
    if (tid < 32) {
        const int *tb = D.pair_tb + (size_t)(X.idx_pair - 1) * n;
        int base = 0;
        for (int l0 = 0; l0 < n; l0 += 32) {
            const int l = l0 + tid;
            const int l_j = l < n ? tb[l] : 0;
            const bool is_pair = l < n && l_j >= l + 1;
            const unsigned m = __ballot_sync(0xffffffffu, is_pair);
            if (is_pair) {
                const int nrand = base + __popc(m & ((1u << tid) - 1u));
                const int l_i = l + 1;
                const double rnd = uniform_counter(D.seed, 0x7fffffffu, X.counter + (unsigned long long)nrand);
                const int rep_i = sh.x_l2r[l_i - 1], rep_j = sh.x_l2r[l_j - 1];
                double temp_m, temp_n;
                if (D.flg_temp) { temp_m = D.lab_temp[l_i - 1]; temp_n = D.lab_temp[l_j - 1]; }
                else temp_m = temp_n = D.tempK_default;
                const double pot_i_m = __ldcg(E + 2 * (rep_i - 1)), pot_i_n = __ldcg(E + 2 * (rep_i - 1) + 1);
                const double pot_j_n = __ldcg(E + 2 * (rep_j - 1)), pot_j_m = __ldcg(E + 2 * (rep_j - 1) + 1);
                const double delta = ((pot_j_m - pot_i_m) / temp_m - (pot_j_n - pot_i_n) / temp_n) / SIS_BOLTZ_KCAL_MOL;
                atomicAdd(&sh.x_natt, 1);
                if (delta < 0.0 || rnd <= exp(-delta)) {
                    atomicAdd(&sh.x_nacc, 1);
                    sh.x_r2l[rep_i - 1] = l_j; sh.x_r2l[rep_j - 1] = l_i;
                    sh.x_l2r[l_i - 1] = rep_j; sh.x_l2r[l_j - 1] = rep_i;
                    const double sf = D.flg_temp ? sqrt(temp_n / temp_m) : 1.0;   // scale_velo
                    if (rep_i - 1 == grep) { sh.x_newlab = l_j; sh.x_vscale = sf; }
                    if (rep_j - 1 == grep) { sh.x_newlab = l_i; sh.x_vscale = 1.0 / sf; }
                }
            }
            base += __popc(m);
        }
    }
    cta_sync();
    // parameter refresh of the replica (src/job_md.F90:538-567); the caller rescales the velocities it holds
    if (tid == 0) {
        FastRemdState Y = X;
        Y.natt += sh.x_natt; Y.nacc += sh.x_nacc;
        Y.idx_pair = (X.idx_pair == D.ncol) ? 1 : X.idx_pair + 1;       // set_forward
        Y.counter += (unsigned long long)((n + 1) / 2);                 // the reference draws ceil(n / 2) numbers up front
        Y.attempt++;
        sh.xs = Y;
        RepParams rp = S.rp[rep];
        remd_apply_label(D, rp, D.lab_rp[sh.x_newlab - 1]);
        S.rp[rep] = rp;
    }
    __threadfence_block();
    cta_sync();
}

// Called by every thread of the CTA at the end of a completed step: on a multiple of nstep_rep_exchange the
// exchange attempt, then the state the roles keep on chip follows the new label -- velocities (scale_velo,
// src/replica_exchange.F90:175-213), noise amplitude and Debye-Hueckel coefficients of the owned beads
// (set_md_coef / set_ele, src/job_md.F90:538-567; ele_cutoff stays: only neighbor_list refreshes it), dG of
// the cached base-pair candidates (src/force_bp_triplet.F90:77)
SIS_D void fast_remd_hook(const DevSys &S, const FFParams &P, const FastCtx &C, FastShared &sh, unsigned long long step, int cur)
{
    if (!C.xch || step % (unsigned long long)S.remd->nex != 0ull) return;
    const int tid = threadIdx.x, rep = C.item0 / C.nmp;
    fast_remd_exchange(S, P, sh, rep, cur);
    const RepParams &Rp = S.rp[rep];
    if (tid < FAST_MAIN) {
        const double vs = sh.x_vscale;
        for (int s = tid; s < C.nb; s += FAST_MAIN) {
            FastBead &B = sh.bead[s];
            B.v[0] *= vs; B.v[1] *= vs; B.v[2] *= vs;
            B.noise = S.coef[B.k].w * Rp.sqrt_tempK;
            B.ele_coef = Rp.ele_coef; B.inv_lambdaD = Rp.inv_lambdaD;
        }
    } else if (tid < FAST_MAIN + FAST_BP) {
        if (P.bp_triplet)
            for (int i = tid - FAST_MAIN; i < min(sh.ncand, FAST_BPC); i += FAST_BP) {
                const double2 cf = S.bp_coef[sh.cand[i].pb];
                sh.cand[i].dG = cf.x - Rp.tempK * cf.y;
            }
    }
    __threadfence_block();
    cta_sync();
}

// XCH: the exchange step of REMD inside the kernel (replica-local mode only); a separate instantiation so that the
// plain kernels do not carry its registers
template <int U, bool LOCAL, bool XCH = false>
__global__ void __launch_bounds__(FAST_THREADS, 1) k_md_fast(const __grid_constant__ DevSys S, const __grid_constant__ FFParams P, StepArgs A,
                                                              long long target, NlArgs NA)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FastShared &sh = *reinterpret_cast<FastShared *>(smem_raw);
    uint32_t *s_rows = reinterpret_cast<uint32_t *>(smem_raw + ((sizeof(FastShared) + 15) & ~size_t(15)));
    const DevCtl ctl = S.ctl[0];
    // steps to run: up to `target`; nothing on lists a device-side rebuild could not fit (the host
    // grows them first).  Uniform over the grid: ctl and misc were written by earlier kernels.
    const int nsteps = md_steps_to_run(S, ctl, target);
    if (nsteps <= 0) return;
    const long long k_entry = clock64();
    FastCtx C;
    C.R = S.nrep; C.nmp = S.nmp; C.K = S.K; C.nitems = S.nrep * S.nmp; C.ncta = gridDim.x;
    C.item0 = S.cta_item0[blockIdx.x]; C.nb = S.cta_item0[blockIdx.x + 1] - C.item0;
    C.bar_pos = S.bar + 32;
    C.dih_on = P.flg_dih_cos || P.flg_dih_exp;
    C.local = LOCAL;
    C.after_rebuild = false; C.seed = A.seed; C.max_bp_per_nt = A.max_bp_per_nt; C.no_loc_help = S.dbg_no_loc_help != 0;
    const int tid = threadIdx.x;
    // A launch is one segment (prologue, step loop) unless the CTA rebuilds its own lists (LOCAL):
    // then a segment ends at the rebuild test and the next one starts on the new lists.  With the
    // exchange step inside the kernel (S.remd) a segment also ends at every multiple of
    // nstep_rep_exchange: the exchange works on the state in global memory and the next segment's
    // prologue picks up the new parameters and the rescaled velocities.
    const bool xch = XCH && LOCAL && S.remd != nullptr;
    C.xch = xch;
    if (xch) {
        const RemdDev &D = *S.remd;
        if (tid == 0) {
            FastRemdState X;
            X.idx_pair = D.idx_pair; X.counter = D.exchange_counter; X.attempt = 0; X.natt = 0; X.nacc = 0;
            sh.xs = X;
        }
        for (int i = tid; i < D.n; i += FAST_THREADS) { sh.x_r2l[i] = D.rep2lab[i]; sh.x_l2r[i] = D.lab2rep[i]; }
        // (cta_sync before the first use: every role's prologue ends with one)
    }
    int done = 0;
    for (;;) {
        C.nsteps = nsteps - done;
        C.step0 = ctl.step + (unsigned long long)done;
        C.cur0 = ctl.cur ^ (done & 1);
        int n;
        if (tid >= FAST_MAIN + FAST_BP) n = fast_loc_role<XCH>(S, P, A, C, sh);
        else if (tid >= FAST_MAIN) n = fast_bp_role<LOCAL, XCH>(S, P, C, sh);
        else n = fast_main_role<U, XCH>(S, P, A, C, sh, s_rows);
        done += n;
        if (!LOCAL || done >= nsteps) break;
        if (sh.ncand > FAST_BPC || sh.overflow) break;     // tables too small (the host checked: cannot happen)
        __threadfence_block();
        cta_sync();                                        // velocities, beta, xyz_move are back in global memory
        if (tid < FAST_MAIN) fast_local_rebuild(S, P, NA, C.item0 / C.nmp, C.cur0 ^ (n & 1));
        C.after_rebuild = true;
        cta_sync();
    }
    if (xch && blockIdx.x == 0 && tid == 0) {              // tables and counters of the exchange back for the host / the next launch
        RemdDev &D = *S.remd;
        const FastRemdState X = sh.xs;
        for (int i = 0; i < D.n; i++) { D.rep2lab[i] = sh.x_r2l[i]; D.lab2rep[i] = sh.x_l2r[i]; }
        D.idx_pair = X.idx_pair; D.exchange_counter = X.counter;
        D.n_attempt += X.natt; D.n_accept += X.nacc;
        D.xch_round0 += (unsigned long long)X.attempt;
    }
    if (tid >= FAST_MAIN) return;
    if (FAST_PROF && S.prof != nullptr && blockIdx.x == 0 && tid == 0) {   // whole-launch clocks: prologue share = total - steps * per-step
        S.prof[10] += (unsigned long long)(clock64() - k_entry);
        S.prof[11] += 1ull;
    }
    // ---- bookkeeping for the host and the other kernels (warp 0) ------------------------------------
    const unsigned long long step = ctl.step + (unsigned long long)done;
    const int cur = ctl.cur ^ (done & 1);
    if (LOCAL) {
        if (tid == 0 && done > 0) {                        // this replica's displacement top-2 (src/job_md.F90:376-388)
            const double2 s = __ldcg(S.top2_slots + (step & 1ull) * FAST_MAX_CTA + blockIdx.x);
            unsigned long long *t = S.top2 + ((size_t)(C.item0 / C.nmp) * 2 + (step & 1ull)) * 2;
            t[0] = (unsigned long long)__double_as_longlong(s.x); t[1] = (unsigned long long)__double_as_longlong(s.y);
            t = S.top2 + ((size_t)(C.item0 / C.nmp) * 2 + ((step + 1ull) & 1ull)) * 2;
            t[0] = 0ull; t[1] = 0ull;
        }
    } else if (blockIdx.x == 0 && tid < 32) {
        if (C.R == 1 && done > 0) {
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
    }
    if (LOCAL) {
        // without grid barriers a CTA may finish before another one has read ctl: the last CTA to
        // finish publishes the new step (S.bar[0] is zeroed by the host before every launch)
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(S.bar, 1u) == gridDim.x - 1) {
                DevCtl a; a.step = step; a.cur = cur; a.blocked = 0;
                S.ctl[0] = a;
            }
        }
    } else if (blockIdx.x == 0 && tid == 0) {
        DevCtl a; a.step = step; a.cur = cur; a.blocked = 0;
        S.ctl[0] = a;
    }
}
