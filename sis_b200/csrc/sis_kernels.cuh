// sis_kernels.cuh -- sm_100a kernels of the SIS hot path.
//
// Layout (DESIGN.md "HBM layout"): positions are double4 AoS (one 32-byte sector per gather),
// double-buffered so that one launch computes forces from buffer `cur` and writes the
// integrated positions to buffer `cur^1`; velocities / beta / displacement are [bead][3]
// doubles; neighbour rows are fixed-capacity (ELL) uint32 rows per bead holding the FULL
// (both directions) list with two flag bits (WCA-listed, DH-listed), so a group of G lanes
// owns one bead, streams its row coalesced and reduces with shuffles -- no atomics.
#pragma once

#include "sis_device.cuh"

#ifdef __CUDACC__
#include <cooperative_groups.h>
#endif

// Step bookkeeping lives on the device so that a whole chunk of steps can be replayed as one
// CUDA graph (or one persistent launch) without host-supplied per-step arguments.
// S.ctl[0] ("A") is read by the force-phase kernels and written only by the integrate kernel;
// S.ctl[1] ("B") is read by the integrate kernel and written only by the force phase -- no kernel
// ever writes the copy its own blocks read, so no "last block" hand-off is needed.
struct DevCtl {
    unsigned long long step;   // completed MD steps (job_md's istep)
    int cur;                   // position buffer that holds the current coordinates
    int blocked;               // B only: some replica is past the list margin, step not executed
};
struct DevMisc {
    int nl_overflow, bp_overflow;   // longest rows seen by the last list build
    int max_force_flag;             // src/job_md.F90:403-407 (forces > 100 warning); bit1: stage abort
    int lists_invalid;              // device-side rebuild: a row / the bp half list overflowed its capacity,
                                    // the persistent kernels do not run until the host has grown the lists
    int nrebuild;                   // replica list rebuilds done by the device-side chain (k_nl_plan)
    int half_overflow;              // largest bp half-list count seen by k_nl_partition
    int inject_invalid_at;          // test hook (option "inject_invalid"): k_nl_plan marks the lists invalid at this rebuild count
    int pad;
    long long stop_step;            // k_md_big / k_nl_plan inside replayed graphs: no step, no rebuild once this step is reached
};

// one contiguous device block the host mirrors with a single copy: header + top2[nrep][2][2]
struct DevSyncHdr {
    DevCtl ctl[2];
    DevMisc misc;
};

struct GridInfo {
    double origin[3], inv_cell[3];
    int n[3], ncell, pbc;
};

struct DevSys {
    int nmp, nrep, K, KB;
    int nblk;                     // integrate-kernel blocks per replica (64 beads each)
    int nblk_ene;                 // energy-kernel blocks per replica
    // topology
    const int4 *meta;
    const int *seq;
    const double4 *coef;          // a, sqrt(b)h/m, sqrt(b)h (0 if frozen), noise0
    const double *mass;
    const int *hash;              // bp3 hash (4096), 1-based values, 0 = none
    BpRestrict rs;                // ct / bpseq / bpl restriction of the bp map (n_ss < 0: none)
    const double *bp3_dH, *bp3_dS;  // 1-based (index 0 unused)
    // state [rep][nmp]
    double4 *pos[2];
    double *vel, *acc, *move;     // [rep][nmp][3]
    double *force;                // [rep][nmp][3]
    double *f_pair, *f_loc, *f_bp;  // partial forces of the three task families [rep][nmp][3]
    // lists
    uint32_t *nbr;                // [rep][nmp][K]
    int *nnbr;                    // [rep][nmp]
    int2 *bprow;                  // [rep][nmp][KB]  {partner, ibp}
    int *nbp_row;                 // [rep][nmp]
    signed char *bp_type;         // [rep][nmp*KB] by ibp = owner*KB + slot
    double2 *bp_coef;             // dH, dS*1e-3
    int *bp_status, *bp_status_MC;
    double *ene_bp;
    int *bp_half;                 // [rep][bp_half_cap] compact list of owner ibp
    int *nbp_half;                // [rep]
    int bp_half_cap;
    int *nt_excess;               // [rep][nmp]
    unsigned long long *mc_counter;  // [rep]
    // cell grid [rep][...]
    int *cell_of, *cell_count, *cell_start, *cell_cursor, *cell_beads;
    int *cell_sorted;             // cell_beads with every cell in ascending bead order (k_nl_rank)
    int ncell_cap;
    GridInfo *grid;               // the cell grid of the current build (host- or k_nl_plan-written)
    int dbg_bp_groups;            // option "bp_groups" = 1: k_md_fast evaluates the base-pair survivors with 8-lane role groups (round-1 form)
    int dbg_no_loc_help;          // option "loc_help" = 0: the LOC warps of k_md_fast do not join the base-pair evaluation
    struct RemdDev *remd;         // non-null: k_md_fast<U, true> runs the exchange step itself every remd->nex steps
    // parameters
    RepParams *rp;
    DevCtl *ctl;                  // [2]: A, B
    DevMisc *misc;
    // displacement top-2 (src/job_md.F90:376-388) as order-preserving bit patterns of |move|^2:
    // [rep][slot = istep & 1][0 = max, 1 = second max]
    unsigned long long *top2;
    double nl_margin;             // 1e300 with the cutoff-free lists (no rebuild trigger)
    unsigned long long *prof;     // optional phase clocks of the persistent kernel (block 0), or null
    unsigned int *bar;            // grid-barrier counters of k_md_fused: [0] bp done, [32] positions written
    double2 *top2_slots;          // k_md_fast: per-CTA (max, 2nd max) |xyz_move|^2, [2][FAST_MAX_CTA]
    int *cta_item0;               // k_md_fast: first flattened (replica, bead) item of each CTA, [ncta + 1]
    double *ene_partial;          // [rep][nblk][ENE_N]
    unsigned int *ene_counter;    // [rep]
    double *ene_out;              // [rep][ENE_N]
    // biases
    const int *twz_pairs;         // [ntwz][2] 0-based
    const double *twz_dir;        // [ntwz][3]
    const int *twzfr_pairs;
    const double *twzfr_k, *twzfr_init, *twzfr_velo;
    const int *rest_id;           // [nrest][3] 0-based i, j, type
    const double *rest_rcut, *rest_para;
};

struct StepArgs {
    unsigned long long seed;
    unsigned term_mask;   // bit k = energies index k
    int use_mask;         // 0 all pairs; 1 only bp_status_MC (nstep_bp_MC > 0); 2 only bp_status
    int only_rep;         // >= 0: single replica (per-irep API); -1: all (grid.y)
    int integrate;        // 1: MD step (job_md loop body); 0: force() only
    int nblk_pair, nblk_loc, nblk_bp;   // per-replica block counts of the three force roles
    int max_bp_per_nt;                  // Monte-Carlo pruning inside k_md_fast (replica-local mode, nstep_bp_MC > 0)
};

// sum over the G lanes that own one bead.  Host builds (tests/csrc/hostcheck.cu) run the same
// force code with G = 1, where the group is a single "lane".
template <int G>
SIS_HD double group_sum(double v)
{
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
#else
    static_assert(G >= 1, "group size");
#endif
    return v;
}

// ---- force tasks --------------------------------------------------------------------------
// The step is three concurrent task families (DESIGN.md "step graph"):
//   pair   : GP lanes own bead k and stream its FULL neighbour row (WCA + DH), plain store
//   local  : 8 lanes own bead k: the 9 bonded term instances k takes part in + per-bead biases
//   bp     : one thread per base-pair candidate (half list), 6-bead scatter with fp64 atomics
// followed by the integrate kernel that sums the three partial forces.

// src/force_wca.F90:34-66, src/force_ele_DH.F90:39-81 in gather form (lane-partial sum)
template <int G>
SIS_HD Vec3 pair_force(const DevSys &S, const FFParams &P, const RepParams &R, const double4 *__restrict__ pos,
                       int rep, int k, int lane, unsigned term_mask)
{
    Vec3 F = mk(0.0, 0.0, 0.0);
    if (!(term_mask & ((1u << ENE_EXV) | (1u << ENE_ELE)))) return F;
    const int nmp = S.nmp;
    const Vec3 xk = v3(pos[k]);
    const uint32_t *row = S.nbr + ((size_t)rep * nmp + k) * S.K;
    const int nn = S.nnbr[(size_t)rep * nmp + k];
    const double cut2 = R.ele_cutoff * R.ele_cutoff;
    const bool do_wca = term_mask & (1u << ENE_EXV);
    const bool do_ele = (term_mask & (1u << ENE_ELE)) && P.flg_ele;
    // batches of 4 entries per lane: the 4 row loads, then the 4 position gathers, are issued
    // back to back so that a bead costs ~2 dependent L2 round trips instead of 2 per entry
    for (int e0 = lane; e0 < nn; e0 += 4 * G) {
        uint32_t ent[4];
        Vec3 xj[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + u * G;
            ent[u] = e < nn ? row[e] : (uint32_t)k;   // padding: self index, no list flags
        }
#pragma unroll
        for (int u = 0; u < 4; u++) xj[u] = v3(pos[ent[u] & NB_IDX]);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            Vec3 v = xk - xj[u];
            if (ent[u] & NB_IMG) {    // only pairs that can cross half a box before the next rebuild (nl_needs_image)
                v.x = minimg1(v.x, P.box[0], P.box_half[0]);
                v.y = minimg1(v.y, P.box[1], P.box_half[1]);
                v.z = minimg1(v.z, P.box[2], P.box_half[2]);
            }
            const double d2 = dot(v, v);
            double s = 0.0;
            const bool w_on = do_wca && (ent[u] & NB_WCA) && d2 < P.wca_sigma2, e_on = do_ele && (ent[u] & NB_ELE) && d2 <= cut2;
            if (P.mixed) { if (w_on || e_on) s = pair_scalar_f32(P, w_on, e_on, d2, R.ele_coef, R.inv_lambdaD); }
            else {
                if (w_on) s += wca_scalar(P, d2);
                if (e_on) s += dh_scalar(R.ele_coef, R.inv_lambdaD, d2);
            }
            F += s * v;
        }
    }
    return F;
}

// per-bead biases in gather form (src/force_stage.F90, force_tweezers.F90, force_rg.F90,
// force_restraint.F90): the force these terms put on bead k
SIS_HD Vec3 bias_force(const DevSys &S, const FFParams &P, const RepParams &R, const double4 *__restrict__ pos,
                       int k, Vec3 xk, unsigned term_mask, unsigned long long istep)
{
    const int nmp = S.nmp;
    Vec3 F = mk(0.0, 0.0, 0.0);
    if ((term_mask & (1u << ENE_STAGE)) && P.flg_stage) {   // src/force_stage.F90:17-37
        double z = xk.z;
        if (z < SIS_SMALL_VALUE) z = SIS_SMALL_VALUE;
        z = P.stage_sigma / z;
        F.z += (24.0 * P.stage_eps / P.stage_sigma) * (2 * powi(z, 13) - powi(z, 7));
    }
    if (term_mask & (1u << ENE_TWZ)) {                      // src/force_tweezers.F90:27-61
        for (int t = 0; t < P.ntwz_DCF; t++) {
            Vec3 f = R.twz_force * mk(S.twz_dir[3 * t], S.twz_dir[3 * t + 1], S.twz_dir[3 * t + 2]);
            if (S.twz_pairs[2 * t] == k) F += f;
            if (S.twz_pairs[2 * t + 1] == k) F -= f;
        }
        for (int t = 0; t < P.ntwz_FR; t++)
            for (int s = 0; s < 2; s++)
                if (S.twzfr_pairs[2 * t + s] == k) {
                    const double *ini = S.twzfr_init + 6 * t + 3 * s, *vel = S.twzfr_velo + 6 * t + 3 * s;
                    double st = (double)istep;
                    Vec3 p = mk(ini[0] + vel[0] * st, ini[1] + vel[1] * st, ini[2] + vel[2] * st);
                    F += S.twzfr_k[2 * t + s] * pbc_vec_d(P, p, xk);
                }
    }
    if ((term_mask & (1u << ENE_RG)) && P.flg_bias_rg) {    // src/force_rg.F90:18-39
        double rg = R.rg_val;
        if (!(P.bias_rg_pott == 2 && rg <= R.bias_rg_0)) {
            double factor = -P.bias_rg_k * (rg - R.bias_rg_0) / (rg * nmp);
            F += factor * (xk - mk(R.rg_cm[0], R.rg_cm[1], R.rg_cm[2]));
        }
    }
    if (term_mask & (1u << ENE_REST)) {                     // src/force_restraint.F90:19-41
        for (int t = 0; t < P.nrest; t++) {
            int ia = S.rest_id[3 * t], ib = S.rest_id[3 * t + 1], ty = S.rest_id[3 * t + 2];
            if (ia != k && ib != k) continue;
            Vec3 v = pbc_vec_d(P, v3(pos[ia]), v3(pos[ib]));
            double r = sqrt(dot(v, v));
            if (r >= S.rest_rcut[t]) continue;
            double eps = S.rest_para[3 * t], d = S.rest_para[3 * t + 1], s = S.rest_para[3 * t + 2];
            double th = tanh((r - d) / s);
            double coef = 0.5 * eps / s * (1.0 - th * th) / r;
            if (ia == k) F -= coef * v;
            else if (ty == 1) F += coef * v;
        }
    }
    return F;
}

// bonded terms (src/force_bond.F90, force_angl*.F90, force_dih_*.F90), one term instance per
// lane: 2 bonds, 3 angles, 4 dihedrals; plus the per-bead biases on one lane
// (src/force_stage.F90, force_tweezers.F90, force_rg.F90, force_restraint.F90)
template <int G>
SIS_HD Vec3 local_force(const DevSys &S, const FFParams &P, const RepParams &R, const double4 *__restrict__ pos,
                        int rep, int k, int lane, unsigned term_mask, unsigned long long istep)
{
    (void)rep;
    const int nmp = S.nmp;
    const int4 mk4 = S.meta[k];
    const Vec3 xk = v3(pos[k]);
    Vec3 F = mk(0.0, 0.0, 0.0);
    for (int inst = lane; inst < 9; inst += G) {
        int kind, p;   // kind 0 bond, 1 angle, 2 dihedral; p = position of k in the term
        if (inst < 2) { kind = 0; p = inst; }
        else if (inst < 5) { kind = 1; p = inst - 2; }
        else { kind = 2; p = inst - 5; }
        const int span = kind + 1;   // term covers beads a .. a+span
        const bool on = (kind == 0) ? (term_mask & (1u << ENE_BOND))
                        : (kind == 1) ? (term_mask & (1u << ENE_ANGL))
                                      : ((term_mask & (1u << ENE_DIHE)) && (P.flg_dih_cos || P.flg_dih_exp));
        // first bead a = k - p; the term exists iff a and a+span exist (circRNA: always, n > span)
        bool exists;
        if (mk4.w & MF_CIRC) exists = mk4.y > span;
        else exists = (mk4.z - p >= 0) && (mk4.z - p + span < mk4.y);
        if (on && exists) {
            Vec3 x[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (q <= span) x[q] = (q == p) ? xk : v3(pos[chain_nb(mk4, q - p)]);
            Vec3 f;
            if (kind == 0) f = bond_force(P, x[0], x[1], p);
            else if (kind == 1) f = P.flg_angl_ReB ? angl_ReB_force(P, x[0], x[1], x[2], p)
                                                  : angl_force(P, x[0], x[1], x[2], p);
            else f = dih_force(P, x[0], x[1], x[2], x[3], p);
            F += f;
        }
    }
    if (lane == (G > 9 ? 9 : 0)) F += bias_force(S, P, R, pos, k, xk, term_mask, istep);
    return F;
}

// one base-pair candidate (owner bead k, row slot e): src/force_bp.F90:50-204,
// force_bp_triplet.F90:62-240.  Returns how many bead forces to scatter (0, 2 for the Bias_SS
// branch, 6) into beads[] / f[]; writes bp_status / ene_bp when write_bp.
// use_mask: 0 all pairs; 1 only bp_status_MC; 2 only bp_status (after the MC pruning)
SIS_HD int bp_pair_force(const DevSys &S, const FFParams &P, const RepParams &R, const double4 *__restrict__ pos,
                         int rep, int k, int e, int use_mask, int write_bp, int *beads, Vec3 *f)
{
    const int nmp = S.nmp;
    const size_t rbase = (size_t)rep * nmp, pbase = rbase * S.KB;
    const int2 ent = S.bprow[(rbase + k) * S.KB + e];
    const int j = ent.x, ibp = ent.y;
    bool active = true;
    if (use_mask == 1) active = S.bp_status_MC[pbase + ibp] != 0;
    else if (use_mask == 2) active = S.bp_status[pbase + ibp] != 0;
    const int type = S.bp_type[pbase + ibp];
    double dG = 1.0;
    if (P.bp_triplet) {
        const double2 cf = S.bp_coef[pbase + ibp];
        dG = cf.x - R.tempK * cf.y;     // src/force_bp_triplet.F90:77
        if (dG >= 0.0) active = false;
    }
    int nout = 0, st = 0;
    double en = 0.0;
    if (active) {
        const int4 mi = S.meta[k], mj = S.meta[j];
        beads[0] = k; beads[1] = j;
        beads[2] = chain_nb(mi, -1); beads[3] = chain_nb(mj, -1);
        beads[4] = chain_nb(mi, +1); beads[5] = chain_nb(mj, +1);
        Vec3 xs[6];
#pragma unroll
        for (int s = 0; s < 6; s++) xs[s] = v3(pos[beads[s]]);
        double u, a12;
        Vec3 v12;
        if (bp_force_all(P, P.bp[type], xs, &u, f, &a12, &v12)) {
            const double ene = P.bp_triplet ? P.bp[type].U0 * dG * exp(-u) : P.bp[type].U0 * exp(-u);
            if (ene <= P.bp_cutoff_energy) {
                st = 1; en = ene; nout = 6;
#pragma unroll
                for (int s = 0; s < 6; s++) f[s] = ene * f[s];
            }
        } else if (P.bp_triplet && P.flg_bias_ss) {   // src/force_bp_triplet.F90:85-93
            const Vec3 t = (P.bias_ss_force / a12) * v12;
            f[0] = -t; f[1] = t; nout = 2;
        }
    }
    if (write_bp) {
        S.bp_status[pbase + ibp] = st;
        S.ene_bp[pbase + ibp] = en;
    }
    return nout;
}

// merge two (max, second max) pairs
SIS_D void top2_merge(double &a1, double &a2, double b1, double b2)
{
    double m1 = fmax(a1, b1);
    double m2 = fmax(fmin(a1, b1), fmax(a2, b2));
    a1 = m1; a2 = m2;
}

// the rebuild test of src/job_md.F90:388 on the displacements left by step `step`:
// sqrt(d2max) + sqrt(d2max_2nd) > nl_margin for any replica of this launch
SIS_D bool step_blocked(const DevSys &S, unsigned long long step, int only_rep)
{
    const int r0 = only_rep >= 0 ? only_rep : 0, r1 = only_rep >= 0 ? only_rep + 1 : S.nrep;
    bool blocked = false;
    for (int r = r0; r < r1; r++) {
        const unsigned long long *t = S.top2 + ((size_t)r * 2 + (step & 1ull)) * 2;
        const double m1 = __longlong_as_double((long long)t[0]), m2 = __longlong_as_double((long long)t[1]);
        blocked |= (sqrt(m1) + sqrt(m2) > S.nl_margin);
    }
    return blocked;
}

// ---- force phase, one block's share -----------------------------------------------------------
// pair role: GP lanes per bead, beads [blk*256/GP, ...)
template <int GP>
SIS_D void phase_pair(const DevSys &S, const FFParams &P, const StepArgs &A, const double4 *pos, int rep, int blk)
{
    constexpr int BPB = 256 / GP;
    const int nmp = S.nmp;
    const int lane = threadIdx.x % GP;
    const int kraw = blk * BPB + threadIdx.x / GP;
    const bool valid = kraw < nmp;
    const int k = valid ? kraw : nmp - 1;   // tail groups shadow the last bead (no writes)
    const RepParams &R = S.rp[rep];
    Vec3 F = pair_force<GP>(S, P, R, pos, rep, k, lane, A.term_mask);
    F.x = group_sum<GP>(F.x); F.y = group_sum<GP>(F.y); F.z = group_sum<GP>(F.z);
    if (valid && lane < 3) S.f_pair[((size_t)rep * nmp + k) * 3 + lane] = lane == 0 ? F.x : lane == 1 ? F.y : F.z;
}
// local role: 8 lanes per bead
SIS_D void phase_local(const DevSys &S, const FFParams &P, const StepArgs &A, const double4 *pos, int rep, int blk,
                       unsigned long long istep)
{
    constexpr int GL = 8, BPB = 256 / GL;
    const int nmp = S.nmp;
    const int lane = threadIdx.x % GL;
    const int kraw = blk * BPB + threadIdx.x / GL;
    const bool valid = kraw < nmp;
    const int k = valid ? kraw : nmp - 1;
    const RepParams &R = S.rp[rep];
    Vec3 F = local_force<GL>(S, P, R, pos, rep, k, lane, A.term_mask, istep);
    F.x = group_sum<GL>(F.x); F.y = group_sum<GL>(F.y); F.z = group_sum<GL>(F.z);
    if (valid && lane < 3) S.f_loc[((size_t)rep * nmp + k) * 3 + lane] = lane == 0 ? F.x : lane == 1 ? F.y : F.z;
}
// bp role: one thread per half-list candidate t
SIS_D void phase_bp(const DevSys &S, const FFParams &P, const StepArgs &A, const double4 *pos, int rep, int t)
{
    if (t >= S.nbp_half[rep]) return;
    const int nmp = S.nmp;
    const int ibp = S.bp_half[(size_t)rep * S.bp_half_cap + t];
    int beads[6];
    Vec3 f[6];
    const int n = bp_pair_force(S, P, S.rp[rep], pos, rep, ibp / S.KB, ibp % S.KB, A.use_mask, A.use_mask != 2, beads, f);
    double *fb = S.f_bp + (size_t)rep * nmp * 3;
    for (int s = 0; s < n; s++) {
        atomicAdd(fb + 3 * beads[s], f[s].x);
        atomicAdd(fb + 3 * beads[s] + 1, f[s].y);
        atomicAdd(fb + 3 * beads[s] + 2, f[s].z);
    }
}
// integrate (src/job_md.F90:409-456) for the 64 beads of block `blk`; 4 lanes per bead, lanes 0..2
// own x, y, z.  Consumes and clears the bp scatter buffer, pushes the block's displacement top-2
// into the slot of `istep`.
SIS_D void phase_integrate(const DevSys &S, const FFParams &P, const StepArgs &A, int cur, int rep, int blk,
                           unsigned long long istep)
{
    const int lane = threadIdx.x % 4;
    const int k = blk * 64 + threadIdx.x / 4;
    const int nmp = S.nmp;
    const bool valid = k < nmp;
    const double4 *pos = S.pos[cur] + (size_t)rep * nmp;
    double xnew = 0.0, mv = 0.0;
    if (valid && lane < 3) {
        const size_t o = ((size_t)rep * nmp + k) * 3 + lane;
        const double f = S.f_pair[o] + S.f_loc[o] + S.f_bp[o];
        S.f_bp[o] = 0.0;
        if (!A.integrate) {
            S.force[o] = f;
        } else {
            const RepParams &R = S.rp[rep];
            const double4 xo = pos[k];
            const double4 c = S.coef[k];
            const double x = lane == 0 ? xo.x : lane == 1 ? xo.y : xo.z;
            const double g = gauss_component(A.seed, (uint32_t)R.stream, istep, (uint32_t)k, lane);
            double dx;
            if (P.integrator == 1) {   // GJF-2GJ
                const double accels_pre = (c.w * R.sqrt_tempK) * g;
                const double v = c.x * S.vel[o] + c.y * f + (S.acc[o] + accels_pre);
                S.vel[o] = v;
                S.acc[o] = accels_pre;
                dx = c.z * v;
            } else {                   // Ermak-McCammon
                dx = c.x * f + (c.w * R.sqrt_tempK) * g;
            }
            xnew = x + dx;
            mv = S.move[o] + dx;
            S.move[o] = mv;
            if (f > 100.0) S.misc->max_force_flag |= 1;
        }
    }
    if (!A.integrate) return;
    const double nx = __shfl_sync(0xffffffffu, xnew, 0, 4);
    const double ny = __shfl_sync(0xffffffffu, xnew, 1, 4);
    const double nz = __shfl_sync(0xffffffffu, xnew, 2, 4);
    const double s0 = __shfl_sync(0xffffffffu, mv, 0, 4);
    const double s1 = __shfl_sync(0xffffffffu, mv, 1, 4);
    const double s2 = __shfl_sync(0xffffffffu, mv, 2, 4);
    const double d2 = s0 * s0 + s1 * s1 + s2 * s2;
    if (valid && lane == 0) S.pos[cur ^ 1][(size_t)rep * nmp + k] = make_double4(nx, ny, nz, 0.0);
    // warp top-2, then two order-preserving integer atomic maxima per warp:
    //   old = atomicMax(m1, a1);  atomicMax(m2, max(min(old, a1), a2))
    // leaves (m1, m2) = the two largest |move|^2 of the replica whatever the arrival order
    double t1 = (valid && lane == 0) ? d2 : 0.0, t2 = 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double b1 = __shfl_xor_sync(0xffffffffu, t1, o), b2 = __shfl_xor_sync(0xffffffffu, t2, o);
        top2_merge(t1, t2, b1, b2);
    }
    if (threadIdx.x % 32 == 0 && t1 > 0.0) {
        unsigned long long *t = S.top2 + ((size_t)rep * 2 + (istep & 1ull)) * 2;
        const unsigned long long old = atomicMax(t, (unsigned long long)__double_as_longlong(t1));
        const double oldd = __longlong_as_double((long long)old);
        const double c2 = fmax(fmin(oldd, t1), t2);
        if (c2 > 0.0) atomicMax(t + 1, (unsigned long long)__double_as_longlong(c2));
    }
}

// pair + local task kernel: blocks [0, nblk_pair) run the pair role with GP lanes per bead,
// blocks [nblk_pair, nblk_pair + nblk_loc) the local role with 8 lanes per bead
template <int GP>
__global__ void __launch_bounds__(256, 4) k_force_local(DevSys S, FFParams P, StepArgs A)
{
    const DevCtl ctl = S.ctl[0];
    const int rep = (A.only_rep >= 0) ? A.only_rep : blockIdx.y;
    bool blocked = false;
    if (A.integrate) {
        blocked = step_blocked(S, ctl.step, A.only_rep);
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
            DevCtl b; b.step = ctl.step; b.cur = ctl.cur; b.blocked = blocked;
            S.ctl[1] = b;
        }
        if (blocked) return;
    }
    const double4 *pos = S.pos[ctl.cur] + (size_t)rep * S.nmp;
    if ((int)blockIdx.x < A.nblk_pair) phase_pair<GP>(S, P, A, pos, rep, blockIdx.x);
    else phase_local(S, P, A, pos, rep, blockIdx.x - A.nblk_pair, ctl.step + (A.integrate ? 1ull : 0ull));
}

// bp task kernel: one thread per half-list candidate
__global__ void __launch_bounds__(128) k_force_bp(DevSys S, FFParams P, StepArgs A)
{
    const DevCtl ctl = S.ctl[0];
    if (A.integrate && step_blocked(S, ctl.step, A.only_rep)) return;
    const int rep = (A.only_rep >= 0) ? A.only_rep : blockIdx.y;
    phase_bp(S, P, A, S.pos[ctl.cur] + (size_t)rep * S.nmp, rep, blockIdx.x * blockDim.x + threadIdx.x);
}

// integrate kernel; integrate == 0: only sums the partial forces into S.force (per-irep force())
__global__ void __launch_bounds__(256) k_integrate(DevSys S, FFParams P, StepArgs A)
{
    const DevCtl ctl = A.integrate ? S.ctl[1] : S.ctl[0];
    if (A.integrate && ctl.blocked) return;
    const int rep = (A.only_rep >= 0) ? A.only_rep : blockIdx.y;
    const unsigned long long istep = ctl.step + 1ull;
    phase_integrate(S, P, A, ctl.cur, rep, blockIdx.x, istep);
    if (A.integrate && blockIdx.x == 0 && threadIdx.x == 0) {
        // clear the slot the NEXT step will accumulate into; publish step / buffer for the next force phase
        unsigned long long *t = S.top2 + ((size_t)rep * 2 + ((istep + 1ull) & 1ull)) * 2;
        t[0] = 0ull; t[1] = 0ull;
        if (blockIdx.y == 0) {
            DevCtl a; a.step = istep; a.cur = ctl.cur ^ 1; a.blocked = 0;
            S.ctl[0] = a;
        }
    }
}

// ---- base-pair evaluation pass (owner entries only) ------------------------------------------
// SIDE 0: force-side formulas (pass 1 of src/force_bp_limit_triplet.F90:85-257)
// SIDE 1: energy-side formulas (src/energy_bp*.F90), used when flg_bp_energy = .False.
// count_excess: nt_bp_excess accumulation for the MC pruning
template <int SIDE>
__global__ void __launch_bounds__(256) k_bp_eval(DevSys S, FFParams P, int only_rep, double tK_override,
                                                 int use_override, int use_mask, int count_excess, int in_md)
{
    if (in_md && step_blocked(S, S.ctl[0].step, only_rep)) return;
    const int rep = (only_rep >= 0) ? only_rep : blockIdx.y;
    const int nmp = S.nmp;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over nmp*KB slots
    if (idx >= nmp * S.KB) return;
    const int k = idx / S.KB, e = idx % S.KB;
    const size_t rbase = (size_t)rep * nmp, pbase = rbase * S.KB;
    if (e >= S.nbp_row[rbase + k]) return;
    const int2 ent = S.bprow[(rbase + k) * S.KB + e];
    const int j = ent.x, ibp = ent.y;
    if (j < k) return;   // owner = lower bead
    const double4 *pos = S.pos[S.ctl[0].cur] + rbase;
    const RepParams R = S.rp[rep];
    int st = 0;
    double en = 0.0;
    bool active = !(use_mask && !S.bp_status_MC[pbase + ibp]);
    const int type = S.bp_type[pbase + ibp];
    double dG = 1.0;
    if (P.bp_triplet) {
        double tK = use_override ? tK_override : R.tempK;
        if (SIDE == 1 && P.temp_independent) tK = 0.0;   // src/energy_bp_triplet.F90:27-31
        double2 cf = S.bp_coef[pbase + ibp];
        dG = cf.x - tK * cf.y;
        if (dG >= 0.0) active = false;
    }
    if (active) {
        const int4 mi = S.meta[k], mj = S.meta[j];
        Vec3 xs[6];
        xs[0] = v3(pos[k]); xs[1] = v3(pos[j]);
        xs[2] = v3(pos[chain_nb(mi, -1)]); xs[3] = v3(pos[chain_nb(mj, -1)]);
        xs[4] = v3(pos[chain_nb(mi, +1)]); xs[5] = v3(pos[chain_nb(mj, +1)]);
        double u, a12; Vec3 fr[6], v12;
        bool pass = (SIDE == 0) ? bp_force_all(P, P.bp[type], xs, &u, fr, &a12, &v12)
                                : bp_energy_u(P, P.bp[type], xs, &u);
        if (pass) {
            double ene = P.bp_triplet ? P.bp[type].U0 * dG * exp(-u) : P.bp[type].U0 * exp(-u);
            if (ene <= P.bp_cutoff_energy) { st = 1; en = ene; }
        }
    }
    S.bp_status[pbase + ibp] = st;
    S.ene_bp[pbase + ibp] = en;
    if (count_excess && st) {
        atomicAdd(&S.nt_excess[rbase + k], 1);
        atomicAdd(&S.nt_excess[rbase + j], 1);
    }
}

__global__ void k_excess_init(DevSys S, int only_rep, int max_bp_per_nt, int in_md)
{
    if (in_md && step_blocked(S, S.ctl[0].step, only_rep)) return;
    const int rep = (only_rep >= 0) ? only_rep : blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S.nmp) S.nt_excess[(size_t)rep * S.nmp + i] = -max_bp_per_nt;
}

// pass 2 of src/force_bp_limit_triplet.F90:260-328: one warp per replica runs the reference's
// sequential delete loop; uniforms come from the counter-based stream 0x40000000 + grep.
// deterministic != 0: temp_independent branch of src/energy_bp_limit_triplet.F90:167-176.
SIS_D void bp_prune_warp(const DevSys &S, const FFParams &P, unsigned long long seed, int rep, double tK_override,
                         int use_override, int deterministic, int copy_to_mc)
{
    const int lane = threadIdx.x & 31;
    const int nmp = S.nmp, KB = S.KB;
    const size_t rbase = (size_t)rep * nmp, pbase = rbase * KB;
    const RepParams R = S.rp[rep];
    const double beta = 1.0 / (SIS_BOLTZ_KCAL_MOL * (use_override ? tK_override : R.tempK));
    int *excess = S.nt_excess + rbase;
    unsigned long long ctr = S.mc_counter[rep];
    const uint32_t stream = 0x40000000u + (uint32_t)R.stream;
    for (;;) {
        int nnt = 0;
        for (int i = lane; i < nmp; i += 32) nnt += excess[i] > 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nnt += __shfl_xor_sync(0xffffffffu, nnt, o);
        if (nnt == 0) break;
        double rnd = uniform_counter(seed, stream, ctr++);
        int pick = max(1, min((int)ceil(rnd * nnt), nnt));
        // the pick-th bead with excess > 0 in ascending bead order (uniform across the warp)
        int nt_delete = -1, seen = 0;
        for (int base = 0; base < nmp && nt_delete < 0; base += 32) {
            const int i = base + lane;
            const unsigned m = __ballot_sync(0xffffffffu, i < nmp && excess[i] > 0);
            const int c = __popc(m);
            if (seen + c >= pick) {
                int want = pick - seen;
                for (int b = 0; b < 32; b++)
                    if ((m >> b) & 1u)
                        if (--want == 0) { nt_delete = base + b; break; }
            }
            seen += c;
        }
        unsigned long long used = 0;
        if (lane == 0) {
            const int nrow = S.nbp_row[rbase + nt_delete];
            const int2 *row = S.bprow + (rbase + nt_delete) * KB;
            int seq[64], nseq = 0;
            for (int e = 0; e < nrow && nseq < 64; e++)
                if (S.bp_status[pbase + row[e].y]) seq[nseq++] = row[e].y;
            int ibp_delete = seq[0];
            if (!deterministic) {
                for (int i = 0; i < nseq; i++) {   // the reference's shuffle, :295-302
                    rnd = uniform_counter(seed, stream, ctr + used++);
                    const int sw = max(1, min((int)ceil(rnd * nseq), nseq));
                    const int t = seq[i]; seq[i] = seq[sw - 1]; seq[sw - 1] = t;
                }
                ibp_delete = seq[0];
                for (int i = 1; i < nseq; i++) {   // Metropolis sweep, :305-316
                    const double ratio = exp((S.ene_bp[pbase + seq[i]] - S.ene_bp[pbase + ibp_delete]) * beta);
                    rnd = uniform_counter(seed, stream, ctr + used++);
                    if (rnd < ratio) ibp_delete = seq[i];
                }
            } else {
                for (int i = 1; i < nseq; i++)
                    if (S.ene_bp[pbase + seq[i]] > S.ene_bp[pbase + ibp_delete]) ibp_delete = seq[i];
            }
            S.bp_status[pbase + ibp_delete] = 0;
            const int owner = ibp_delete / KB;
            const int partner = S.bprow[(rbase + owner) * KB + ibp_delete % KB].x;
            excess[owner] -= 1;
            excess[partner] -= 1;
        }
        ctr += __shfl_sync(0xffffffffu, used, 0);
        __syncwarp();
    }
    if (lane == 0) S.mc_counter[rep] = ctr;
    __syncwarp();
    if (copy_to_mc)   // src/job_md.F90:399-401
        for (int i = lane; i < nmp * KB; i += 32) S.bp_status_MC[pbase + i] = S.bp_status[pbase + i];
}
__global__ void __launch_bounds__(32) k_bp_prune(DevSys S, FFParams P, unsigned long long seed, int only_rep,
                                                  double tK_override, int use_override, int deterministic,
                                                  int copy_to_mc, int in_md)
{
    if (in_md && step_blocked(S, S.ctl[0].step, only_rep)) return;
    bp_prune_warp(S, P, seed, (only_rep >= 0) ? only_rep : blockIdx.x, tK_override, use_override, deterministic, copy_to_mc);
}

__global__ void k_copy_status_mc(DevSys S, int only_rep)
{
    const int rep = (only_rep >= 0) ? only_rep : blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t pbase = (size_t)rep * S.nmp * S.KB;
    if (i < S.nmp * S.KB) S.bp_status_MC[pbase + i] = S.bp_status[pbase + i];
}

// ---- Rg pre-pass (src/force_rg.F90:22-30, src/energy_rg.F90:22-28): one block per replica ----
__global__ void __launch_bounds__(256) k_rg(DevSys S, int only_rep)
{
    const int rep = (only_rep >= 0) ? only_rep : blockIdx.x;
    const int nmp = S.nmp;
    const double4 *pos = S.pos[S.ctl[0].cur] + (size_t)rep * nmp;
    __shared__ double sh[4][8];
    __shared__ double cm[3];
    double sx = 0, sy = 0, sz = 0;
    for (int i = threadIdx.x; i < nmp; i += blockDim.x) { double4 p = pos[i]; sx += p.x; sy += p.y; sz += p.z; }
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (threadIdx.x % 32 == 0) { sh[0][threadIdx.x / 32] = sx; sh[1][threadIdx.x / 32] = sy; sh[2][threadIdx.x / 32] = sz; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0;
        for (int w = 0; w < 8; w++) { a += sh[0][w]; b += sh[1][w]; c += sh[2][w]; }
        cm[0] = a / nmp; cm[1] = b / nmp; cm[2] = c / nmp;
    }
    __syncthreads();
    double s2 = 0;
    for (int i = threadIdx.x; i < nmp; i += blockDim.x) {
        double4 p = pos[i];
        double dx = p.x - cm[0], dy = p.y - cm[1], dz = p.z - cm[2];
        s2 += dx * dx + dy * dy + dz * dz;
    }
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (threadIdx.x % 32 == 0) sh[3][threadIdx.x / 32] = s2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; w++) t += sh[3][w];
        RepParams *R = S.rp + rep;
        R->rg_cm[0] = cm[0]; R->rg_cm[1] = cm[1]; R->rg_cm[2] = cm[2];
        R->rg_val = sqrt(t / nmp);
    }
}

// ---- energies (src/energy_sumup.F90:15-54) ------------------------------------------------
// pair and bonded terms are counted once (by the lower bead / the first bead of the term);
// bp energies are the cached ene_bp written by k_step / k_bp_eval, masked by bp_status.
// rp_alt != nullptr evaluates with another parameter block (REMD cross energies).
template <int G>
__global__ void __launch_bounds__(256) k_energy(DevSys S, FFParams P, const RepParams *rp_alt, int only_rep)
{
    constexpr int BPB = 256 / G;
    const int rep = (only_rep >= 0) ? only_rep : blockIdx.y;
    const int lane = threadIdx.x % G;
    const int k = blockIdx.x * BPB + threadIdx.x / G;
    const int nmp = S.nmp;
    const double4 *pos = S.pos[S.ctl[0].cur] + (size_t)rep * nmp;
    const RepParams R = rp_alt ? rp_alt[rep] : S.rp[rep];
    double e[ENE_N];
#pragma unroll
    for (int t = 0; t < ENE_N; t++) e[t] = 0.0;
    if (k < nmp) {
        const int4 mk4 = S.meta[k];
        const Vec3 xk = v3(pos[k]);
        const uint32_t *row = S.nbr + ((size_t)rep * nmp + k) * S.K;
        const int nn = S.nnbr[(size_t)rep * nmp + k];
        for (int q = lane; q < nn; q += G) {
            uint32_t ent = row[q];
            int j = ent & NB_IDX;
            if (j < k) continue;
            Vec3 v = pbc_vec_d(P, xk, v3(pos[j]));
            double d2 = dot(v, v);
            if (ent & NB_WCA) e[ENE_EXV] += wca_energy(P, d2);
            if ((ent & NB_ELE) && P.flg_ele) e[ENE_ELE] += dh_energy(P, R, d2);
        }
        if (lane < 3) {
            const int span = lane + 1;
            bool exists = (mk4.w & MF_CIRC) ? (mk4.y > span) : (mk4.z + span < mk4.y);
            if (exists) {
                Vec3 x1 = v3(pos[chain_nb(mk4, 1)]);
                if (lane == 0) e[ENE_BOND] += bond_energy(P, xk, x1);
                else if (lane == 1) e[ENE_ANGL] += angl_energy(P, xk, x1, v3(pos[chain_nb(mk4, 2)]));
                else if (P.flg_dih_cos || P.flg_dih_exp)
                    e[ENE_DIHE] += dih_energy(P, xk, x1, v3(pos[chain_nb(mk4, 2)]), v3(pos[chain_nb(mk4, 3)]));
            }
        }
        {
            const size_t rbase = (size_t)rep * nmp, pbase = rbase * S.KB;
            const int nrow = S.nbp_row[rbase + k];
            for (int q = lane; q < nrow; q += G) {
                int2 ent = S.bprow[(rbase + k) * S.KB + q];
                if (ent.x > k && S.bp_status[pbase + ent.y]) e[ENE_BP] += S.ene_bp[pbase + ent.y];
            }
        }
        if (lane == 0 && P.flg_stage) {   // src/energy_stage.F90:28-36
            double z = P.stage_sigma / xk.z;
            e[ENE_STAGE] += 4 * P.stage_eps * (powi(z, 12) - powi(z, 6));
            if (xk.z < 0.85 * P.stage_sigma) S.misc->max_force_flag |= 2;   // reference aborts here
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int t = 0; t < P.ntwz_DCF; t++) {   // src/energy_tweezers.F90:21-31
            Vec3 v = pbc_vec_d(P, v3(pos[S.twz_pairs[2 * t]]), v3(pos[S.twz_pairs[2 * t + 1]]));
            e[ENE_TWZ] -= R.twz_force * (v.x * S.twz_dir[3 * t] + v.y * S.twz_dir[3 * t + 1] + v.z * S.twz_dir[3 * t + 2]);
        }
        for (int t = 0; t < P.nrest; t++) {      // src/energy_restraint.F90:19-33
            Vec3 v = pbc_vec_d(P, v3(pos[S.rest_id[3 * t]]), v3(pos[S.rest_id[3 * t + 1]]));
            double r = sqrt(dot(v, v));
            if (r >= S.rest_rcut[t]) continue;
            e[ENE_REST] -= 0.5 * S.rest_para[3 * t] * (1.0 - tanh((r - S.rest_para[3 * t + 1]) / S.rest_para[3 * t + 2]));
        }
        if (P.flg_bias_rg) {                     // src/energy_rg.F90:30-35 (rg from k_rg)
            double rg = R.rg_val;
            if (!(P.bias_rg_pott == 2 && rg <= R.bias_rg_0))
                e[ENE_RG] += 0.5 * P.bias_rg_k * (rg - R.bias_rg_0) * (rg - R.bias_rg_0);
        }
    }
    // block reduction (fixed order -> deterministic sums)
    __shared__ double sh[8][ENE_N];
    __shared__ bool s_last;
#pragma unroll
    for (int t = 1; t < ENE_N; t++) {
        double v = e[t];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x % 32 == 0) sh[threadIdx.x / 32][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < ENE_N && threadIdx.x > 0) {
        double v = 0;
        for (int w = 0; w < 8; w++) v += sh[w][threadIdx.x];
        S.ene_partial[((size_t)rep * S.nblk_ene + blockIdx.x) * ENE_N + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int done = atomicAdd(&S.ene_counter[rep], 1u);
        s_last = (done == (unsigned)S.nblk_ene - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (threadIdx.x < ENE_N && threadIdx.x > 0) {
            const volatile double *part = S.ene_partial + (size_t)rep * S.nblk_ene * ENE_N;
            double v = 0;
            for (int b = 0; b < S.nblk_ene; b++) v += part[(size_t)b * ENE_N + threadIdx.x];
            S.ene_out[(size_t)rep * ENE_N + threadIdx.x] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0;
            for (int t = 1; t < ENE_N; t++) tot += ((volatile double *)S.ene_out)[(size_t)rep * ENE_N + t];
            S.ene_out[(size_t)rep * ENE_N] = tot;
            S.ene_counter[rep] = 0;
        }
    }
}

// src/energy_kinetic.F90:17-29 (one block per replica; zeroes frozen velocities first)
__global__ void __launch_bounds__(256) k_kinetic(DevSys S, double *out, int only_rep)
{
    const int rep = (only_rep >= 0) ? only_rep : blockIdx.x;
    const int nmp = S.nmp;
    double *vel = S.vel + (size_t)rep * nmp * 3;
    double s = 0;
    for (int i = threadIdx.x; i < nmp; i += blockDim.x) {
        if (S.meta[i].w & MF_FROZEN) vel[3 * i] = vel[3 * i + 1] = vel[3 * i + 2] = 0.0;
        s += S.mass[i] * (vel[3 * i] * vel[3 * i] + vel[3 * i + 1] * vel[3 * i + 1] + vel[3 * i + 2] * vel[3 * i + 2]);
    }
    __shared__ double sh[8];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x % 32 == 0) sh[threadIdx.x / 32] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; w++) t += sh[w];
        out[rep] = 0.5 * t;
    }
}

// src/job_md.F90:157-176 initial velocities and beta(0) from the counter-based generator
__global__ void k_init_velocities(DevSys S, unsigned long long seed, int rep)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.nmp * 3) return;
    const int k = i / 3, c = i % 3;
    const RepParams R = S.rp[rep];
    const size_t o = ((size_t)rep * S.nmp + k) * 3 + c;
    const double c1 = sqrt(R.tempK * SIS_BOLTZ_KCAL_MOL / S.mass[k]);
    S.vel[o] = c1 * gauss_component(seed, (uint32_t)R.stream, 0x8000000000000000ull, (uint32_t)k, c);
    S.acc[o] = (S.coef[k].w * R.sqrt_tempK) * gauss_component(seed, (uint32_t)R.stream, 0x8000000000000001ull, (uint32_t)k, c);
}

// displacement top-2 of every replica recomputed from xyz_move into the slot of `step` (the other slot
// cleared): after sisgpu_set_step moved the step counter the pending rebuild trigger of
// src/job_md.F90:376-388 must stay where the new parity looks for it.  One block per replica.
__global__ void __launch_bounds__(256) k_top2_from_move(DevSys S, unsigned long long step)
{
    const int rep = blockIdx.x, nmp = S.nmp;
    const double *mv = S.move + (size_t)rep * nmp * 3;
    double t1 = 0.0, t2 = 0.0;
    for (int i = threadIdx.x; i < nmp; i += blockDim.x)
        top2_merge(t1, t2, mv[3 * i] * mv[3 * i] + mv[3 * i + 1] * mv[3 * i + 1] + mv[3 * i + 2] * mv[3 * i + 2], 0.0);
    __shared__ double sh[8][2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double b1 = __shfl_xor_sync(0xffffffffu, t1, o), b2 = __shfl_xor_sync(0xffffffffu, t2, o);
        top2_merge(t1, t2, b1, b2);
    }
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][0] = t1; sh[threadIdx.x >> 5][1] = t2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a1 = 0.0, a2 = 0.0;
        for (int w = 0; w < 8; w++) top2_merge(a1, a2, sh[w][0], sh[w][1]);
        unsigned long long *t = S.top2 + ((size_t)rep * 2 + (step & 1ull)) * 2;
        t[0] = (unsigned long long)__double_as_longlong(a1); t[1] = (unsigned long long)__double_as_longlong(a2);
        t = S.top2 + ((size_t)rep * 2 + ((step + 1ull) & 1ull)) * 2;
        t[0] = 0ull; t[1] = 0ull;
    }
}

// T-REMD velocity rescale (src/replica_exchange.F90:175-213)
__global__ void k_scale_velocities(DevSys S, int rep, double factor)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S.nmp * 3) S.vel[(size_t)rep * S.nmp * 3 + i] *= factor;
}

// ---- neighbour list on a cell grid (replaces the O(N^2) sweep of src/neighbor_list.F90) ------
// Every kernel of the chain reads the grid (S.grid) and the replica flags (rep_active) from device
// memory, so the same launches serve the host-driven build (build_lists: flags and grid written by
// the host) and the device-driven rebuild inside the MD loop (k_nl_plan writes both; a chain whose
// replicas are all within the list margin falls through in a few microseconds without a host
// round trip).
struct NlPlanArgs {
    double rc;              // cell edge lower bound: largest list cutoff + margin
    double ele_cutoff_inp;  // cutoff_type 2: ele_cutoff = ele_cutoff_inp * lambdaD, refreshed at every rebuild
    int ele_cutoff_type, flg_ele, pbc;
    int force_all;          // rebuild every replica whatever its displacement
    long long stop;         // no rebuild once this step is reached: the lists of step stop + 1 belong to the next call
};
// 0. which replicas rebuild (the test of src/job_md.F90:388 on the displacement top-2 of the last
//    completed step), the ele_cutoff refresh of src/neighbor_list.F90:80-82, and -- without PBC --
//    the bounding box of those replicas and the cell grid over it.  One block.
__global__ void __launch_bounds__(1024) k_nl_plan(DevSys S, NlPlanArgs A, int *rep_active)
{
    __shared__ int s_any;
    __shared__ double s_lo[3][32], s_hi[3][32];
    const int tid = threadIdx.x, R = S.nrep, nmp = S.nmp;
    const DevCtl ctl = S.ctl[0];
    const long long stop = A.stop < 0 ? S.misc->stop_step : A.stop;   // < 0: the device-side stop step (k_md_big cycles)
    // lists already marked invalid (a row of an earlier chain of this batch did not fit): nothing is
    // rebuilt and the overflow marks are kept for the host, which grows the capacities
    const bool invalid = S.misc->lists_invalid != 0;
    if (tid == 0) { s_any = 0; if (!invalid) { S.misc->nl_overflow = 0; S.misc->bp_overflow = 0; } }
    __syncthreads();
    for (int r = tid; r < R; r += blockDim.x) {
        const unsigned long long *t = S.top2 + ((size_t)r * 2 + (ctl.step & 1ull)) * 2;
        const double m1 = __longlong_as_double((long long)t[0]), m2 = __longlong_as_double((long long)t[1]);
        const int need = !invalid && (A.force_all || ((long long)ctl.step < stop && sqrt(m1) + sqrt(m2) > S.nl_margin));
        rep_active[r] = need;
        if (need) {
            s_any = 1;
            S.nbp_half[r] = 0;
            atomicAdd(&S.misc->nrebuild, 1);
            if (A.flg_ele && A.ele_cutoff_type == 2) S.rp[r].ele_cutoff = A.ele_cutoff_inp * S.rp[r].lambdaD;
        }
    }
    __syncthreads();
    if (tid == 0 && s_any && S.misc->inject_invalid_at > 0 && S.misc->nrebuild >= S.misc->inject_invalid_at) {
        S.misc->lists_invalid = 1;      // exercised by tests: the host must recover exactly as after a row overflow
        S.misc->inject_invalid_at = 0;
    }
    if (!s_any || A.pbc) return;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    const double4 *pos = S.pos[ctl.cur];
    for (int it = tid; it < R * nmp; it += blockDim.x) {
        if (!rep_active[it / nmp]) continue;
        const double4 p = pos[it];
        lo[0] = fmin(lo[0], p.x); hi[0] = fmax(hi[0], p.x);
        lo[1] = fmin(lo[1], p.y); hi[1] = fmax(hi[1], p.y);
        lo[2] = fmin(lo[2], p.z); hi[2] = fmax(hi[2], p.z);
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        if ((tid & 31) == 0) { s_lo[d][tid >> 5] = lo[d]; s_hi[d][tid >> 5] = hi[d]; }
    }
    __syncthreads();
    if (tid == 0) {
        GridInfo G;
        G.pbc = 0;
        long long ncell = 1;
        for (int d = 0; d < 3; d++) {
            double l = 1e300, h = -1e300;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) { l = fmin(l, s_lo[d][w]); h = fmax(h, s_hi[d][w]); }
            const double ext = fmax(h - l, 1e-9);
            int n = (int)floor(ext / (A.rc * 1.000001));
            n = max(1, min(n, 128));
            G.n[d] = n; G.origin[d] = l; G.inv_cell[d] = n / ext;
            ncell *= n;
        }
        while (ncell > S.ncell_cap) {       // stay inside the allocated grid: coarser cells are still exact
            int d = G.n[0] >= G.n[1] ? (G.n[0] >= G.n[2] ? 0 : 2) : (G.n[1] >= G.n[2] ? 1 : 2);
            G.n[d]--;
            ncell = (long long)G.n[0] * G.n[1] * G.n[2];
        }
        for (int d = 0; d < 3; d++) {
            double l = 1e300, h = -1e300;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) { l = fmin(l, s_lo[d][w]); h = fmax(h, s_hi[d][w]); }
            G.inv_cell[d] = G.n[d] / fmax(h - l, 1e-9);
        }
        G.ncell = (int)ncell;
        *S.grid = G;
    }
}
// 1. wrap (src/pbc.F90:196-226) + cell id + per-cell count
__global__ void k_nl_wrap_count(DevSys S, FFParams P, const int *rep_active)
{
    const int rep = blockIdx.y;
    if (!rep_active[rep]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.nmp) return;
    const GridInfo &Gd = *S.grid;
    double4 *pos = S.pos[S.ctl[0].cur] + (size_t)rep * S.nmp;
    double4 p = pos[i];
    if (P.flg_pbc) {
        p.x = wrap1(p.x, P.box[0]); p.y = wrap1(p.y, P.box[1]); p.z = wrap1(p.z, P.box[2]);
        pos[i] = p;
    }
    int cx = min(max((int)((p.x - Gd.origin[0]) * Gd.inv_cell[0]), 0), Gd.n[0] - 1);
    int cy = min(max((int)((p.y - Gd.origin[1]) * Gd.inv_cell[1]), 0), Gd.n[1] - 1);
    int cz = min(max((int)((p.z - Gd.origin[2]) * Gd.inv_cell[2]), 0), Gd.n[2] - 1);
    int c = (cz * Gd.n[1] + cy) * Gd.n[0] + cx;
    S.cell_of[(size_t)rep * S.nmp + i] = c;
    atomicAdd(&S.cell_count[(size_t)rep * S.ncell_cap + c], 1);
}
// 2. exclusive scan of the cell counts (one block per replica); leaves the counts cleared for
//    the next build
__global__ void __launch_bounds__(1024) k_nl_scan(DevSys S, const int *rep_active)
{
    const int rep = blockIdx.x;
    if (!rep_active[rep]) return;
    const int ncell = S.grid->ncell;
    int *cnt = S.cell_count + (size_t)rep * S.ncell_cap;
    int *start = S.cell_start + (size_t)rep * (S.ncell_cap + 1);
    int *cur = S.cell_cursor + (size_t)rep * S.ncell_cap;
    __shared__ int wsum[32];
    __shared__ int carry, chunk_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ncell; base += 1024) {
        const int i = base + tid;
        const int v = i < ncell ? cnt[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wsum[lane] = wi - w;            // exclusive prefix of the warp totals
            if (lane == 31) chunk_total = wi;
        }
        __syncthreads();
        if (i < ncell) { start[i] = carry + wsum[warp] + incl - v; cur[i] = 0; cnt[i] = 0; }
        __syncthreads();
        if (tid == 0) carry += chunk_total;
        __syncthreads();
    }
    if (tid == 0) start[ncell] = carry;
}
// 3. scatter beads into their cells (arrival order)
__global__ void k_nl_fill(DevSys S, const int *rep_active)
{
    const int rep = blockIdx.y;
    if (!rep_active[rep]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.nmp) return;
    int c = S.cell_of[(size_t)rep * S.nmp + i];
    int slot = S.cell_start[(size_t)rep * (S.ncell_cap + 1) + c] + atomicAdd(&S.cell_cursor[(size_t)rep * S.ncell_cap + c], 1);
    S.cell_beads[(size_t)rep * S.nmp + slot] = i;
}
// 4. every cell in ascending bead order (deterministic rows): bead i goes to the slot given by
//    the number of smaller bead ids in its cell
__global__ void k_nl_rank(DevSys S, const int *rep_active)
{
    const int rep = blockIdx.y;
    if (!rep_active[rep]) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.nmp) return;
    const size_t rbase = (size_t)rep * S.nmp;
    const int c = S.cell_of[rbase + i];
    const int *start = S.cell_start + (size_t)rep * (S.ncell_cap + 1);
    const int s = start[c], e = start[c + 1];
    const int *b = S.cell_beads + rbase;
    int rank = 0;
    for (int q = s; q < e; q++) rank += b[q] < i;
    S.cell_sorted[rbase + s + rank] = i;
}

SIS_D int uniq_cells(int c, int n, int pbc, int out[3])
{
    int cnt = 0;
    for (int d = -1; d <= 1; d++) {
        int q = c + d;
        if (pbc) q = (q + n) % n;
        else if (q < 0 || q >= n) continue;
        bool dup = false;
        for (int t = 0; t < cnt; t++) dup |= (out[t] == q);
        if (!dup) out[cnt++] = q;
    }
    return cnt;
}

// 5. rows: the three lists of src/neighbor_list.F90:99-229 for bead i against every bead of the
//    27 surrounding cells, same predicates (d2 <= cut^2 on wrapped coordinates, min image),
//    same exclusion rules; FULL rows (j on both sides of i).  One warp per bead, ballot-compacted
//    appends keep the row order deterministic (cell order, then bead id).
struct NlArgs {
    double wca_nl_cut2, bp_nl_cut2, nl_margin;
    int excl_bond, excl_angle, bp_model, bp_min_loop;
    int all_pairs;   // 0 cell grid; 1 list_exv/ele/bp (no cutoffs); 2 all-pairs sweep with cutoffs
    int ele_cutoff_type;     // in-kernel rebuilds (k_md_fast, replica-local mode): the ele_cutoff refresh
    double ele_cutoff_inp;   // of src/neighbor_list.F90:80-82
};
// the list predicates of src/neighbor_list.F90:122-180 (and of src/list_exv.F90:36-52,
// src/list_ele.F90:36-56 when A.all_pairs == 1) for the unordered bead pair {i, j}
SIS_HD uint32_t nl_pair_flags(const NlArgs &A, const FFParams &P, int4 mi, int4 mj, int i, int j, double d2,
                              double ele_nl_cut2)
{
    const bool full = A.all_pairs == 1;
    uint32_t ent = 0;
    const bool same = mi.x == mj.x;
    // (lo, hi) = the reference's local (i, j) with i < j inside a chain
    const int lo = min(mi.z, mj.z) + 1, hi = max(mi.z, mj.z) + 1, len = mi.y;
    const bool circ = mi.w & MF_CIRC;
    if ((mi.w & MF_SEQ_MASK) != SEQ_D && (mj.w & MF_SEQ_MASK) != SEQ_D && (full || d2 <= A.wca_nl_cut2)) {
        bool add = true;
        if (same) {
            if (lo + 2 >= hi) add = false;
            else if (circ && !full && hi + 2 >= lo + len) add = false;
        }
        if (add) ent |= NB_WCA;
    }
    if (P.flg_ele && (full || d2 <= ele_nl_cut2) && (mi.w & MF_CHARGE) && (mj.w & MF_CHARGE)) {
        bool add = true;
        if (full) {
            // list_ele tests j == i+1 / i+2 on LOCAL indices for every chain pair
            const int li = (i < j ? mi.z : mj.z) + 1, lj = (i < j ? mj.z : mi.z) + 1;
            if (lj == li + 1 && A.excl_bond) add = false;
            if (lj == li + 2 && A.excl_angle) add = false;
        } else if (same) {
            if (hi == lo + 1 && A.excl_bond) add = false;
            if (hi == lo + 2 && A.excl_angle) add = false;
            if (circ) {
                if (lo == 1) {
                    if (hi == len && A.excl_bond) add = false;
                    if (hi == len - 1 && A.excl_angle) add = false;
                } else if (lo == 2) {
                    if (hi == len && A.excl_angle) add = false;
                }
            }
        }
        if (add) ent |= NB_ELE;
    }
    return ent;
}
// A listed pair can need the periodic image shift of pbc_vec_d before the next rebuild only if a
// raw coordinate difference can reach L/2: both beads together move at most nl_margin between
// rebuilds (the trigger of src/job_md.F90:388), so |raw| + margin < L/2 on every axis proves
// that v = x_i - x_j is already the minimum image for the whole life of the list.
SIS_HD bool nl_needs_image(const FFParams &P, const NlArgs &A, Vec3 xi, Vec3 xj)
{
    if (!P.flg_pbc) return false;
    if (A.all_pairs == 1) return true;   // cutoff-free lists are never rebuilt
    const double m = A.nl_margin * 1.0001 + 1.0e-9;
    return !(fabs(xi.x - xj.x) + m < P.box_half[0] && fabs(xi.y - xj.y) + m < P.box_half[1] &&
             fabs(xi.z - xj.z) + m < P.box_half[2]);
}
// distance / index gate of the bp list: src/neighbor_list.F90:183, src/list_bp.F90:57
SIS_HD bool nl_bp_geometry(const NlArgs &A, int i, int j, double d2)
{
    return A.all_pairs == 1 ? (abs(i - j) >= A.bp_min_loop + 1) : (d2 <= A.bp_nl_cut2);
}
SIS_D void nl_rows_bead_at(const DevSys &S, const FFParams &P, const NlArgs &A, int rep, int i, int lane,
                           const double4 *pos, int *s_off, int *s_beg)
{
    const int nmp = S.nmp;
    const size_t rbase = (size_t)rep * nmp;
    const RepParams R = S.rp[rep];
    const int4 mi = S.meta[i];
    const Vec3 xi = v3(pos[i]);
    const double ele_cut = R.ele_cutoff + A.nl_margin;
    const double ele_nl_cut2 = ele_cut * ele_cut;
    const double max_nl_cut2 = fmax(fmax(A.wca_nl_cut2, A.bp_nl_cut2), P.flg_ele ? ele_nl_cut2 : 0.0);
    uint32_t *row = S.nbr + (rbase + i) * S.K;
    int2 *brow = S.bprow + (rbase + i) * S.KB;
    int nn = 0, nb = 0;
    const int *cstart = S.cell_start + (size_t)rep * (S.ncell_cap + 1);
    const int *cbeads = S.cell_sorted + rbase;

    const bool sweep_all = A.all_pairs != 0;   // iterate every bead instead of the 27 cells
    // candidate stream: the beads of the (up to 27) surrounding cells, concatenated in the fixed
    // order (z, y, x neighbour; ascending bead id inside a cell).  Lane r < nrange holds range r;
    // a warp scan gives the stream offsets, so every iteration tests 32 candidates whatever the
    // cell occupancies.
    int total;
    if (sweep_all) {
        total = nmp;
    } else {
        const GridInfo &Gd = *S.grid;
        int lx[3], ly[3], lz[3];
        const int c = S.cell_of[rbase + i];
        const int cx = c % Gd.n[0], cy = (c / Gd.n[0]) % Gd.n[1], cz = c / (Gd.n[0] * Gd.n[1]);
        const int nx = uniq_cells(cx, Gd.n[0], Gd.pbc, lx);
        const int ny = uniq_cells(cy, Gd.n[1], Gd.pbc, ly);
        const int nz = uniq_cells(cz, Gd.n[2], Gd.pbc, lz);
        const int nrange = nx * ny * nz;
        int beg = 0, len = 0;
        if (lane < nrange) {
            const int ix = lane % nx, iy = (lane / nx) % ny, iz = lane / (nx * ny);
            const int cell = (lz[iz] * Gd.n[1] + ly[iy]) * Gd.n[0] + lx[ix];
            beg = cstart[cell];
            len = cstart[cell + 1] - beg;
        }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane < 28) s_off[lane] = incl - len;   // lanes >= nrange: the total (sentinels)
        if (lane < 27) s_beg[lane] = beg;
        total = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
    }
    for (int t0 = 0; t0 < total; t0 += 32) {
        const int t = t0 + lane;
        uint32_t ent = 0;
        int bp3 = 0, bptype = 0, j = -1;
        if (t < total) {
            if (sweep_all) j = t;
            else {
                // range r with off[r] <= t < off[r + 1] (empty ranges have equal offsets)
                int lo = 0, hi = 27;
#pragma unroll
                for (int s = 0; s < 5; s++) {
                    const int mid = (lo + hi) >> 1;
                    if (s_off[mid] <= t) lo = mid; else hi = mid;
                }
                j = cbeads[s_beg[lo] + (t - s_off[lo])];
            }
            const Vec3 xj = j >= 0 ? v3(pos[j]) : xi;
            const Vec3 v = pbc_vec_d(P, xi, xj);
            // most candidates of the 27 cells lie outside every list radius (a sphere fills 15 % of the cube):
            // they leave here, before their meta data is touched.  Same predicates for those that stay.
            if (j != i && j >= 0 && (A.all_pairs == 1 || dot(v, v) <= max_nl_cut2)) {
                const int4 mj = S.meta[j];
                ent = nl_pair_flags(A, P, mi, mj, i, j, dot(v, v), ele_nl_cut2);
                if (ent && nl_needs_image(P, A, xi, xj)) ent |= NB_IMG;
                if (nl_bp_geometry(A, i, j, dot(v, v))) {
                    if (i < j) bp3 = bp_lookup_r(S.rs, A.bp_model, A.bp_min_loop, S.hash, S.seq, i, mi, j, mj, &bptype);
                    else bp3 = bp_lookup_r(S.rs, A.bp_model, A.bp_min_loop, S.hash, S.seq, j, mj, i, mi, &bptype);
                }
            }
        }
        // compacted appends
        unsigned m = __ballot_sync(0xffffffffu, ent != 0);
        if (ent) {
            int p = nn + __popc(m & ((1u << lane) - 1));
            if (p < S.K) row[p] = ent | (uint32_t)j;
        }
        nn += __popc(m);
        unsigned mb = __ballot_sync(0xffffffffu, bp3 > 0);
        if (bp3 > 0) {
            int p = nb + __popc(mb & ((1u << lane) - 1));
            if (p < S.KB) {
                brow[p] = make_int2(j, -1);
                if (i < j) {   // owner stores the pair data at ibp = i*KB + p
                    const size_t ibp = rbase * S.KB + (size_t)i * S.KB + p;
                    S.bp_type[ibp] = (signed char)bptype;
                    S.bp_coef[ibp] = (A.bp_model == 5) ? make_double2(S.bp3_dH[bp3], S.bp3_dS[bp3])
                                                       : make_double2(0.0, 0.0);
                    S.bp_status[ibp] = 0;
                    S.bp_status_MC[ibp] = 0;
                    S.ene_bp[ibp] = 0.0;
                }
            }
        }
        nb += __popc(mb);
    }
    if (lane == 0) {
        S.nnbr[rbase + i] = min(nn, S.K);
        S.nbp_row[rbase + i] = min(nb, S.KB);
        if (nn > S.K) atomicMax(&S.misc->nl_overflow, nn);
        if (nb > S.KB) atomicMax(&S.misc->bp_overflow, nb);
        S.move[(rbase + i) * 3] = 0.0; S.move[(rbase + i) * 3 + 1] = 0.0; S.move[(rbase + i) * 3 + 2] = 0.0;
    }
}
__global__ void __launch_bounds__(256) k_nl_rows(DevSys S, FFParams P, NlArgs A, const int *rep_active)
{
    const int rep = blockIdx.y;
    if (!rep_active[rep]) return;
    __shared__ int s_off[8][28], s_beg[8][27];
    const int lane = threadIdx.x % 32, w = threadIdx.x / 32;
    const int i = blockIdx.x * 8 + w;
    if (i >= S.nmp) return;
    nl_rows_bead_at(S, P, A, rep, i, lane, S.pos[S.ctl[0].cur] + (size_t)rep * S.nmp, s_off[w], s_beg[w]);
}
// 6. pair ids: ibp = owner*KB + slot-in-owner-row, resolved for both directions
SIS_D void nl_bp_slot(const DevSys &S, int rep, int idx)
{
    const int nmp = S.nmp, KB = S.KB;
    const int k = idx / KB, e = idx % KB;
    const size_t rbase = (size_t)rep * nmp;
    if (e >= S.nbp_row[rbase + k]) return;
    int2 *ent = S.bprow + (rbase + k) * KB + e;
    const int j = ent->x;
    if (j > k) {
        ent->y = k * KB + e;
        const int t = atomicAdd(&S.nbp_half[rep], 1);
        if (t < S.bp_half_cap) S.bp_half[(size_t)rep * S.bp_half_cap + t] = k * KB + e;
        return;
    }
    const int2 *jr = S.bprow + (rbase + j) * KB;
    const int nj = S.nbp_row[rbase + j];
    int found = -1;
    for (int q = 0; q < nj; q++)
        if (jr[q].x == k) { found = q; break; }
    ent->y = j * KB + max(found, 0);
}
__global__ void k_nl_bp_slots(DevSys S, const int *rep_active)
{
    const int rep = blockIdx.y;
    if (!rep_active[rep]) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < S.nmp * S.KB) nl_bp_slot(S, rep, idx);
}
__global__ void k_nl_begin_slots(DevSys S, const int *rep_active)
{
    const int rep = blockIdx.x * blockDim.x + threadIdx.x;
    if (rep < S.nrep && rep_active[rep]) S.nbp_half[rep] = 0;
}
// 7. end of a build.  (a) a rebuilt replica restarts its displacement bookkeeping (xyz_move = 0,
//    src/job_md.F90:390); (b) capacity check: a row or the bp half list that did not fit marks the
//    lists invalid so that no persistent kernel runs on them (the host grows and rebuilds);
//    (c) the CTA bead ranges of k_md_fast: `ncta` contiguous ranges of at most `per_cta` flattened
//    (replica, bead) items, cut where the running pair work (2 per row octet + 1 per bead) crosses
//    c * total / ncta.  One block of 1024 threads, nitems ints of dynamic shared memory;
//    nitems <= ncta * per_cta <= PART_MAX_ITEMS when ncta > 0.
constexpr int PART_MAX_ITEMS = 148 * 128;
__global__ void __launch_bounds__(1024) k_nl_partition(DevSys S, const int *rep_active, int ncta, int per_cta, int local)
{
    extern __shared__ int s_w[];                 // [nitems] inclusive prefix of the weights
    __shared__ int s_part[1024];
    __shared__ int s_cut[1025];
    __shared__ int s_any;
    const int tid = threadIdx.x, R = S.nrep, nitems = R * S.nmp;
    if (tid == 0) s_any = 0;
    __syncthreads();
    for (int r = tid; r < R; r += blockDim.x) {
        if (!rep_active[r]) continue;
        s_any = 1;
        for (int q = 0; q < 4; q++) S.top2[(size_t)r * 4 + q] = 0ull;
        const int nh = S.nbp_half[r];
        if (nh > S.bp_half_cap) { atomicMax(&S.misc->half_overflow, nh); S.misc->lists_invalid = 1; }
    }
    if (tid == 0 && (S.misc->nl_overflow > S.K || S.misc->bp_overflow > S.KB)) S.misc->lists_invalid = 1;
    __syncthreads();
    if (!s_any || ncta <= 0) return;
    if (local) {     // replica-local mode of k_md_fast: CTA c owns replica c
        for (int c = tid; c <= ncta; c += blockDim.x) S.cta_item0[c] = c * S.nmp;
        return;
    }
    // block scan: thread t owns items [t * chunk, (t + 1) * chunk)
    const int chunk = (nitems + 1023) / 1024;
    const int i0 = tid * chunk, i1 = min(i0 + chunk, nitems);
    int sum = 0;
    for (int i = i0; i < i1; i++) { sum += 2 * ((S.nnbr[i] + 7) >> 3) + 1; s_w[i] = sum; }
    s_part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int t = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += t;
        __syncthreads();
    }
    const int base = s_part[tid] - sum;
    for (int i = i0; i < i1; i++) s_w[i] += base;
    __syncthreads();
    const long long total = s_part[1023];
    // weight cuts: first item whose exclusive prefix reaches c * total / ncta
    for (int c = tid; c <= ncta; c += blockDim.x) {
        const long long want = total * c / ncta;
        int lo = 0, hi = nitems;                 // first i with (i == 0 ? 0 : s_w[i - 1]) >= want
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((mid == 0 ? 0 : s_w[mid - 1]) >= want) hi = mid; else lo = mid + 1;
        }
        s_cut[c] = lo;
    }
    __syncthreads();
    if (tid == 0) {
        int b = 0;
        S.cta_item0[0] = 0;
        for (int c = 1; c < ncta; c++) {
            int nb = min(s_cut[c], b + per_cta);
            nb = max(nb, nitems - (ncta - c) * per_cta);
            nb = max(nb, b);
            S.cta_item0[c] = nb;
            b = nb;
        }
        S.cta_item0[ncta] = nitems;
    }
}

// ---- REMD exchange step on the device (src/energy_replica.F90:45-163, src/job_md.F90:531-567,
// src/replica_exchange.F90:44-154, 175-213) -------------------------------------------------
// The host tables of var_replica / init_replica are mirrored in RemdDev; a label's parameter block
// (temperature, Debye-Hueckel coefficients, tweezers force, Rg0) is precomputed per label, so an
// exchange is a permutation of those blocks.  Every rank replays the same decisions from the
// all-gathered energies and the shared counter-based uniform stream: no second collective.
constexpr int REMD_MAX_PEERS = 16;          // ranks of one node whose exchange buffers are peer-mapped
constexpr int REMD_DEV_MAX_LABELS = 4096;   // replicas the device-side exchange handles (shared table of k_remd_exchange)
struct RemdDev {
    int n, ncol, idx_pair, R, rank;
    int flg_temp, flg_ele_var, flg_twz, flg_rg;   // which fields of a label block replace the replica's
    unsigned long long seed, exchange_counter;
    long long n_attempt, n_accept;
    double tempK_default;
    const int *pair_tb;       // [ncol][n]
    int *rep2lab, *lab2rep;   // [n], 1-based values
    const double *lab_temp;   // [n]
    const RepParams *lab_rp;  // [n]
    RepParams *rp_alt;        // [R] partner-parameter blocks of the cross-energy pass
    int *has_opp;             // [R]
    double *send, *recv;      // [2 R], [2 n]: replica_energies(2, :) of src/job_md.F90:475
    double *vscale;           // [R]
    // exchange inside the replica-local persistent kernel (k_md_fast<U, true>, fast_remd_exchange).
    // Every rank holds one buffer {E[2 parities][n][2] doubles; arrival counter}; the buffers of the other
    // ranks of the node are mapped through CUDA IPC (sisgpu_peer_export / _import), so a replica's CTA stores
    // its two energies straight into every rank's copy over NVLink and bumps every rank's counter: the
    // all-gather of src/job_md.F90:531-533 and its barrier without leaving the kernel.
    int nex;                              // nstep_rep_exchange
    int nranks_x;                         // ranks taking part (1: this rank alone)
    unsigned long long xch_round0;        // exchange rounds done before this launch (parity of E, base of the counters)
    double *xE_peer[REMD_MAX_PEERS];      // [rank] -> that rank's E buffer (own included)
    unsigned long long *xbar_peer[REMD_MAX_PEERS];   // [rank] -> that rank's arrival counter (monotonic, never reset)
};
SIS_D void remd_apply_label(const RemdDev &D, RepParams &dst, const RepParams &lab)
{
    if (D.flg_temp) { dst.tempK = lab.tempK; dst.sqrt_tempK = lab.sqrt_tempK; }
    if (D.flg_ele_var) {
        dst.ele_coef = lab.ele_coef; dst.inv_lambdaD = lab.inv_lambdaD; dst.lambdaD = lab.lambdaD;
        dst.diele_dTcoef = lab.diele_dTcoef;
    }
    if (D.flg_twz) dst.twz_force = lab.twz_force;
    if (D.flg_rg) dst.bias_rg_0 = lab.bias_rg_0;
}
// own total energies -> send(1, :); partner blocks for the cross pass (src/energy_replica.F90:76-118)
__global__ void k_remd_prepare(DevSys S, const RemdDev *Dp)
{
    const RemdDev &D = *Dp;
    const int irep = blockIdx.x * blockDim.x + threadIdx.x;
    if (irep >= D.R) return;
    const int label_own = D.rep2lab[D.rank * D.R + irep];
    const int label_opp = D.pair_tb[(size_t)(D.idx_pair - 1) * D.n + label_own - 1];
    RepParams alt = S.rp[irep];
    int has = 0;
    if (label_opp != 0 && label_opp <= D.n) { has = 1; remd_apply_label(D, alt, D.lab_rp[label_opp - 1]); }
    D.rp_alt[irep] = alt;
    D.has_opp[irep] = has;
    D.send[2 * irep] = S.ene_out[(size_t)irep * ENE_N];
}
__global__ void k_remd_pack(DevSys S, const RemdDev *Dp)
{
    const RemdDev &D = *Dp;
    const int irep = blockIdx.x * blockDim.x + threadIdx.x;
    if (irep < D.R) D.send[2 * irep + 1] = D.has_opp[irep] ? S.ene_out[(size_t)irep * ENE_N] : 0.0;
}
// decisions, label tables, parameter refresh of the local replicas, velocity scale factors of
// accepted temperature swaps.  The pairs of one attempt are disjoint (every label is in at most one
// pair of the current pair table), so they are decided in parallel, one thread per lower label;
// pair k in label order uses uniform k of the attempt, as the reference's sequential loop does.
// One block.
__global__ void __launch_bounds__(256) k_remd_exchange(DevSys S, RemdDev *Dp, int scale_here)
{
    RemdDev &D = *Dp;
    __shared__ int s_rank[REMD_DEV_MAX_LABELS];
    __shared__ int s_natt, s_nacc;
    const int tid = threadIdx.x, n = D.n, R = D.R;
    const int idx_pair = D.idx_pair;
    const unsigned long long counter = D.exchange_counter;
    for (int r = tid; r < R; r += blockDim.x) D.vscale[r] = 1.0;
    for (int l = tid; l < n; l += blockDim.x) {
        const int l_j = D.pair_tb[(size_t)(idx_pair - 1) * n + l];
        s_rank[l] = l_j >= l + 1 ? 1 : 0;
    }
    if (tid == 0) { s_natt = 0; s_nacc = 0; }
    __syncthreads();
    if (tid == 0) {                      // exclusive scan: index of the pair's uniform
        int run = 0;
        for (int l = 0; l < n; l++) { const int v = s_rank[l]; s_rank[l] = v ? run : -1; run += v; }
    }
    __syncthreads();
    for (int l = tid; l < n; l += blockDim.x) {
        const int nrand = s_rank[l];
        if (nrand < 0) continue;
        const int l_i = l + 1, l_j = D.pair_tb[(size_t)(idx_pair - 1) * n + l];
        const double rnd = uniform_counter(D.seed, 0x7fffffffu, counter + (unsigned long long)nrand);
        const int rep_i = D.lab2rep[l_i - 1], rep_j = D.lab2rep[l_j - 1];
        double temp_m, temp_n;
        if (D.flg_temp) { temp_m = D.lab_temp[l_i - 1]; temp_n = D.lab_temp[l_j - 1]; }
        else temp_m = temp_n = D.tempK_default;
        const double *E = D.recv;
        const double pot_i_m = E[2 * (rep_i - 1)], pot_i_n = E[2 * (rep_i - 1) + 1];
        const double pot_j_n = E[2 * (rep_j - 1)], pot_j_m = E[2 * (rep_j - 1) + 1];
        const double delta = ((pot_j_m - pot_i_m) / temp_m - (pot_j_n - pot_i_n) / temp_n) / SIS_BOLTZ_KCAL_MOL;
        atomicAdd(&s_natt, 1);
        if (delta < 0.0 || rnd <= exp(-delta)) {
            atomicAdd(&s_nacc, 1);
            D.rep2lab[rep_i - 1] = l_j; D.rep2lab[rep_j - 1] = l_i;
            D.lab2rep[l_i - 1] = rep_j; D.lab2rep[l_j - 1] = rep_i;
            if (D.flg_temp) {   // scale_velo, src/replica_exchange.F90:175-213
                const double sf = sqrt(temp_n / temp_m);
                const int li = rep_i - 1 - D.rank * R, lj = rep_j - 1 - D.rank * R;
                if (li >= 0 && li < R) D.vscale[li] = sf;
                if (lj >= 0 && lj < R) D.vscale[lj] = 1.0 / sf;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        D.idx_pair = (idx_pair == D.ncol) ? 1 : idx_pair + 1;
        D.exchange_counter = counter + (unsigned long long)((n + 1) / 2);   // the reference draws ceil(n/2) numbers up front
        D.n_attempt += s_natt; D.n_accept += s_nacc;
    }
    for (int r = tid; r < R; r += blockDim.x) {
        RepParams rp = S.rp[r];
        remd_apply_label(D, rp, D.lab_rp[D.rep2lab[D.rank * R + r] - 1]);
        S.rp[r] = rp;
    }
    // small systems: the velocity rescale of accepted temperature swaps right here (one launch less)
    if (scale_here) {
        __syncthreads();                       // vscale complete
        const int n3 = S.nmp * 3;
#pragma unroll 4
        for (int idx = tid; idx < R * n3; idx += blockDim.x) {
            const double f = D.vscale[idx / n3];
            if (f != 1.0) S.vel[idx] *= f;
        }
    }
}
__global__ void k_remd_scale(DevSys S, const RemdDev *Dp)
{
    const int rep = blockIdx.y;
    const double f = Dp->vscale[rep];
    if (f == 1.0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S.nmp * 3) S.vel[(size_t)rep * S.nmp * 3 + i] *= f;
}

// ---- job_dcd batch mode (src/job_dcd.F90:65-79): frames of a DCD evaluated nrep_proc at a time ----
// float32 DCD coordinates of `nf` frames -> the double4 positions of replica slots 0 .. nf-1
__global__ void k_dcd_load(DevSys S, const float *xyz32, int nf)
{
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf || i >= S.nmp) return;
    const float *p = xyz32 + ((size_t)f * S.nmp + i) * 3;
    S.pos[S.ctl[0].cur][(size_t)f * S.nmp + i] = make_double4((double)p[0], (double)p[1], (double)p[2], 0.0);
}

// ---- output staging: what the writers of a save step need, reduced on the device ----
// float32 DCD frame (src/dcd.F90:584-592) and the .bp / .bpall records (src/write_bp.F90:36-49,
// 76-88): base pairs with bp_status and (all or ene_bp <= -kT), as (imp, jmp, real(ene, 4)), 1-based.
struct SnapDev {
    float *xyz32;        // [nmp][3]
    int *bp_cnt;         // [nmp + 1] per-owner record counts, then their exclusive scan
    int2 *bp_ij;         // [cap]
    float *bp_e;         // [cap]
    int *nrec;           // records found (may exceed cap: the host retries with a larger buffer)
    int cap;
};
__global__ void k_snap_frame(DevSys S, SnapDev D, int rep)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.nmp) return;
    const double4 p = S.pos[S.ctl[0].cur][(size_t)rep * S.nmp + i];
    D.xyz32[3 * i] = (float)p.x; D.xyz32[3 * i + 1] = (float)p.y; D.xyz32[3 * i + 2] = (float)p.z;
}
SIS_D bool snap_bp_keep(const DevSys &S, size_t pbase, int2 ent, int k, double kT, int all)
{
    if (ent.x <= k) return false;                       // the owner (lower bead) reports the pair
    if (!S.bp_status[pbase + ent.y]) return false;
    return all || S.ene_bp[pbase + ent.y] <= -kT;
}
__global__ void k_snap_bp_count(DevSys S, SnapDev D, int rep, double kT, int all)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S.nmp) return;
    const size_t rbase = (size_t)rep * S.nmp, pbase = rbase * S.KB;
    const int2 *row = S.bprow + (rbase + k) * S.KB;
    int c = 0;
    for (int e = 0; e < S.nbp_row[rbase + k]; e++) c += snap_bp_keep(S, pbase, row[e], k, kT, all);
    D.bp_cnt[k] = c;
}
__global__ void __launch_bounds__(1024) k_snap_bp_scan(DevSys S, SnapDev D)
{
    __shared__ int wsum[32];
    __shared__ int carry, chunk_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = S.nmp;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = i < n ? D.bp_cnt[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wsum[lane] = wi - w;
            if (lane == 31) chunk_total = wi;
        }
        __syncthreads();
        if (i < n) D.bp_cnt[i] = carry + wsum[warp] + incl - v;
        __syncthreads();
        if (tid == 0) carry += chunk_total;
        __syncthreads();
    }
    if (tid == 0) { D.bp_cnt[n] = carry; *D.nrec = carry; }
}
__global__ void k_snap_bp_fill(DevSys S, SnapDev D, int rep, double kT, int all)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S.nmp) return;
    const size_t rbase = (size_t)rep * S.nmp, pbase = rbase * S.KB;
    const int2 *row = S.bprow + (rbase + k) * S.KB;
    int o = D.bp_cnt[k];
    for (int e = 0; e < S.nbp_row[rbase + k]; e++) {
        const int2 ent = row[e];
        if (!snap_bp_keep(S, pbase, ent, k, kT, all)) continue;
        if (o < D.cap) { D.bp_ij[o] = make_int2(k + 1, ent.x + 1); D.bp_e[o] = (float)S.ene_bp[pbase + ent.y]; }
        o++;
    }
}

// ---- state transfer: xyz(3, nmp) of the Fortran driver <-> the double4 positions ----------------
__global__ void k_pos_unpack(DevSys S, const double *xyz, int rep)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S.nmp) S.pos[S.ctl[0].cur][(size_t)rep * S.nmp + i] = make_double4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0);
}
__global__ void k_pos_pack(DevSys S, double *xyz, int rep)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.nmp) return;
    const double4 p = S.pos[S.ctl[0].cur][(size_t)rep * S.nmp + i];
    xyz[3 * i] = p.x; xyz[3 * i + 1] = p.y; xyz[3 * i + 2] = p.z;
}
