// sis_rep.cuh -- k_md_rep: the MD loop body (src/job_md.F90:374-460) for SMALL replicas (at most 64 beads:
// BASELINE configs C1 and C3, the REMD inputs of check/), sized to the replica instead of to the SM.
//
// One CTA of 128, 256 or 512 threads owns one whole replica for the whole launch; four, two or one CTAs share
// an SM, so a GPU holds 592 / 296 / 148 replicas (the host picks the widest CTA the replica count leaves room
// for: a replica's step is a latency chain that more warps shorten; without an exchange barrier inside the
// kernel more replicas than that run in waves).  Everything a step touches lives in the CTA's shared memory:
// positions, velocities, beta, xyz_move, the STATIC pair table and the STATIC base-pair candidate table.
//
// No neighbour list on the hot path.  With <= 64 beads the list of src/neighbor_list.F90 is a
// superset filter and nothing else: a pair is listed when it is inside cutoff + margin at the last
// rebuild, the force routines then test the cutoff itself (src/force_wca.F90:44, force_ele_DH.F90:55,
// force_bp.F90:64), and the rebuild trigger of src/job_md.F90:376-388 guarantees that no unlisted pair
// can be inside a cutoff.  So every sequence-allowed pair is kept (n (n - 1) / 2 <= 2016 entries with
// the exclusion rules of nl_pair_flags, every pair with bp_lookup_r > 0 as a base-pair candidate) and
// the distance tests of the force routines alone decide -- the same forces as the list path, without
// the per-rebuild prologue that made the 64 replicas of C3 wait for the slowest one at every exchange.
// A rebuild is still an EVENT the dynamics can see, and it is reproduced at the same steps:
// pbc_wrap, xyz_move = 0, the ele_cutoff refresh of src/neighbor_list.F90:80-82, the forced
// Monte-Carlo step of src/job_md.F90:396-401, the rebuild counter.  The global lists other entry
// points read (energies, snapshots, the other kernels) are rebuilt ONCE at the end of the launch from
// the positions saved at the last rebuild event (nl_rows_bead_at, the same code as every other list
// build), and bp_status / bp_status_MC / ene_bp are written back into their slots.
//
// A step = one WORK QUEUE drained by all warps + one assemble / integrate phase, two block barriers.
// The queue: chunks of 32 items drawn from a shared counter, every kind of work on chunk boundaries (a warp runs one
// code path); a chunk that needs earlier chunks waits on a shared-memory counter (stores, fence, counter add on the
// producer side; sleeping spin, fence on the consumer side):
//   gate        distance gate of 32 base-pair candidates -> survivor list
//   dihedrals, bonds + angles (each term ONCE, into one table of force vectors)
//   pairs       64 / 128 pairs per chunk, each pair ONCE, 2 / 4 per lane side by side, branch-free Debye-Hueckel; both
//               directions into [bead][component][partner] slots
//   Gaussians   of the NEXT step (Philox + Box-Muller): nothing in this step waits for them
//   roles       (after the gate) one angle, or one dihedral / the bond, of one survivor per lane: bp_role_eval
//   energies    (after the roles) per survivor: exponent, U0 dG exp(-u), status, Monte-Carlo count;
//   per-bead forces  per (survivor, bead): the unscaled force into a slot listed at the bead
//   owners' sums     (after pairs + bonded) per (bead, component): 128-bit loads of the pair row, bonded entries, tweezers
// then [Monte-Carlo steps: the reference's sequential delete loop on one warp, rep_mc_prune], owners add the
// survivors' forces scaled by the pair energy, integrate (src/job_md.F90:409-456), and the rebuild verdict for the
// next step costs a flag unless a displacement passes half the margin (then every warp reduces the top-2 itself).
// The exchange step of REMD (src/job_md.F90:474-570) runs inside the kernel (rep_remd_exchange): own and cross
// energies item by item, the two numbers into every rank's exchange buffer (NVLink peer stores), two-level arrival
// (local word, then R on every rank's counter), every CTA replays every Metropolis decision on its own label tables.
#pragma once

#include "sis_fast.cuh"

// -DSIS_REP_PROF: per-warp barrier waits, per-chunk-kind cycles of the work queue and the sections of the exchange
// (block 0; tools/rep_profile.py prints them).  The per-CTA phase clocks behind option "profile" are always there.
#ifdef SIS_REP_PROF
constexpr bool REP_PROF = true;
#else
constexpr bool REP_PROF = false;
#endif
constexpr int REP_MAXB = 64;       // beads per replica
constexpr int REP_MAXC = 1024;     // base-pair candidates per replica
constexpr int REP_ACT = 32;        // gate survivors evaluated role by role with slots for their results (beyond: one thread each, atomics)
constexpr int REP_ROLE = BP_ROLE;       // doubles one survivor's seven roles leave: 4 angles x {u, fi, fk}, 2 dihedrals x {u, fi, fl, ra, rb}, bond {u, fi}
constexpr int REP_BPL = 16;        // base-pair force slots listed per bead (beyond: shared-memory atomics)
constexpr int REP_BENT = 9;        // bonded terms a bead takes part in: 2 bonds, 3 angles, 4 dihedrals

// shared-memory carve-up: identical arithmetic on host (size) and device (pointers)
struct RepSm {
    double *x, *v, *a, *m, *fbp, *xreb, *fsum;    // [3 n]
    double *g;                                    // [2][3 n] Gaussians by step parity
    double *ca, *cb, *cc, *nz0, *d2;              // [n]
    double *bf;                                   // [8 n][3]  bond k | angle n + 3 k + p | dihedral 4 n + 4 k + p
    double *Fp;                                   // [n][3][npad]: component c of the force on bead k from bead j at (3 k + c) npad + j (both directions stored; 0 = no pair)
    double *bpF;                                  // [REP_ACT][6][3]
    double *bpR;                                  // [REP_ACT][REP_ROLE]
    double *eA;                                   // [REP_ACT] pair energy of a survivor, 0 = not formed: the owners' scale factor
    double *c_ene, *c_dH, *c_dS;                  // [ccap]  (dG = dH - T dS, src/force_bp_triplet.F90:77, on the fly: rep_dG)
    double *red;                                  // [2 * 16]: two partial sums per warp
    int4 *meta;                                   // [n]
    uint32_t *pl;                                 // [npair]  i | j << 8 | flags (NB_WCA, NB_ELE); 0 = pair excluded
    unsigned short *bent;                         // [n][REP_BENT]  index into bf | 0x8000 (subtract)
    unsigned short *blst;                         // [n][REP_BPL]   index into bpF
    unsigned char *c_b;                           // [ccap][8]  beads i, j, i-1, j-1, i+1, j+1, type, status (bit0 bp_status, bit1 bp_status_MC)
    int *c_pb;                                    // [ccap]  slot in bp_status / ene_bp of the global lists (-1: not listed)
    int *act;                                     // [ccap]
    int *bn, *bcnt, *excess;                      // [n]
    unsigned short *r2l, *l2r;                    // [nlab] label tables of the in-kernel exchange (1-based values, nlab <= REMD_LOCAL_MAX)
    int *ctl;                                     // [16] counters and flags (RC_*)
    RepParams *rp;
    FastRemdState *xs;
    double *xE;                                   // [4]: own, cross, vscale
    size_t bytes;
};
enum { RC_NACT = 0, RC_WORK, RC_NCAND, RC_NEWLAB, RC_NATT, RC_NACC, RC_OVER, RC_FAR0, RC_FAR1, RC_ROLES, RC_STAT, RC_SPILL0, RC_SPILL1, RC_GATE };

SIS_HD RepSm rep_carve(unsigned char *base, int n, int ccap, int nlab)
{
    RepSm s;
    const size_t npair = (size_t)n * (n - 1) / 2;
    size_t o = 0;
    auto take = [&](size_t bytes, size_t align) { o = (o + align - 1) & ~(align - 1); const size_t at = o; o += bytes; return base + at; };
    s.rp = (RepParams *)take(sizeof(RepParams), 16);
    s.xs = (FastRemdState *)take(sizeof(FastRemdState), 16);
    s.meta = (int4 *)take(sizeof(int4) * n, 16);
    s.x = (double *)take(8 * 3 * n, 8); s.v = (double *)take(8 * 3 * n, 8); s.a = (double *)take(8 * 3 * n, 8);
    s.m = (double *)take(8 * 3 * n, 8); s.g = (double *)take(8 * 6 * n, 8); s.fbp = (double *)take(8 * 3 * n, 8);
    s.fsum = (double *)take(8 * 3 * n, 8);
    s.xreb = (double *)take(8 * 3 * n, 8);
    s.ca = (double *)take(8 * n, 8); s.cb = (double *)take(8 * n, 8); s.cc = (double *)take(8 * n, 8);
    s.nz0 = (double *)take(8 * n, 8); s.d2 = (double *)take(8 * REP_MAXB, 8);
    s.bf = (double *)take(8 * 3 * 8 * (size_t)n, 8);
    s.Fp = (double *)take(8 * 3 * (size_t)n * ((n + 1) & ~1), 16);
    s.bpF = (double *)take(8 * 18 * REP_ACT, 8);
    s.bpR = (double *)take(8 * REP_ROLE * REP_ACT, 8);
    s.eA = (double *)take(8 * REP_ACT, 8);
    s.c_ene = (double *)take(8 * ccap, 8);
    s.c_dH = (double *)take(8 * ccap, 8); s.c_dS = (double *)take(8 * ccap, 8);
    s.red = (double *)take(8 * 2 * 16, 8);
    s.xE = (double *)take(8 * 4, 8);
    s.pl = (uint32_t *)take(4 * npair, 4);
    s.c_pb = (int *)take(4 * ccap, 4);
    s.act = (int *)take(4 * ccap, 4);
    s.bn = (int *)take(4 * n, 4); s.bcnt = (int *)take(4 * n, 4); s.excess = (int *)take(4 * n, 4);
    s.ctl = (int *)take(4 * 16, 4);
    s.bent = (unsigned short *)take(2 * REP_BENT * (size_t)n, 2);
    s.blst = (unsigned short *)take(2 * REP_BPL * (size_t)n, 2);
    s.r2l = (unsigned short *)take(2 * (size_t)nlab, 2); s.l2r = (unsigned short *)take(2 * (size_t)nlab, 2);
    s.c_b = (unsigned char *)take(8 * (size_t)ccap, 8);
    s.bytes = (o + 15) & ~size_t(15);
    return s;
}

#ifdef __CUDACC__
// wait until a shared-memory counter reaches `target` (chunks of the work queue that need earlier chunks complete)
SIS_D void rep_wait(const int *ctr, int target, int *flag)
{
    if (*(volatile const int *)ctr < target) {
        const long long t0 = clock64();
        while (*(volatile const int *)ctr < target) {
            __nanosleep(32);
            if (clock64() - t0 > 4000000000ll) { atomicOr(flag, 16); break; }     // ~2 s: never hang the device on a lost chunk
        }
    }
    __threadfence_block();
}
// NT == blockDim.x: the plain CTA barrier (the compiler reconverges the warp in front of it)
template <int NT> SIS_D void rep_sync() { __syncthreads(); }

SIS_D Vec3 rep_x(const double *x, int i) { return mk(x[3 * i], x[3 * i + 1], x[3 * i + 2]); }
SIS_D void rep_put(double *t, int i, Vec3 f) { t[3 * i] = f.x; t[3 * i + 1] = f.y; t[3 * i + 2] = f.z; }
// does the bonded term of `span` bonds that STARTS at the bead with meta m exist (src/list_local.F90:16-119)
SIS_D bool rep_term(int4 m, int span) { return (m.w & MF_CIRC) ? m.y > span : m.z + span < m.y; }
SIS_D int rep_rowbase(int i, int n) { return i * (2 * n - i - 1) / 2; }
SIS_D double rep_dG(const FFParams &P, const RepSm &sm, int c) { return P.bp_triplet ? sm.c_dH[c] - sm.rp->tempK * sm.c_dS[c] : 1.0; }

// one pair of the table: force on its first bead (the second one takes the negative)
SIS_D Vec3 rep_pair_force(const FFParams &P, uint32_t e, const double *x, double ele_coef, double inv_lambdaD, double cut2)
{
    const Vec3 v = pbc_vec_d(P, rep_x(x, e & 0xFF), rep_x(x, (e >> 8) & 0xFF));
    const double d2 = dot(v, v);
    double s = 0.0;
    if (P.mixed) {
        const bool w_on = (e & NB_WCA) && d2 < P.wca_sigma2, e_on = (e & NB_ELE) && d2 <= cut2;
        if (w_on || e_on) s = pair_scalar_f32(P, w_on, e_on, d2, ele_coef, inv_lambdaD);
        return s * v;
    }
    if ((e & NB_WCA) && d2 < P.wca_sigma2) s += wca_scalar(P, d2);
    if ((e & NB_ELE) && d2 <= cut2) {
        const double ri = rsqrt_sl(d2);
        s += ele_coef * ri * ri * (ri + inv_lambdaD) * exp_neg_sl((d2 * ri) * inv_lambdaD);
    }
    return s * v;
}

// survivors beyond the REP_ACT role slots: full evaluation of one candidate by one thread (bp_force_all), forces by atomics
SIS_D void rep_bp_eval(const FFParams &P, const RepSm &sm, int c, bool forces)
{
    unsigned char *b = sm.c_b + 8 * c;
    const int type = (signed char)b[6];
    Vec3 xs[6], fr[6], v12;
#pragma unroll
    for (int t = 0; t < 6; t++) xs[t] = rep_x(sm.x, b[t]);
    double u, a12, ene = 0.0;
    bool on = false;
    if (bp_force_all(P, P.bp[type], xs, &u, fr, &a12, &v12)) {
        ene = P.bp_triplet ? P.bp[type].U0 * rep_dG(P, sm, c) * exp(-u) : P.bp[type].U0 * exp(-u);
        on = ene <= P.bp_cutoff_energy;
    }
    if (on) { b[7] |= 1; sm.c_ene[c] = ene; } else { b[7] &= ~1; sm.c_ene[c] = 0.0; }
    if (!forces) {
        if (on) { atomicAdd(&sm.excess[b[0]], 1); atomicAdd(&sm.excess[b[1]], 1); }
        return;
    }
    if (!on) return;
#pragma unroll
    for (int t = 0; t < 6; t++) {
        const Vec3 F = ene * fr[t];
        const int bead = b[t];
        atomicAdd(&sm.fbp[3 * bead], F.x); atomicAdd(&sm.fbp[3 * bead + 1], F.y); atomicAdd(&sm.fbp[3 * bead + 2], F.z);
    }
}

// Role-major evaluation of the gate survivors (the first REP_ACT of them): one ITEM = one angle or one dihedral of one
// survivor, so a warp runs one code path (the 8-lane groups of k_md_fast run bond / angle / dihedral lanes one after the
// other).  Same expressions as bp_force_role (src/force_bp.F90:60-190); an item leaves {u, fi, fk} or {u, fi, fl, ra, rb}.
SIS_D void rep_bp_role_item(const FFParams &P, const RepSm &sm, int a, int role)
{
    const unsigned char *b = sm.c_b + 8 * sm.act[a];
    const Vec3 x0 = rep_x(sm.x, b[0]), x1 = rep_x(sm.x, b[1]);
    const int tA = bp_role_A(role), tB = bp_role_B(role);
    const Vec3 xa = tA >= 0 ? rep_x(sm.x, b[tA]) : x0, xb = tB >= 0 ? rep_x(sm.x, b[tB]) : x0;
    bp_role_eval(P, P.bp[(signed char)b[6]], role, x0, x1, xa, xb, sm.bpR + REP_ROLE * a);
}
// the six roles of survivor `a` are in -- two independent pieces of work:
// (1) its energy: U0 dG exp(-sum u) (src/force_bp_triplet.F90:200), status, the per-nucleotide count of a Monte-Carlo step;
//     eA[a] = the factor the owners scale the pair's forces with (0 = not formed)
SIS_D void rep_bp_energy(const FFParams &P, const RepSm &sm, int a, bool count_excess)
{
    const int c = sm.act[a];
    unsigned char *b = sm.c_b + 8 * c;
    const BpPara &p = P.bp[(signed char)b[6]];
    const double *r = sm.bpR + REP_ROLE * a;
    const double u = bp_roles_u(r);
    const double ene = P.bp_triplet ? p.U0 * rep_dG(P, sm, c) * exp(-u) : p.U0 * exp(-u);
    const bool on = ene <= P.bp_cutoff_energy;
    if (on) { b[7] |= 1; sm.c_ene[c] = ene; } else { b[7] &= ~1; sm.c_ene[c] = 0.0; }
    sm.eA[a] = on ? ene : 0.0;
    if (on && count_excess) { atomicAdd(&sm.excess[b[0]], 1); atomicAdd(&sm.excess[b[1]], 1); }
}
// (2) the unscaled force on bead t (0..5 = i, j, i-1, j-1, i+1, j+1) of survivor `a` into its slot, listed at the bead
SIS_D void rep_bp_gather(const FFParams &P, const RepSm &sm, int a, int t, int spill_slot)
{
    const unsigned char *b = sm.c_b + 8 * sm.act[a];
    const double *r = sm.bpR + REP_ROLE * a;
    const Vec3 F = bp_roles_bead_force(r, t);
    const int bead = b[t];
    const int slot = atomicAdd(&sm.bcnt[bead], 1);
    rep_put(sm.bpF, a * 6 + t, F);
    if (slot < REP_BPL) sm.blst[bead * REP_BPL + slot] = (unsigned short)(a * 6 + t);
    else atomicOr(&sm.ctl[spill_slot], 1);      // more pairs at one bead than list slots: the owners scan every survivor
}

// the distance gate of 32 candidates (one lane each): survivors -> act[], the others lose bp_status.
// |r - r0| <= cutoff_ddist (src/force_bp.F90:64) tested on r^2.  use_mask: only pairs with bp_status_MC
SIS_D void rep_bp_gate_warp(const FFParams &P, const RepSm &sm, int c, int ncand, bool use_mask)
{
    const int lane = threadIdx.x & 31;
    bool pass = false;
    if (c < ncand) {
        unsigned char *b = sm.c_b + 8 * c;
        const int type = (signed char)b[6];
        if ((!use_mask || (b[7] & 2)) && (!P.bp_triplet || rep_dG(P, sm, c) < 0.0)) {
            const Vec3 v12 = pbc_vec_d(P, rep_x(sm.x, b[0]), rep_x(sm.x, b[1]));
            const double d2 = dot(v12, v12), r0 = P.bp[type].bond_r, dc = P.bp[type].cutoff_ddist;
            const double lo = fmax(r0 - dc, 0.0), hi = r0 + dc;
            pass = d2 >= lo * lo && d2 <= hi * hi;
        }
        if (!pass) { b[7] &= ~1; sm.c_ene[c] = 0.0; }
    }
    const unsigned mb = __ballot_sync(0xffffffffu, pass);
    if (mb) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&sm.ctl[RC_NACT], __popc(mb));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) sm.act[base + __popc(mb & ((1u << lane) - 1u))] = c;
    }
}

// Monte-Carlo step (src/force_bp_limit_triplet.F90:85-328).  Pass 1 (every candidate evaluated, excess of formed pairs
// per nucleotide) runs in the step's work queue; this is pass 2 = the reference's sequential delete loop on warp 0
// (bp_prune_warp's arithmetic and uniform stream, on the shared-memory tables: a nucleotide's formed pairs in ascending
// partner order = table order; the uniforms of one deletion are drawn side by side, one per lane); then the forces of
// the pairs that stay, and bp_status_MC = bp_status (src/job_md.F90:399-401)
template <int NT>
SIS_D void rep_mc_prune(const FFParams &P, const RepSm &sm, int n, int ncand, unsigned long long seed, int max_bp_per_nt, unsigned long long *mc_ctr,
                         unsigned long long *prof_slot)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const int nact = sm.ctl[RC_NACT];
    if (tid < 32) {
        const RepParams &R = *sm.rp;
        const double beta = 1.0 / (SIS_BOLTZ_KCAL_MOL * R.tempK);
        unsigned long long ctr = *mc_ctr;
        const uint32_t stream = 0x40000000u + (uint32_t)R.stream;
        int *seq = reinterpret_cast<int *>(sm.bpF);          // the slots are free until the force pass
        int ndel = 0;
        unsigned long long wbase = ctr - 64ull;     // no window yet: the first deletion draws it
        double umine = 0.0;
        for (;;) {
            const bool e0 = lane < n && sm.excess[lane] > 0, e1 = lane + 32 < n && sm.excess[lane + 32] > 0;
            const unsigned m0 = __ballot_sync(0xffffffffu, e0), m1 = __ballot_sync(0xffffffffu, e1);
            const int nnt = __popc(m0) + __popc(m1);
            if (nnt == 0) break;
            // uniforms ctr, ctr + 1, ...: [0] picks the nucleotide, [1 .. nseq] shuffle, [nseq + 1 .. 2 nseq - 1] Metropolis sweep.
            // A window of 32 of the stream is drawn side by side (one per lane) and serves several deletions
            if (ctr - wbase >= 24ull) { wbase = ctr; umine = uniform_counter(seed, stream, wbase + (unsigned long long)lane); }
            auto uni = [&](int idx) {
                const unsigned long long off = ctr + (unsigned long long)idx - wbase;
                double u = __shfl_sync(0xffffffffu, umine, (int)(off & 31ull));
                if (off >= 32ull) u = uniform_counter(seed, stream, ctr + (unsigned long long)idx);
                return u;
            };
            double rnd = uni(0);
            const int pick = max(1, min((int)ceil(rnd * nnt), nnt));
            const int nt_delete = pick <= __popc(m0) ? (int)__fns(m0, 0, pick) : 32 + (int)__fns(m1, 0, pick - __popc(m0));
            // formed pairs of the nucleotide in table order
            int nseq = 0;
            for (int c0 = 0; c0 < ncand; c0 += 32) {
                const int c = c0 + lane;
                bool hit = false;
                if (c < ncand) { const unsigned char *b = sm.c_b + 8 * c; hit = (b[7] & 1) && (b[0] == nt_delete || b[1] == nt_delete); }
                const unsigned mh = __ballot_sync(0xffffffffu, hit);
                const int at = nseq + __popc(mh & ((1u << lane) - 1u));
                if (hit && at < REP_MAXB) seq[at] = c;
                nseq += __popc(mh);
            }
            nseq = min(nseq, REP_MAXB);
            __syncwarp();
            if (nseq == 0) {                        // cannot happen (excess > 0 means formed pairs); never spin on it
                if (lane == 0) sm.excess[nt_delete] = 0;
                ctr += 1ull;
                __syncwarp();
                continue;
            }
            int used = 0;
            for (int i = 0; i < nseq; i++) {       // the reference's shuffle, :295-302 (every lane runs it: uniform)
                rnd = uni(1 + used++);
                const int sw = max(1, min((int)ceil(rnd * nseq), nseq));
                if (lane == 0) { const int t = seq[i]; seq[i] = seq[sw - 1]; seq[sw - 1] = t; }
                __syncwarp();
            }
            int del = seq[0];
            for (int i = 1; i < nseq; i++) {       // Metropolis sweep, :305-316
                const double ratio = exp((sm.c_ene[seq[i]] - sm.c_ene[del]) * beta);
                rnd = uni(1 + used++);
                if (rnd < ratio) del = seq[i];
            }
            if (lane == 0) {
                unsigned char *b = sm.c_b + 8 * del;
                b[7] &= ~1;
                sm.excess[b[0]] -= 1;
                sm.excess[b[1]] -= 1;
            }
            ctr += 1ull + (unsigned long long)used;
            ndel++;
            __syncwarp();
        }
        if (lane == 0) *mc_ctr = ctr;
        if (lane == 0 && prof_slot) *prof_slot += (unsigned long long)ndel;
    }
    rep_sync<NT>();
    for (int a = tid; a < nact; a += NT) {      // the pairs that stay keep their forces
        const bool on = sm.c_b[8 * sm.act[a] + 7] & 1;
        if (a < REP_ACT) { if (!on) sm.eA[a] = 0.0; }
        else if (on) rep_bp_eval(P, sm, sm.act[a], true);
    }
    rep_sync<NT>();
    for (int c = tid; c < ncand; c += NT) {
        unsigned char *b = sm.c_b + 8 * c;
        b[7] = (unsigned char)((b[7] & 1) * 3);
    }
    for (int i = tid; i < n; i += NT) sm.excess[i] = -max_bp_per_nt;     // for the next Monte-Carlo step
    rep_sync<NT>();
}

// energy_replica (src/energy_replica.F90:45-163) for the CTA's replica on the shared-memory state: total energy
// with its own parameters (bp energies as cached by the last force evaluation: flg_bp_energy) and with the partner
// label's parameters (bp term recomputed on the energy side, :72, :121) -> sm.xE[0], sm.xE[1].
// The cross base-pair energies follow the force path's shape: distance gate, survivors compacted, one angle or
// dihedral of one survivor per item (mp_angle / mp_dihedral of src/energy_bp.F90:76-118), then the exponent per survivor.
template <int NT>
SIS_D void rep_remd_energies(const DevSys &S, const FFParams &P, const RepSm &sm, int n, int npair, int ncand,
                             const RepParams &Rown, const RepParams &Ralt, bool has_opp)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const bool dih_on = P.flg_dih_cos || P.flg_dih_exp;
    const double tK_alt = P.temp_independent ? 0.0 : Ralt.tempK;     // src/energy_bp_triplet.F90:27-31
    double eo = 0.0, ea = 0.0;
    // gate with the partner's dG (k_bp_eval<1>); own bp energy = the cached one
    for (int c0 = 0; c0 < ncand; c0 += NT) {
        const int c = c0 + tid;
        bool pass = false;
        if (c < ncand) {
            const unsigned char *b = sm.c_b + 8 * c;
            if (b[7] & 1) eo += sm.c_ene[c];
            if (has_opp && (!P.bp_triplet || sm.c_dH[c] - tK_alt * sm.c_dS[c] < 0.0)) {
                const BpPara &p = P.bp[(signed char)b[6]];
                const Vec3 v = pbc_vec_d(P, rep_x(sm.x, b[0]), rep_x(sm.x, b[1]));
                pass = fabs(sqrt(dot(v, v)) - p.bond_r) <= p.cutoff_ddist;
            }
        }
        const unsigned mb = __ballot_sync(0xffffffffu, pass);
        if (mb) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&sm.ctl[RC_NACT], __popc(mb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) sm.act[base + __popc(mb & ((1u << lane) - 1u))] = c;
        }
    }
    rep_sync<NT>();
    const int nact = sm.ctl[RC_NACT], nar = min(nact, REP_ACT);
    double *uR = sm.bpR;        // [nar][6]: free between two force evaluations
    for (int it = tid; it < 6 * nar + (nact - nar) + 3 * n + npair; it += NT) {
        if (it < 6 * nar) {
            const int a = it / 6, r = it - 6 * a;
            const unsigned char *b = sm.c_b + 8 * sm.act[a];
            const BpPara &p = P.bp[(signed char)b[6]];
            const Vec3 x0 = rep_x(sm.x, b[0]), x1 = rep_x(sm.x, b[1]);
            double u;
            if (r < 4) {
                // angles 1=2-4, 3-1=2, 1=2-6, 5-1=2 in the order of bp_energy_u
                const double t = r == 0 ? mp_angle(P, x0, x1, rep_x(sm.x, b[3])) : r == 1 ? mp_angle(P, rep_x(sm.x, b[2]), x0, x1)
                               : r == 2 ? mp_angle(P, x0, x1, rep_x(sm.x, b[5])) : mp_angle(P, rep_x(sm.x, b[4]), x0, x1);
                const double k = r == 0 ? p.angl_k1 : r == 1 ? p.angl_k2 : r == 2 ? p.angl_k3 : p.angl_k4;
                const double th = r == 0 ? p.angl_theta1 : r == 1 ? p.angl_theta2 : r == 2 ? p.angl_theta3 : p.angl_theta4;
                u = k * (t - th) * (t - th);
            } else {
                const double t = r == 4 ? mp_dihedral(P, rep_x(sm.x, b[2]), x0, x1, rep_x(sm.x, b[3]))
                                        : mp_dihedral(P, rep_x(sm.x, b[4]), x0, x1, rep_x(sm.x, b[5]));
                u = r == 4 ? p.dihd_k1 * (1.0 + cos(t + p.dihd_phi1)) : p.dihd_k2 * (1.0 + cos(t + p.dihd_phi2));
            }
            uR[it] = u;
        } else if (it < 6 * nar + (nact - nar)) {       // survivors beyond the slots: one thread each
            const int c = sm.act[REP_ACT + it - 6 * nar];
            const unsigned char *b = sm.c_b + 8 * c;
            const int type = (signed char)b[6];
            Vec3 xs[6];
#pragma unroll
            for (int t = 0; t < 6; t++) xs[t] = rep_x(sm.x, b[t]);
            double u;
            if (bp_energy_u(P, P.bp[type], xs, &u)) {
                const double dG = P.bp_triplet ? sm.c_dH[c] - tK_alt * sm.c_dS[c] : 1.0;
                const double ene = P.bp_triplet ? P.bp[type].U0 * dG * exp(-u) : P.bp[type].U0 * exp(-u);
                if (ene <= P.bp_cutoff_energy) ea += ene;
            }
        } else if (it < 6 * nar + (nact - nar) + 3 * n) {
            const int t = it - 6 * nar - (nact - nar), span = 3 - t / n, k = t % n;       // dihedrals first
            const int4 m = sm.meta[k];
            if (!rep_term(m, span) || (span == 3 && !dih_on)) continue;
            const Vec3 x0 = rep_x(sm.x, k), x1 = rep_x(sm.x, chain_nb(m, 1));
            double w;
            if (span == 1) w = bond_energy(P, x0, x1);
            else if (span == 2) w = angl_energy(P, x0, x1, rep_x(sm.x, chain_nb(m, 2)));
            else w = dih_energy(P, x0, x1, rep_x(sm.x, chain_nb(m, 2)), rep_x(sm.x, chain_nb(m, 3)));
            eo += w; ea += w;
        } else {
            const uint32_t e = sm.pl[it - 6 * nar - (nact - nar) - 3 * n];
            if (!e) continue;
            const Vec3 v = pbc_vec_d(P, rep_x(sm.x, e & 0xFF), rep_x(sm.x, (e >> 8) & 0xFF));
            const double d2 = dot(v, v);
            if (e & NB_WCA) { const double w = wca_energy(P, d2); eo += w; ea += w; }
            if ((e & NB_ELE) && P.flg_ele) { eo += dh_energy(P, Rown, d2); if (has_opp) ea += dh_energy(P, Ralt, d2); }
        }
    }
    if (tid == 0)      // the only bias this kernel admits: dual constant-force tweezers (src/energy_tweezers.F90:21-31)
        for (int t = 0; t < P.ntwz_DCF; t++) {
            const Vec3 v = pbc_vec_d(P, rep_x(sm.x, S.twz_pairs[2 * t]), rep_x(sm.x, S.twz_pairs[2 * t + 1]));
            const double proj = v.x * S.twz_dir[3 * t] + v.y * S.twz_dir[3 * t + 1] + v.z * S.twz_dir[3 * t + 2];
            eo -= Rown.twz_force * proj; ea -= Ralt.twz_force * proj;
        }
    rep_sync<NT>();
    for (int a = tid; a < nar; a += NT) {               // exponent and energy of a survivor
        const int c = sm.act[a];
        const unsigned char *b = sm.c_b + 8 * c;
        const BpPara &p = P.bp[(signed char)b[6]];
        const Vec3 v = pbc_vec_d(P, rep_x(sm.x, b[0]), rep_x(sm.x, b[1]));
        const double d = sqrt(dot(v, v)) - p.bond_r;
        const double *ur = uR + 6 * a;
        const double u = p.bond_k * d * d + ur[0] + ur[1] + ur[2] + ur[3] + ur[4] + ur[5];     // the order of bp_energy_u
        const double dG = P.bp_triplet ? sm.c_dH[c] - tK_alt * sm.c_dS[c] : 1.0;
        const double ene = P.bp_triplet ? p.U0 * dG * exp(-u) : p.U0 * exp(-u);
        if (ene <= P.bp_cutoff_energy) ea += ene;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { eo += __shfl_xor_sync(0xffffffffu, eo, o); ea += __shfl_xor_sync(0xffffffffu, ea, o); }
    if (lane == 0) { sm.red[2 * (tid >> 5)] = eo; sm.red[2 * (tid >> 5) + 1] = ea; }
    rep_sync<NT>();
    if (tid == 0) {
        double so = 0.0, sa = 0.0;
        for (int w = 0; w < NT / 32; w++) { so += sm.red[2 * w]; sa += sm.red[2 * w + 1]; }
        sm.xE[0] = so;
        sm.xE[1] = has_opp ? sa : 0.0;
        sm.ctl[RC_NACT] = 0;
    }
    rep_sync<NT>();
}

// One exchange attempt (src/job_md.F90:474-481,514-570): see fast_remd_exchange -- the same protocol (energies and
// arrival counters into every rank's exchange buffer, every CTA replays every decision), on this kernel's tables.
// Afterwards the on-chip state follows the new label: velocities (scale_velo), T / Debye-Hueckel parameters /
// tweezers force (set_md_coef, set_ele; ele_cutoff stays: only neighbor_list refreshes it), dG of the candidates.
template <int NT>
SIS_D void rep_remd_exchange(const DevSys &S, const FFParams &P, const RepSm &sm, int rep, int n, int npair, int ncand)
{
    const FastRemdState X = *sm.xs;
    RemdDev &D = *S.remd;
    const int tid = threadIdx.x, nl = D.n, R = D.R;
    const int grep = D.rank * R + rep;
    const int label_own = sm.r2l[grep];
    const int label_opp = D.pair_tb[(size_t)(X.idx_pair - 1) * nl + label_own - 1];
    const bool has_opp = label_opp != 0 && label_opp <= nl;
    const RepParams Rown = *sm.rp;
    RepParams Ralt = Rown;
    if (has_opp) remd_apply_label(D, Ralt, D.lab_rp[label_opp - 1]);
    const bool xprof = REP_PROF && S.prof != nullptr && blockIdx.x == 0 && tid == 0;
    unsigned long long *xp = S.prof + 16 + 8 * 1003;
    long long x0 = xprof ? clock64() : 0, x1;
    rep_remd_energies<NT>(S, P, sm, n, npair, ncand, Rown, Ralt, has_opp);
    if (xprof) { x1 = clock64(); xp[0] += (unsigned long long)(x1 - x0); x0 = x1; }
    const unsigned long long round = D.xch_round0 + (unsigned long long)X.attempt;
    const size_t eoff = (size_t)(round & 1ull) * 2 * nl;
    const double *E = D.xE_peer[D.rank] + eoff;
    if (tid == 0) {
        for (int r = 0; r < D.nranks_x; r++) {
            double *Er = D.xE_peer[r] + eoff;
            Er[2 * grep] = sm.xE[0];
            Er[2 * grep + 1] = sm.xE[1];
        }
        if (D.nranks_x > 1) {
            // two-level arrival: the replicas of this rank count themselves on a local word (the slot behind the
            // arrival counter); the last one of the round adds R to every rank's counter -- nranks remote atomics per
            // rank and round on a peer's counter instead of one per replica (512 on one address at 8 x 64 replicas)
            __threadfence_system();
            const unsigned long long mine = atomicAdd(D.xbar_peer[D.rank] + 1, 1ull) + 1ull;
            if (mine == (round + 1ull) * (unsigned long long)R) {
                __threadfence_system();
                for (int r = 0; r < D.nranks_x; r++) atomicAdd_system(D.xbar_peer[r], (unsigned long long)R);
            }
        } else {
            __threadfence();
            atomicAdd(D.xbar_peer[0], 1ull);
        }
        const unsigned long long target = (round + 1ull) * (unsigned long long)nl;     // one CTA per replica on the whole node
        const long long t0 = clock64();
        if (xprof) { xp[1] += (unsigned long long)(t0 - x0); }
        while (ld_acquire_sys_u64(D.xbar_peer[D.rank]) < target) {     // (sleeping: the SM's other replicas keep their issue slots)
            __nanosleep(64);
            if (clock64() - t0 > 8000000000ll) { atomicOr(&S.misc->max_force_flag, 16); break; }   // ~4 s: a CTA / a rank is missing
        }
        if (REP_PROF && S.prof) S.prof[16 + 8 * (blockIdx.x % FAST_MAX_CTA) + 6] += (unsigned long long)(clock64() - t0);
        sm.ctl[RC_NEWLAB] = label_own; sm.xE[2] = 1.0; sm.ctl[RC_NATT] = 0; sm.ctl[RC_NACC] = 0;
        if (xprof) { x0 = clock64(); xp[2] += (unsigned long long)(x0 - t0); }
    }
    rep_sync<NT>();
    {   // replay of the decisions: one lane per label, the blocks of 32 labels on the warps side by side (the pairs are
        // disjoint); pair k in label order uses uniform k
        const int *tb = D.pair_tb + (size_t)(X.idx_pair - 1) * nl;
        const int lane = tid & 31;
        for (int l0 = 32 * (tid >> 5); l0 < nl; l0 += NT) {
            int base = 0;
            for (int b0 = 0; b0 < l0; b0 += 32) base += __popc(__ballot_sync(0xffffffffu, tb[b0 + lane] >= b0 + lane + 1));
            const int l = l0 + lane;
            const int l_j = l < nl ? tb[l] : 0;
            const bool is_pair = l < nl && l_j >= l + 1;
            const unsigned m = __ballot_sync(0xffffffffu, is_pair);
            if (is_pair) {
                const int nrand = base + __popc(m & ((1u << lane) - 1u));
                const int l_i = l + 1;
                const double rnd = uniform_counter(D.seed, 0x7fffffffu, X.counter + (unsigned long long)nrand);
                const int rep_i = sm.l2r[l_i - 1], rep_j = sm.l2r[l_j - 1];
                double temp_m, temp_n;
                if (D.flg_temp) { temp_m = D.lab_temp[l_i - 1]; temp_n = D.lab_temp[l_j - 1]; }
                else temp_m = temp_n = D.tempK_default;
                const double pot_i_m = __ldcg(E + 2 * (rep_i - 1)), pot_i_n = __ldcg(E + 2 * (rep_i - 1) + 1);
                const double pot_j_n = __ldcg(E + 2 * (rep_j - 1)), pot_j_m = __ldcg(E + 2 * (rep_j - 1) + 1);
                const double delta = ((pot_j_m - pot_i_m) / temp_m - (pot_j_n - pot_i_n) / temp_n) / SIS_BOLTZ_KCAL_MOL;
                atomicAdd(&sm.ctl[RC_NATT], 1);
                if (delta < 0.0 || rnd <= exp(-delta)) {
                    atomicAdd(&sm.ctl[RC_NACC], 1);
                    sm.r2l[rep_i - 1] = (unsigned short)l_j; sm.r2l[rep_j - 1] = (unsigned short)l_i;
                    sm.l2r[l_i - 1] = (unsigned short)rep_j; sm.l2r[l_j - 1] = (unsigned short)rep_i;
                    const double sf = D.flg_temp ? sqrt(temp_n / temp_m) : 1.0;   // scale_velo
                    if (rep_i - 1 == grep) { sm.ctl[RC_NEWLAB] = l_j; sm.xE[2] = sf; }
                    if (rep_j - 1 == grep) { sm.ctl[RC_NEWLAB] = l_i; sm.xE[2] = 1.0 / sf; }
                }
            }
        }
    }
    rep_sync<NT>();
    if (xprof) { x1 = clock64(); xp[3] += (unsigned long long)(x1 - x0); x0 = x1; }
    if (tid == 0) {     // parameter refresh of the replica (src/job_md.F90:538-567)
        FastRemdState Y = X;
        Y.natt += sm.ctl[RC_NATT]; Y.nacc += sm.ctl[RC_NACC];
        Y.idx_pair = (X.idx_pair == D.ncol) ? 1 : X.idx_pair + 1;       // set_forward
        Y.counter += (unsigned long long)((nl + 1) / 2);                // the reference draws ceil(n / 2) numbers up front
        Y.attempt++;
        *sm.xs = Y;
        RepParams rp = *sm.rp;
        remd_apply_label(D, rp, D.lab_rp[sm.ctl[RC_NEWLAB] - 1]);
        *sm.rp = rp;
    }
    rep_sync<NT>();
    const double vs = sm.xE[2];
    for (int i = tid; i < 3 * n; i += NT) sm.v[i] *= vs;
    rep_sync<NT>();
    if (xprof) { xp[4] += (unsigned long long)(clock64() - x0); xp[5] += 1ull; }
}

template <int NT, bool XCH>
__global__ void __launch_bounds__(NT, NT >= 512 ? 1 : 512 / NT) k_md_rep(const __grid_constant__ DevSys S, const __grid_constant__ FFParams P, StepArgs A,
                                                          long long target, NlArgs NA, int ccap, int nlab)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_mc_ctr;
    const DevCtl ctl = S.ctl[0];
    const int nsteps = md_steps_to_run(S, ctl, target);
    if (nsteps <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = S.nmp, KB = S.KB, rep = blockIdx.x;
    const int npair = n * (n - 1) / 2;
    const size_t rbase = (size_t)rep * n, pbase = rbase * KB;
    const RepSm sm = rep_carve(smem_raw, n, ccap, nlab);
    const bool dih_on = P.flg_dih_cos || P.flg_dih_exp;
    const bool use_mc = P.nstep_bp_MC > 0;
    const bool xch = XCH && S.remd != nullptr;

    // ---- prologue: state and static tables -> shared memory ---------------------------------------------------
    for (int t = tid; t < 3 * n; t += NT) {
        const int k = t / 3, c = t - 3 * k;
        const double4 p = S.pos[ctl.cur][rbase + k];
        sm.x[t] = c == 0 ? p.x : c == 1 ? p.y : p.z;
        sm.v[t] = S.vel[rbase * 3 + t]; sm.a[t] = S.acc[rbase * 3 + t]; sm.m[t] = S.move[rbase * 3 + t];
        sm.fbp[t] = 0.0;
        sm.xreb[t] = 0.0;
    }
    for (int k = tid; k < n; k += NT) {
        const double4 cf = S.coef[k];
        sm.ca[k] = cf.x; sm.cb[k] = cf.y; sm.cc[k] = cf.z; sm.nz0[k] = cf.w;
        sm.meta[k] = S.meta[k];
        sm.bcnt[k] = 0;
        sm.excess[k] = -A.max_bp_per_nt;
    }
    for (int k = tid; k < REP_MAXB; k += NT) sm.d2[k] = 0.0;
    for (int t = tid; t < 3 * n * ((n + 1) & ~1); t += NT) sm.Fp[t] = 0.0;
    if (tid == 0) {
        *sm.rp = S.rp[rep];
        for (int i = 0; i < 16; i++) sm.ctl[i] = 0;
        s_mc_ctr = use_mc ? S.mc_counter[rep] : 0ull;
        if (xch) {
            const RemdDev &D = *S.remd;
            FastRemdState X;
            X.idx_pair = D.idx_pair; X.counter = D.exchange_counter; X.attempt = 0; X.natt = 0; X.nacc = 0;
            *sm.xs = X;
        }
    }
    if (xch) for (int i = tid; i < S.remd->n; i += NT) { sm.r2l[i] = (unsigned short)S.remd->rep2lab[i]; sm.l2r[i] = (unsigned short)S.remd->lab2rep[i]; }
    rep_sync<NT>();
    // pair table: every pair the exclusion rules of src/neighbor_list.F90:122-180 admit (flags), 0 for the others
    for (int i = warp; i + 1 < n; i += NT / 32) {
        const int4 mi = sm.meta[i];
        const int base = rep_rowbase(i, n);
        for (int j = i + 1 + lane; j < n; j += 32) {
            const uint32_t ent = nl_pair_flags(NA, P, mi, sm.meta[j], i, j, 0.0, 1.0);
            sm.pl[base + j - i - 1] = ent ? ((uint32_t)i | ((uint32_t)j << 8) | ent) : 0u;
        }
    }
    // bonded terms a bead takes part in: position p of the term that starts at the bead p places before it
    for (int k = tid; k < n; k += NT) {
        const int4 m = sm.meta[k];
        int ne = 0;
        unsigned short *be = sm.bent + REP_BENT * k;
        for (int span = 1; span <= (dih_on ? 3 : 2); span++)
            for (int p = 0; p <= span; p++) {
                const int s = chain_nb(m, -p);
                if (s < 0 || !rep_term(sm.meta[s], span)) continue;
                if (span == 1) be[ne++] = (unsigned short)(s | (p == 0 ? 0x8000 : 0));        // bond vector f: F(first) = -f, F(second) = +f
                else if (span == 2) be[ne++] = (unsigned short)(n + 3 * s + p);
                else be[ne++] = (unsigned short)(4 * n + 4 * s + p);
            }
        sm.bn[k] = ne;
    }
    if (warp == NT / 32 - 1) {
        // base-pair candidates in (i, j) order: every pair of bp_map (src/init_bp.F90:184-313)
        int nc = 0;
        for (int i = 0; i + 1 < n; i++) {
            const int4 mi = sm.meta[i];
            const int nrow = S.nbp_row[rbase + i];
            for (int j0 = i + 1; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                int bp3 = 0, bptype = 0;
                int4 mj = mi;
                if (j < n) { mj = sm.meta[j]; bp3 = bp_lookup_r(S.rs, NA.bp_model, NA.bp_min_loop, S.hash, S.seq, i, mi, j, mj, &bptype); }
                const unsigned mb = __ballot_sync(0xffffffffu, bp3 > 0);
                if (bp3 > 0) {
                    const int c = nc + __popc(mb & ((1u << lane) - 1u));
                    if (c < ccap) {
                        unsigned char *b = sm.c_b + 8 * c;
                        b[0] = (unsigned char)i; b[1] = (unsigned char)j;
                        b[2] = (unsigned char)chain_nb(mi, -1); b[3] = (unsigned char)chain_nb(mj, -1);
                        b[4] = (unsigned char)chain_nb(mi, +1); b[5] = (unsigned char)chain_nb(mj, +1);
                        b[6] = (unsigned char)bptype;
                        int pb = -1;
                        for (int e = 0; e < nrow; e++) {
                            const int2 ent = S.bprow[(rbase + i) * KB + e];
                            if (ent.x == j) { pb = ent.y; break; }
                        }
                        sm.c_pb[c] = pb;
                        b[7] = pb >= 0 ? (unsigned char)((S.bp_status[pbase + pb] ? 1 : 0) | (S.bp_status_MC[pbase + pb] ? 2 : 0)) : 0;
                        sm.c_ene[c] = pb >= 0 ? S.ene_bp[pbase + pb] : 0.0;
                        const bool m5 = NA.bp_model == 5;
                        sm.c_dH[c] = m5 ? S.bp3_dH[bp3] : 0.0;
                        sm.c_dS[c] = m5 ? S.bp3_dS[bp3] : 0.0;
                    }
                }
                nc += __popc(mb);
            }
        }
        if (lane == 0) { sm.ctl[RC_NCAND] = min(nc, ccap); if (nc > ccap) { sm.ctl[RC_OVER] = 1; atomicOr(&S.misc->max_force_flag, 4); } }
    }
    // Gaussians of the first step (every later step finds its own drawn one step ahead)
    for (int it = tid; it < 2 * n; it += NT) {
        const int k = it >> 1, draw = it & 1;
        double g0, g1;
        gauss_pair(A.seed, (uint32_t)S.rp[rep].stream, ctl.step + 1ull, (uint32_t)k, (uint32_t)draw, &g0, &g1);
        double *gn = sm.g + 3 * n * (int)((ctl.step + 1ull) & 1ull);
        if (draw == 0) { gn[3 * k] = g0; gn[3 * k + 1] = g1; } else gn[3 * k + 2] = g0;
    }
    // rebuild verdict on the displacements the previous launch left (src/job_md.F90:376-388)
    bool blocked;
    double top1, top2v;
    {
        const unsigned long long *t = S.top2 + ((size_t)rep * 2 + (ctl.step & 1ull)) * 2;
        top1 = __longlong_as_double((long long)__ldcg(t));
        top2v = __longlong_as_double((long long)__ldcg(t + 1));
        blocked = sqrt(top1) + sqrt(top2v) > S.nl_margin;
    }
    const double d2_any = 0.25 * S.nl_margin * S.nl_margin;     // sqrt(t1) + sqrt(t2) > margin needs t1 > (margin / 2)^2
    int nreb = 0;
    bool after_rebuild = false;
    rep_sync<NT>();
    const int ncand = sm.ctl[RC_NCAND];
    const bool overflow = sm.ctl[RC_OVER] != 0;
    const uint32_t stream = (uint32_t)sm.rp->stream;
    // work of P1 in chunks of one warp: base-pair survivors, Gaussians, dihedrals, bond + angle, then the pairs (64 per chunk)
    // (each kind starts on a chunk boundary: a warp runs one code path)
    constexpr int PPL = NT >= 512 ? 2 : 4;       // pairs per lane of a pair chunk
    const int cD = dih_on ? (n + 31) >> 5 : 0, cB = (n + 31) >> 5, cpair = (npair + 32 * PPL - 1) / (32 * PPL), cG = (2 * n + 31) >> 5, cA = (3 * n + 31) >> 5;
    const int cgate = (ncand + 31) >> 5;
    const int npad = (n + 1) & ~1;               // row length of the pair slots (16-byte rows)
    const int cstat = cD + cB + cpair;          // chunks the owners' sums wait for
    // optional phase clocks (option "profile"): per CTA {P0, P1 wait for the first chunk, P1 work, P2, verdict, exchange, exchange wait, steps}
    const bool prof = S.prof != nullptr && tid == 0;
    unsigned long long *pc = S.prof + 16 + 8 * (blockIdx.x % FAST_MAX_CTA);
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
    const bool wprof = REP_PROF && S.prof != nullptr && blockIdx.x == 0 && lane == 0 && warp < 4;     // block 0, warps 0-3: {B2 wait, P2, B3 wait}
    long long w0 = 0, w1 = 0, aw[3] = {0, 0, 0};
    const bool tprof = REP_PROF && S.prof != nullptr && blockIdx.x == 0;      // block 0: cycles and count per chunk type of the P1 queue

    int done = 0;
    if (!overflow)
    for (; done < nsteps; done++) {
        const unsigned long long istep = ctl.step + (unsigned long long)done + 1ull;
        if (prof) c0 = clock64();
        // ---- P0: rebuild event, Monte-Carlo pass, distance gate ------------------------------------------------
        if (blocked) {
            for (int t = tid; t < 3 * n; t += NT) {
                double xv = sm.x[t];
                if (P.flg_pbc) { xv = wrap1(xv, P.box[t % 3]); sm.x[t] = xv; }
                sm.xreb[t] = xv;
                sm.m[t] = 0.0;
            }
            if (tid == 0 && NA.ele_cutoff_type == 2) sm.rp->ele_cutoff = NA.ele_cutoff_inp * sm.rp->lambdaD;
            nreb++;
            after_rebuild = true;
            rep_sync<NT>();
        }
        const bool mc = use_mc && ((istep % (unsigned long long)P.nstep_bp_MC == 0ull) || after_rebuild);
        after_rebuild = false;
        const bool gate_mask = use_mc && !mc;
        const RepParams &Rp = *sm.rp;
        const double ele_coef = Rp.ele_coef, inv_lambdaD = Rp.inv_lambdaD;
        const double cut2 = P.flg_ele ? Rp.ele_cutoff * Rp.ele_cutoff : -1.0;
        const double noise_T = Rp.sqrt_tempK, twz_force = Rp.twz_force;
        if (prof) c1 = clock64();
        // ---- P1: every warp draws chunks from the shared counter ----------------------------------------------
        int nact = -1, nar = 0, cang = 0, cdih = 0, cov = 0, cene = 0, ccomb = 0, cbp = 0;
        const int cfix = cgate + cstat + cG;        // chunks whose meaning does not depend on the gate
        {
            bool first = true;
            for (;;) {
                int ch = 0;
                if (lane == 0) ch = atomicAdd(&sm.ctl[RC_WORK], 1);
                ch = __shfl_sync(0xffffffffu, ch, 0);
                if (prof && first) { c2 = clock64(); first = false; }
                const long long tq0 = tprof ? clock64() : 0;
                int ttype = 0;
                if (ch < cgate) {                       // distance gate of 32 candidates
                    ttype = 8;
                    rep_bp_gate_warp(P, sm, 32 * ch + lane, ncand, gate_mask);
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) atomicAdd(&sm.ctl[RC_GATE], 1);
                } else if (ch < cgate + cstat) {
                    const int cs = ch - cgate;
                    if (cs < cD) {                      // dihedral starting at bead k
                        ttype = 4;
                        const int k = 32 * cs + lane;
                        if (k < n) {
                            const int4 m = sm.meta[k];
                            if (rep_term(m, 3)) {
                                Vec3 f4[4];
                                dih_force_vecs(P, rep_x(sm.x, k), rep_x(sm.x, chain_nb(m, 1)), rep_x(sm.x, chain_nb(m, 2)),
                                               rep_x(sm.x, chain_nb(m, 3)), f4);
#pragma unroll
                                for (int p = 0; p < 4; p++) rep_put(sm.bf, 4 * n + 4 * k + p, f4[p]);
                            }
                        }
                    } else if (cs < cD + cB) {          // bond and angle starting at bead k
                        ttype = 5;
                        const int k = 32 * (cs - cD) + lane;
                        if (k < n) {
                            const int4 m = sm.meta[k];
                            if (rep_term(m, 1)) {
                                const Vec3 x0 = rep_x(sm.x, k), x1 = rep_x(sm.x, chain_nb(m, 1));
                                rep_put(sm.bf, k, bond_force_vec(P, x0, x1));
                                if (rep_term(m, 2)) {
                                    Vec3 fa, fc;
                                    if (P.flg_angl_ReB) angl_ReB_force_vecs(P, x0, x1, rep_x(sm.x, chain_nb(m, 2)), &fa, &fc);
                                    else angl_force_vecs(P, x0, x1, rep_x(sm.x, chain_nb(m, 2)), &fa, &fc);
                                    rep_put(sm.bf, n + 3 * k, fa); rep_put(sm.bf, n + 3 * k + 1, -(fa + fc)); rep_put(sm.bf, n + 3 * k + 2, fc);
                                }
                            }
                        }
                    } else {                            // 32 * PPL pairs: PPL per lane side by side
                        ttype = 6;
                        const int q0 = 32 * PPL * (cs - cD - cB) + lane;
                        uint32_t e[PPL];
                        Vec3 f[PPL];
#pragma unroll
                        for (int u = 0; u < PPL; u++) e[u] = q0 + 32 * u < npair ? sm.pl[q0 + 32 * u] : 0u;
#pragma unroll
                        for (int u = 0; u < PPL; u++) f[u] = e[u] ? rep_pair_force(P, e[u], sm.x, ele_coef, inv_lambdaD, cut2) : mk(0.0, 0.0, 0.0);
#pragma unroll
                        for (int u = 0; u < PPL; u++)
                            if (e[u]) {
                                const int i = e[u] & 0xFF, j = (e[u] >> 8) & 0xFF;
                                double *ri = sm.Fp + 3 * i * npad + j, *rj = sm.Fp + 3 * j * npad + i;
                                ri[0] = f[u].x; ri[npad] = f[u].y; ri[2 * npad] = f[u].z;
                                rj[0] = -f[u].x; rj[npad] = -f[u].y; rj[2 * npad] = -f[u].z;
                            }
                    }
                    __threadfence_block();
                    __syncwarp();
                    if (lane == 0) atomicAdd(&sm.ctl[RC_STAT], 1);
                } else if (ch < cfix) {                 // Gaussians of the NEXT step (nothing waits for them in this one): draw 0 -> (x, y), draw 1 -> z
                    ttype = 3;
                    const int it = 32 * (ch - cgate - cstat) + lane;
                    if (it < 2 * n) {
                        const int k = it >> 1, draw = it & 1;
                        double g0, g1;
                        gauss_pair(A.seed, stream, istep + 1ull, (uint32_t)k, (uint32_t)draw, &g0, &g1);
                        double *gn = sm.g + 3 * n * (int)((istep + 1ull) & 1ull);
                        if (draw == 0) { gn[3 * k] = g0; gn[3 * k + 1] = g1; } else gn[3 * k + 2] = g0;
                    }
                } else {
                    if (nact < 0) {                     // the chunks behind depend on the survivors: the gate must be complete
                        rep_wait(&sm.ctl[RC_GATE], cgate, &S.misc->max_force_flag);
                        nact = sm.ctl[RC_NACT];
                        nar = min(nact, REP_ACT);
                        // angle roles, dihedral roles, survivors beyond the role slots, energies + per-bead forces, owners' sums
                        cang = (4 * nar + 31) >> 5; cdih = (3 * nar + 31) >> 5; cov = (nact - nar + 31) >> 5;
                        cene = (nar + 31) >> 5; ccomb = cene + (((NT >= 512 ? 6 : 1) * nar + 31) >> 5);
                        cbp = cang + cdih + cov;
                    }
                    ch -= cfix;
                    if (ch >= cbp + ccomb + cA) break;
                    if (ch < cang + cdih) {             // base-pair roles, one code path per warp
                        ttype = ch < cang ? 0 : 1;
                        const bool ang = ch < cang;
                        const int it = 32 * (ang ? ch : ch - cang) + lane;
                        if (ang) { if (it < 4 * nar) rep_bp_role_item(P, sm, it >> 2, 1 + (it & 3)); }
                        else if (it < 3 * nar) { const int a = it / 3, q = it - 3 * a; rep_bp_role_item(P, sm, a, q == 2 ? 0 : 5 + q); }
                        __threadfence_block();
                        __syncwarp();
                        if (lane == 0) atomicAdd(&sm.ctl[RC_ROLES], 1);
                    } else if (ch < cbp) {
                        ttype = 2;
                        const int a = REP_ACT + 32 * (ch - cang - cdih) + lane;
                        if (a < nact) rep_bp_eval(P, sm, sm.act[a], !mc);
                    } else if (ch < cbp + ccomb) {      // energies, per-bead forces of the survivors (they need the roles)
                        ttype = 2;
                        rep_wait(&sm.ctl[RC_ROLES], cang + cdih, &S.misc->max_force_flag);
                        const int cc = ch - cbp;
                        if (cc < cene) { const int a = 32 * cc + lane; if (a < nar) rep_bp_energy(P, sm, a, mc); }
                        else if (NT >= 512) { const int it = 32 * (cc - cene) + lane; if (it < 6 * nar) rep_bp_gather(P, sm, it / 6, it % 6, RC_SPILL0 + (done & 1)); }
                        else {
                            const int a = 32 * (cc - cene) + lane;
                            if (a < nar)
                                for (int t = 0; t < 6; t++) rep_bp_gather(P, sm, a, t, RC_SPILL0 + (done & 1));
                        }
                    } else {                            // owners' sums: pair row, bonded terms, tweezers (last: they need the pairs)
                        ttype = 7;
                        rep_wait(&sm.ctl[RC_STAT], cstat, &S.misc->max_force_flag);
                        const int it = 32 * (ch - cbp - ccomb) + lane;
                        if (it < 3 * n) {
                            const int k = it / 3, comp = it - 3 * k;
                            double f0 = 0.0, f1 = 0.0, f2 = 0.0, f3 = 0.0;
                            const double2 *row = reinterpret_cast<const double2 *>(sm.Fp + (size_t)it * npad);
                            int j = 0;
                            for (; j + 1 < npad / 2; j += 2) { const double2 p0 = row[j], p1 = row[j + 1]; f0 += p0.x; f1 += p0.y; f2 += p1.x; f3 += p1.y; }
                            if (j < npad / 2) { const double2 p0 = row[j]; f0 += p0.x; f1 += p0.y; }
                            double f = (f0 + f1) + (f2 + f3);
                            const unsigned short *be = sm.bent + REP_BENT * k;
                            const int ne = sm.bn[k];
                            for (int e = 0; e < ne; e++) {
                                const unsigned w = be[e];
                                const double t = sm.bf[3 * (w & 0x7FFF) + comp];
                                f += (w & 0x8000) ? -t : t;
                            }
                            // dual constant-force tweezers (src/force_tweezers.F90:27-35): +f on the first bead of a pair, -f on the second
                            for (int t = 0; t < P.ntwz_DCF; t++) {
                                const double ft = twz_force * S.twz_dir[3 * t + comp];
                                if (S.twz_pairs[2 * t] == k) f += ft;
                                if (S.twz_pairs[2 * t + 1] == k) f -= ft;
                            }
                            sm.fsum[it] = f;
                        }
                    }
                }
                if (tprof) {
                    __syncwarp();
                    if (lane == 0) { atomicAdd(S.prof + 16 + 8 * 1000 + 2 * ttype, (unsigned long long)(clock64() - tq0)); atomicAdd(S.prof + 16 + 8 * 1000 + 2 * ttype + 1, 1ull); }
                }
            }
        }
        if (wprof) w0 = clock64();
        rep_sync<NT>();
        if (mc) rep_mc_prune<NT>(P, sm, n, ncand, A.seed, A.max_bp_per_nt, &s_mc_ctr, S.prof ? pc + 6 : nullptr);
        if (prof) c3 = clock64();
        // ---- P2: assemble and integrate (src/job_md.F90:409-456) ------------------------------------------------
        bool far = false;
        const bool spill = sm.ctl[RC_SPILL0 + (done & 1)] != 0;
        if (wprof) { w1 = clock64(); aw[0] += w1 - w0; }
        for (int k = tid >> 2; k < ((n + NT / 4 - 1) / (NT / 4)) * (NT / 4); k += NT / 4) {
            const int comp = tid & 3;
            double mv = 0.0;
            if (k < n && comp < 3) {
                double f = sm.fsum[3 * k + comp];
                const int nb = sm.bcnt[k];
                if (!spill) {
                    for (int s = 0; s < nb; s++) { const int w = sm.blst[REP_BPL * k + s]; f += sm.eA[w / 6] * sm.bpF[3 * w + comp]; }
                } else if (nb) {
                    for (int a = 0; a < nar; a++) {
                        const unsigned char *b = sm.c_b + 8 * sm.act[a];
#pragma unroll
                        for (int t = 0; t < 6; t++)
                            if (b[t] == k) f += sm.eA[a] * sm.bpF[3 * (6 * a + t) + comp];
                    }
                }
                const double fo = sm.fbp[3 * k + comp];      // base pairs beyond the slot tables
                if (fo != 0.0) { f += fo; sm.fbp[3 * k + comp] = 0.0; }
                const double gz = sm.g[3 * n * (int)(istep & 1ull) + 3 * k + comp], noise = sm.nz0[k] * noise_T;
                double dx;
                if (P.integrator == 1) {
                    const double accels_pre = noise * gz;
                    const double vn = sm.ca[k] * sm.v[3 * k + comp] + sm.cb[k] * f + (sm.a[3 * k + comp] + accels_pre);
                    sm.v[3 * k + comp] = vn;
                    sm.a[3 * k + comp] = accels_pre;
                    dx = sm.cc[k] * vn;
                } else {
                    dx = sm.ca[k] * f + noise * gz;
                }
                sm.x[3 * k + comp] += dx;
                mv = sm.m[3 * k + comp] + dx;
                sm.m[3 * k + comp] = mv;
                if (f > 100.0) atomicOr(&S.misc->max_force_flag, 1);
            }
            const double s0 = __shfl_sync(0xffffffffu, mv, 0, 4), s1 = __shfl_sync(0xffffffffu, mv, 1, 4), s2 = __shfl_sync(0xffffffffu, mv, 2, 4);
            const double d2 = s0 * s0 + s1 * s1 + s2 * s2;
            if (k < n && comp == 0) { sm.d2[k] = d2; far |= d2 > d2_any; }
            __syncwarp();
            if (k < n && comp == 3) sm.bcnt[k] = 0;
        }
        if (tid == 0) { sm.ctl[RC_NACT] = 0; sm.ctl[RC_WORK] = 0; sm.ctl[RC_ROLES] = 0; sm.ctl[RC_STAT] = 0; sm.ctl[RC_GATE] = 0; sm.ctl[RC_SPILL0 + ((done + 1) & 1)] = 0; sm.ctl[RC_FAR0 + ((done + 1) & 1)] = 0; }
        // "is any displacement beyond half the margin": only then the exact top-2 test (flag by step parity)
        if (__any_sync(0xffffffffu, far) && lane == 0) sm.ctl[RC_FAR0 + (done & 1)] = 1;
        if (wprof) { w0 = clock64(); aw[1] += w0 - w1; }
        rep_sync<NT>();
        if (wprof) { w1 = clock64(); aw[2] += w1 - w0; }
        const int any_far = sm.ctl[RC_FAR0 + (done & 1)];
        if (prof) c4 = clock64();
        blocked = false;
        if (any_far) {
            double t1 = sm.d2[lane], t2 = sm.d2[lane + 32];
            if (t2 > t1) { const double t = t1; t1 = t2; t2 = t; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double b1 = __shfl_xor_sync(0xffffffffu, t1, o), b2 = __shfl_xor_sync(0xffffffffu, t2, o);
                top2_merge(t1, t2, b1, b2);
            }
            blocked = sqrt(t1) + sqrt(t2) > S.nl_margin;
        }
        if (prof) c5 = clock64();
        if (XCH) {
            if (xch && istep % (unsigned long long)S.remd->nex == 0ull) rep_remd_exchange<NT>(S, P, sm, rep, n, npair, ncand);
        }
        if (prof) {
            pc[0] += (unsigned long long)(c1 - c0); pc[1] += (unsigned long long)(c2 - c1); pc[2] += (unsigned long long)(c3 - c2);
            pc[3] += (unsigned long long)(c4 - c3); pc[4] += (unsigned long long)(c5 - c4); pc[5] += (unsigned long long)(clock64() - c5);
            pc[7] += 1ull;
        }
    }

    if (wprof) for (int i = 0; i < 3; i++) S.prof[3 * warp + i] += (unsigned long long)aw[i];
    // ---- epilogue: state, lists and base-pair records back to global memory ---------------------------------
    const unsigned long long stepf = ctl.step + (unsigned long long)done;
    const int curf = ctl.cur ^ (done & 1);
    if (done > 0) {       // exact displacement top-2 of the last step for the next launch
        double t1 = sm.d2[lane], t2 = sm.d2[lane + 32];
        if (t2 > t1) { const double t = t1; t1 = t2; t2 = t; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double b1 = __shfl_xor_sync(0xffffffffu, t1, o), b2 = __shfl_xor_sync(0xffffffffu, t2, o);
            top2_merge(t1, t2, b1, b2);
        }
        top1 = t1; top2v = t2;
    }
    if (tid == 0) S.rp[rep] = *sm.rp;
    if (nreb > 0) {
        // the lists of the last rebuild event (src/neighbor_list.F90:20-324), from the positions saved there
        double4 *scratch = S.pos[curf ^ 1] + rbase;
        for (int k = tid; k < n; k += NT) scratch[k] = make_double4(sm.xreb[3 * k], sm.xreb[3 * k + 1], sm.xreb[3 * k + 2], 0.0);
        if (tid == 0) { S.nbp_half[rep] = 0; atomicAdd(&S.misc->nrebuild, nreb); }
        __threadfence_block();
        rep_sync<NT>();
        NlArgs NB = NA;
        NB.all_pairs = 2;
        for (int i = warp; i < n; i += NT / 32) nl_rows_bead_at(S, P, NB, rep, i, lane, scratch, nullptr, nullptr);
        __threadfence_block();
        rep_sync<NT>();
        for (int idx = tid; idx < n * KB; idx += NT) nl_bp_slot(S, rep, idx);
        __threadfence_block();
        rep_sync<NT>();
        for (int c = tid; c < ncand; c += NT) {
            const unsigned char *b = sm.c_b + 8 * c;
            const int nrow = S.nbp_row[rbase + b[0]];
            int pb = -1;
            for (int e = 0; e < nrow; e++) {
                const int2 ent = S.bprow[(rbase + b[0]) * KB + e];
                if (ent.x == b[1]) { pb = ent.y; break; }
            }
            sm.c_pb[c] = pb;
        }
    }
    for (int c = tid; c < ncand; c += NT) {
        const int pb = sm.c_pb[c];
        if (pb < 0) continue;
        const unsigned char st = sm.c_b[8 * c + 7];
        S.bp_status[pbase + pb] = st & 1;
        S.bp_status_MC[pbase + pb] = (st >> 1) & 1;
        S.ene_bp[pbase + pb] = (st & 1) ? sm.c_ene[c] : 0.0;
    }
    for (int k = tid; k < n; k += NT) S.pos[curf][rbase + k] = make_double4(sm.x[3 * k], sm.x[3 * k + 1], sm.x[3 * k + 2], 0.0);
    for (int t = tid; t < 3 * n; t += NT) { S.vel[rbase * 3 + t] = sm.v[t]; S.acc[rbase * 3 + t] = sm.a[t]; S.move[rbase * 3 + t] = sm.m[t]; }
    if (tid == 0) {
        if (use_mc) S.mc_counter[rep] = s_mc_ctr;
        if (done > 0) {       // this replica's displacement top-2 (src/job_md.F90:376-388)
            unsigned long long *t = S.top2 + ((size_t)rep * 2 + (stepf & 1ull)) * 2;
            t[0] = (unsigned long long)__double_as_longlong(top1); t[1] = (unsigned long long)__double_as_longlong(top2v);
            t = S.top2 + ((size_t)rep * 2 + ((stepf + 1ull) & 1ull)) * 2;
            t[0] = 0ull; t[1] = 0ull;
        }
        if (xch && blockIdx.x == 0) {      // tables and counters of the exchange back for the host / the next launch
            RemdDev &D = *S.remd;
            const FastRemdState X = *sm.xs;
            for (int i = 0; i < D.n; i++) { D.rep2lab[i] = sm.r2l[i]; D.lab2rep[i] = sm.l2r[i]; }
            D.idx_pair = X.idx_pair; D.exchange_counter = X.counter;
            D.n_attempt += X.natt; D.n_accept += X.nacc;
            D.xch_round0 += (unsigned long long)X.attempt;
        }
        // the last CTA to finish publishes the new step (S.bar[0] is zeroed by the host before every launch)
        __threadfence();
        if (atomicAdd(S.bar, 1u) == gridDim.x - 1) {
            DevCtl a; a.step = stepf; a.cur = curf; a.blocked = 0;
            S.ctl[0] = a;
        }
    }
}
#endif
