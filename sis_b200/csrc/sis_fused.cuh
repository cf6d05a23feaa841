// sis_fused.cuh -- the MD loop body (src/job_md.F90:374-460) as ONE persistent, warp-specialised
// cooperative kernel ("k_md_fused").
//
// One CTA per SM, 20 warps:
//   * 16 MAIN warps = 64 groups of 8 lanes.  A group owns bead k of replica r for the whole
//     launch: it streams k's neighbour row (WCA + Debye-Hueckel, src/force_wca.F90,
//     src/force_ele_DH.F90), evaluates the bonded terms and per-bead biases k takes part in
//     (src/force_bond.F90 ... force_restraint.F90) in gather form, keeps the force in registers,
//     and integrates k (GJF-2GJ / Ermak-McCammon, src/job_md.F90:409-456) as soon as the base-pair
//     forces of the step are complete.
//   * 4 BP warps = 16 groups of 8 lanes.  A group evaluates one base-pair candidate
//     cooperatively (one role per lane: bond, 4 angles, 2 dihedrals; src/force_bp.F90:50-204,
//     src/force_bp_triplet.F90:62-240) and scatters the six bead forces with fp64 atomics.
// Two grid-wide barriers per step, both split into arrive / wait:
//   bar[0] "bp done"  : arrived by the BP warps, awaited by the MAIN warps only after their own
//                       force work -- it is normally complete by then, so it costs nothing;
//   bar[1] "positions": arrived after the position update; the only exposed barrier of a step.
// The kernel leaves as soon as a replica passes the neighbour-list margin (src/job_md.F90:388)
// so that the host can rebuild the lists and relaunch.
#pragma once

#include "sis_kernels.cuh"

constexpr int FUSED_MAIN_WARPS = 16, FUSED_BP_WARPS = 4;
constexpr int FUSED_THREADS = 32 * (FUSED_MAIN_WARPS + FUSED_BP_WARPS);
constexpr int FUSED_G = 8;                                    // lanes per bead / per bp candidate
constexpr int FUSED_GROUPS = 32 * FUSED_MAIN_WARPS / FUSED_G;   // bead groups per CTA
constexpr int FUSED_BPGROUPS = 32 * FUSED_BP_WARPS / FUSED_G;   // bp groups per CTA

SIS_D unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
SIS_D void red_release_add(unsigned *p, unsigned v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
SIS_D unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SIS_D void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// sum over the 8 lanes named by `mask` (the groups of one warp may have diverged)
SIS_D double group_sum_masked(double v, unsigned mask)
{
#pragma unroll
    for (int o = FUSED_G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, FUSED_G);
    return v;
}

// one base-pair candidate on the 8 lanes of a group (role = lane)
SIS_D void fused_bp_candidate(const DevSys &S, const FFParams &P, const RepParams &R, const double4 *__restrict__ pos,
                              int rep, int ibp, int lane)
{
    const int nmp = S.nmp, KB = S.KB;
    const size_t rbase = (size_t)rep * nmp, pbase = rbase * KB;
    const int k = ibp / KB, e = ibp - k * KB;
    const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~(FUSED_G - 1));
    const int j = S.bprow[(rbase + k) * KB + e].x;
    const int type = S.bp_type[pbase + ibp];
    double dG = 1.0;
    bool active = true;
    if (P.bp_triplet) {
        const double2 cf = S.bp_coef[pbase + ibp];
        dG = cf.x - R.tempK * cf.y;     // src/force_bp_triplet.F90:77
        if (dG >= 0.0) active = false;
    }
    int st = 0;
    double en = 0.0;
    if (active) {                       // uniform across the group
        const BpPara &p = P.bp[type];
        const Vec3 x0 = v3(pos[k]), x1 = v3(pos[j]);
        const Vec3 v12 = pbc_vec_d(P, x0, x1);
        const double d1212 = dot(v12, v12), a12 = sqrt(d1212);
        if (fabs(a12 - p.bond_r) <= p.cutoff_ddist) {
            const int4 mi = S.meta[k], mj = S.meta[j];
            // beads 0..5 = i, j, i-1, j-1, i+1, j+1
            const int tA = bp_role_A(lane), tB = bp_role_B(lane);
            auto bead_of = [&](int t) -> int {
                return t == 0 ? k : t == 1 ? j : t == 2 ? chain_nb(mi, -1) : t == 3 ? chain_nb(mj, -1)
                     : t == 4 ? chain_nb(mi, +1) : chain_nb(mj, +1);
            };
            Vec3 xa = x0, xb = x0;
            if (tA >= 0) xa = v3(pos[bead_of(tA)]);
            if (tB >= 0) xb = v3(pos[bead_of(tB)]);
            const BpSub o = bp_force_role(P, p, lane, x0, x1, xa, xb, v12, d1212, a12);
            const double u = group_sum_masked(o.u, gmask);
            const double ene = P.bp_triplet ? p.U0 * dG * exp(-u) : p.U0 * exp(-u);
            // forces on the six beads: beads 0, 1 collect every role; 2..5 two roles each
            Vec3 F0, F1;
            F0.x = group_sum_masked(o.c0.x, gmask); F0.y = group_sum_masked(o.c0.y, gmask); F0.z = group_sum_masked(o.c0.z, gmask);
            F1.x = group_sum_masked(o.c1.x, gmask); F1.y = group_sum_masked(o.c1.y, gmask); F1.z = group_sum_masked(o.c1.z, gmask);
            // bead t (= lane, 2..5): ca of the angle role + (ca | cb) of the dihedral role
            //   bead 2: ca(role 1) + cb(role 5)     bead 3: ca(role 2) + ca(role 5)
            //   bead 4: ca(role 3) + cb(role 6)     bead 5: ca(role 4) + ca(role 6)
            const int t = lane;
            const int src_ang = (t >= 2 && t <= 5) ? t - 1 : 0;
            const int src_dih = (t == 2 || t == 3) ? 5 : 6;
            const bool dih_b = (t == 2 || t == 4);
            Vec3 Fa, Fd_a, Fd_b;
            Fa.x = __shfl_sync(gmask, o.ca.x, src_ang, FUSED_G);
            Fa.y = __shfl_sync(gmask, o.ca.y, src_ang, FUSED_G);
            Fa.z = __shfl_sync(gmask, o.ca.z, src_ang, FUSED_G);
            Fd_a.x = __shfl_sync(gmask, o.ca.x, src_dih, FUSED_G);
            Fd_a.y = __shfl_sync(gmask, o.ca.y, src_dih, FUSED_G);
            Fd_a.z = __shfl_sync(gmask, o.ca.z, src_dih, FUSED_G);
            Fd_b.x = __shfl_sync(gmask, o.cb.x, src_dih, FUSED_G);
            Fd_b.y = __shfl_sync(gmask, o.cb.y, src_dih, FUSED_G);
            Fd_b.z = __shfl_sync(gmask, o.cb.z, src_dih, FUSED_G);
            if (ene <= P.bp_cutoff_energy) {
                st = 1; en = ene;
                if (t < 6) {
                    Vec3 F = t == 0 ? F0 : t == 1 ? F1 : Fa + (dih_b ? Fd_b : Fd_a);
                    double *fb = S.f_bp + (rbase + bead_of(t)) * 3;
                    atomicAdd(fb, ene * F.x);
                    atomicAdd(fb + 1, ene * F.y);
                    atomicAdd(fb + 2, ene * F.z);
                }
            }
        } else if (P.bp_triplet && P.flg_bias_ss) {   // src/force_bp_triplet.F90:85-93
            if (lane < 2) {
                const double s = (lane == 0 ? -1.0 : 1.0) * (P.bias_ss_force / a12);
                double *fb = S.f_bp + (rbase + (lane == 0 ? k : j)) * 3;
                atomicAdd(fb, s * v12.x);
                atomicAdd(fb + 1, s * v12.y);
                atomicAdd(fb + 2, s * v12.z);
            }
        }
    }
    if (lane == 0) {
        S.bp_status[pbase + ibp] = st;
        S.ene_bp[pbase + ibp] = en;
    }
}

// integrate one bead (lanes 0..2 = x, y, z of the owning group); returns |xyz_move|^2 on every
// lane of the group.  f = total force component of this lane.
SIS_D double fused_integrate(const DevSys &S, const FFParams &P, const StepArgs &A, const RepParams &R, int cur, int rep,
                             int k, int lane, double f, double g, bool valid)
{
    const int nmp = S.nmp;
    double xnew = 0.0, mv = 0.0;
    if (valid && lane < 3) {
        const size_t o = ((size_t)rep * nmp + k) * 3 + lane;
        const double4 xo = S.pos[cur][(size_t)rep * nmp + k];
        const double4 c = S.coef[k];
        const double x = lane == 0 ? xo.x : lane == 1 ? xo.y : xo.z;
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
    const double nx = __shfl_sync(0xffffffffu, xnew, 0, FUSED_G);
    const double ny = __shfl_sync(0xffffffffu, xnew, 1, FUSED_G);
    const double nz = __shfl_sync(0xffffffffu, xnew, 2, FUSED_G);
    const double s0 = __shfl_sync(0xffffffffu, mv, 0, FUSED_G);
    const double s1 = __shfl_sync(0xffffffffu, mv, 1, FUSED_G);
    const double s2 = __shfl_sync(0xffffffffu, mv, 2, FUSED_G);
    if (valid && lane == 0) S.pos[cur ^ 1][(size_t)rep * nmp + k] = make_double4(nx, ny, nz, 0.0);
    (void)A;
    return valid ? s0 * s0 + s1 * s1 + s2 * s2 : 0.0;
}

SIS_D void top2_push(unsigned long long *t, double t1, double t2)
{
    // two order-preserving integer maxima leave (m1, m2) = the two largest values whatever the
    // arrival order:  old = atomicMax(m1, a1);  atomicMax(m2, max(min(old, a1), a2))
    const unsigned long long old = atomicMax(t, (unsigned long long)__double_as_longlong(t1));
    const double oldd = __longlong_as_double((long long)old);
    const double c2 = fmax(fmin(oldd, t1), t2);
    if (c2 > 0.0) atomicMax(t + 1, (unsigned long long)__double_as_longlong(c2));
}

// number of steps a persistent launch runs: up to step `target`, none on invalid lists
SIS_D int md_steps_to_run(const DevSys &S, const DevCtl &ctl, long long target)
{
    if (S.misc->lists_invalid) return 0;
    const long long left = target - (long long)ctl.step;
    return left <= 0 ? 0 : (int)(left < (1ll << 20) ? left : (1ll << 20));
}

__global__ void __launch_bounds__(FUSED_THREADS, 1) k_md_fused(DevSys S, FFParams P, StepArgs A, long long target)
{
    __shared__ int s_blocked;
    __shared__ double s_top[FUSED_MAIN_WARPS][2];
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool is_main = warp < FUSED_MAIN_WARPS;
    const int R = S.nrep, nmp = S.nmp;
    const int nitems = R * nmp;
    const int ncta = gridDim.x;
    const int niter = (nitems + ncta * FUSED_GROUPS - 1) / (ncta * FUSED_GROUPS);
    const DevCtl ctl = S.ctl[0];
    const int nsteps = md_steps_to_run(S, ctl, target);
    if (nsteps <= 0) return;
    unsigned long long step = ctl.step;
    int cur = ctl.cur;
    unsigned *bar_bp = S.bar, *bar_pos = S.bar + 32;   // separate 128-byte lines
    const bool prof = S.prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;

    int n = 0;
    for (; n < nsteps; n++) {
        // ---- rebuild test of src/job_md.F90:388 on the displacements left by step `step` -------
        if (warp == 0) {
            __syncwarp();   // lanes 1..31 wait here for thread 0 (still inside the grid barrier)
            bool b = false;
            for (int r = tid; r < R; r += 32) {
                const unsigned long long *t = S.top2 + ((size_t)r * 2 + (step & 1ull)) * 2;
                const double m1 = __longlong_as_double((long long)__ldcg(t));
                const double m2 = __longlong_as_double((long long)__ldcg(t + 1));
                b |= (sqrt(m1) + sqrt(m2) > S.nl_margin);
            }
            b = __any_sync(0xffffffffu, b);
            if (tid == 0) s_blocked = b;
        }
        __syncthreads();
        if (s_blocked) break;
        if (prof) c0 = clock64();

        if (!is_main) {
            // ================= BP warps ==========================================================
            const int lane = tid & (FUSED_G - 1);
            const int bg = (tid - 32 * FUSED_MAIN_WARPS) / FUSED_G;
            if (A.term_mask & (1u << ENE_BP)) {
                for (int rep = 0; rep < R; rep++) {
                    const int nb = S.nbp_half[rep];
                    const double4 *pos = S.pos[cur] + (size_t)rep * nmp;
                    const int *half = S.bp_half + (size_t)rep * S.bp_half_cap;
                    for (int t = blockIdx.x * FUSED_BPGROUPS + bg; t < nb; t += ncta * FUSED_BPGROUPS)
                        fused_bp_candidate(S, P, S.rp[rep], pos, rep, half[t], lane);
                }
            }
            named_bar_sync(1, 32 * FUSED_BP_WARPS);
            if (tid == 32 * FUSED_MAIN_WARPS) { __threadfence(); red_release_add(bar_bp, 1u); }
        } else {
            // ================= MAIN warps ========================================================
            const int lane = tid & (FUSED_G - 1);
            const int grp = tid / FUSED_G;
            Vec3 Freg = mk(0.0, 0.0, 0.0);
            double greg = 0.0;
            for (int it = 0; it < niter; it++) {
                const int item = (it * ncta + blockIdx.x) * FUSED_GROUPS + grp;
                const bool valid = item < nitems;
                const int rep = valid ? item / nmp : 0;
                const int k = valid ? item - rep * nmp : 0;
                const RepParams &Rp = S.rp[rep];
                const double4 *pos = S.pos[cur] + (size_t)rep * nmp;
                Vec3 F = pair_force<FUSED_G>(S, P, Rp, pos, rep, k, lane, A.term_mask);
                F += local_force<FUSED_G>(S, P, Rp, pos, rep, k, lane, A.term_mask, step + 1ull);
                F.x = group_sum<FUSED_G>(F.x); F.y = group_sum<FUSED_G>(F.y); F.z = group_sum<FUSED_G>(F.z);
                double g = 0.0;
                // with one bead per group the Gaussian is drawn here, off the critical path of the
                // integrate phase; otherwise in the integrate loop (once per bead either way)
                if (niter == 1 && lane < 3) g = gauss_component(A.seed, (uint32_t)Rp.stream, step + 1ull, (uint32_t)k, lane);
                if (niter == 1) { Freg = F; greg = g; }
                else if (valid && lane < 3)
                    S.f_pair[((size_t)rep * nmp + k) * 3 + lane] = lane == 0 ? F.x : lane == 1 ? F.y : F.z;
            }
            if (prof) c1 = clock64();
            // ---- wait for the base-pair scatter of every CTA -------------------------------------
            if (tid == 0) {
                const unsigned target = (unsigned)(n + 1) * (unsigned)ncta;
                while (ld_acquire_u32(bar_bp) < target) { }
            }
            named_bar_sync(2, 32 * FUSED_MAIN_WARPS);
            if (prof) c2 = clock64();
            // ---- integrate ------------------------------------------------------------------------
            double t1 = 0.0, t2 = 0.0;   // warp top-2 of |xyz_move|^2 (R == 1) / per-group push (R > 1)
            for (int it = 0; it < niter; it++) {
                const int item = (it * ncta + blockIdx.x) * FUSED_GROUPS + grp;
                const bool valid = item < nitems;
                const int rep = valid ? item / nmp : 0;
                const int k = valid ? item - rep * nmp : 0;
                const RepParams &Rp = S.rp[rep];
                double f = 0.0, g = greg;
                if (valid && lane < 3) {
                    const size_t o = ((size_t)rep * nmp + k) * 3 + lane;
                    if (niter == 1) f = lane == 0 ? Freg.x : lane == 1 ? Freg.y : Freg.z;
                    else {
                        f = S.f_pair[o];
                        g = gauss_component(A.seed, (uint32_t)Rp.stream, step + 1ull, (uint32_t)k, lane);
                    }
                    f += __ldcg(S.f_bp + o);
                    S.f_bp[o] = 0.0;
                }
                const double d2 = fused_integrate(S, P, A, Rp, cur, rep, k, lane, f, g, valid);
                if (R == 1) {
                    if (lane == 0) top2_merge(t1, t2, d2, 0.0);
                } else if (valid && lane == 0 && d2 > 0.0) {
                    top2_push(S.top2 + ((size_t)rep * 2 + ((step + 1ull) & 1ull)) * 2, d2, 0.0);
                }
            }
            if (R == 1) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double b1 = __shfl_xor_sync(0xffffffffu, t1, o), b2 = __shfl_xor_sync(0xffffffffu, t2, o);
                    top2_merge(t1, t2, b1, b2);
                }
                if ((tid & 31) == 0) { s_top[warp][0] = t1; s_top[warp][1] = t2; }
            }
            named_bar_sync(2, 32 * FUSED_MAIN_WARPS);
            if (tid == 0) {
                if (R == 1) {
                    double a1 = 0.0, a2 = 0.0;
                    for (int w = 0; w < FUSED_MAIN_WARPS; w++) top2_merge(a1, a2, s_top[w][0], s_top[w][1]);
                    if (a1 > 0.0) top2_push(S.top2 + ((step + 1ull) & 1ull) * 2, a1, a2);
                }
                if (blockIdx.x == 0)   // clear the slots step+2 will accumulate into (nobody touches them before)
                    for (int r = 0; r < R; r++) {
                        unsigned long long *t = S.top2 + ((size_t)r * 2 + ((step + 2ull) & 1ull)) * 2;
                        t[0] = 0ull; t[1] = 0ull;
                    }
                if (prof) c3 = clock64();
                __threadfence();
                red_release_add(bar_pos, 1u);
                const unsigned target = (unsigned)(n + 1) * (unsigned)ncta;
                while (ld_acquire_u32(bar_pos) < target) { }
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
        // the next iteration starts with warp 0 reading top2 and a __syncthreads that releases
        // every warp (BP warps included) only after thread 0 has seen all positions written
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) {
        DevCtl a; a.step = step; a.cur = cur; a.blocked = 0;
        S.ctl[0] = a;
    }
}
