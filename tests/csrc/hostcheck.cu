// hostcheck.cu -- TEST-ONLY host build of the product's force and list code.
//
// There is no GPU in the build container, so the kernels cannot run there.  This file compiles
// the product sources (sis_b200/csrc/sisgpu.cu and the __host__ __device__ math / gather code in
// sis_kernels.cuh) for the HOST and drives the force task code (`pair_force<1>`, `local_force<1>`, `bp_pair_force`) and the list predicates serially, so
// the CPU test-suite can compare the product's arithmetic, term roles and exclusion rules with
// the oracle before any GPU time is spent.  It is not linked into libsisgpu.so and nothing in
// sis_b200/ loads it.
#include "../../sis_b200/csrc/sisgpu.cu"

struct HostCheck {
    sisgpu *h = nullptr;
    std::vector<int4> meta;
    std::vector<double4> coef, pos;
    std::vector<uint32_t> nbr;
    std::vector<int> nnbr, nbp_row, bp_status, bp_status_MC;
    std::vector<int2> bprow;
    std::vector<signed char> bp_type;
    std::vector<double2> bp_coef;
    std::vector<double> ene_bp;
    RepParams rp{};
    DevSys S{};
    int K = 0, KB = 0;
};

extern "C" {

int hc_create(HostCheck **out)
{
    HostCheck *c = new HostCheck();
    c->h = new sisgpu();
    c->h->P.integrator = 1;
    c->h->P.bp_cutoff_energy = 0.001;
    *out = c;
    return 0;
}
sisgpu *hc_handle(HostCheck *c) { return c->h; }
void hc_destroy(HostCheck *c) { delete c->h; delete c; }

int hc_prepare(HostCheck *c)
{
    sisgpu *h = c->h;
    TRY(host_prepare(h, c->meta, c->coef));
    fill_rp(h, 0, c->rp);
    DevSys &S = c->S;
    S.nmp = h->nmp; S.nrep = 1;
    S.meta = c->meta.data(); S.seq = h->seq.data(); S.coef = c->coef.data(); S.mass = h->mass.data();
    S.hash = h->bp3_hash.data(); S.bp3_dH = h->bp3_dH.data(); S.bp3_dS = h->bp3_dS.data();
    S.twz_pairs = h->twz_pairs.data(); S.twz_dir = h->twz_dir.data(); S.twzfr_pairs = h->twzfr_pairs.data();
    S.twzfr_k = h->twzfr_k.data(); S.twzfr_init = h->twzfr_init.data(); S.twzfr_velo = h->twzfr_velo.data();
    S.rest_id = h->rest_id.data(); S.rest_rcut = h->rest_rcut.data(); S.rest_para = h->rest_para.data();
    return 0;
}

// rows from the product's predicates by an all-pairs sweep (what k_nl_rows evaluates per candidate)
int hc_build_lists(HostCheck *c, double *xyz, int all_pairs)
{
    sisgpu *h = c->h;
    const int N = h->nmp;
    const FFParams &P = h->P;
    if (h->P.flg_ele && h->ele_cutoff_type == 2) h->rep[0].ele_cutoff = h->ele_cutoff_inp * h->rep[0].lambdaD;
    fill_rp(h, 0, c->rp);
    c->pos.resize(N);
    for (int i = 0; i < N; i++) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        if (P.flg_pbc && !all_pairs) { x = wrap1(x, P.box[0]); y = wrap1(y, P.box[1]); z = wrap1(z, P.box[2]); }
        xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = z;
        c->pos[i] = make_double4(x, y, z, 0.0);
    }
    NlArgs A = nl_args(h, all_pairs != 0);
    A.all_pairs = all_pairs ? 1 : 0;
    const double ec = c->rp.ele_cutoff + A.nl_margin, ele_nl_cut2 = ec * ec;
    std::vector<std::vector<uint32_t>> rows(N);
    std::vector<std::vector<int>> brows(N);
    std::vector<std::vector<int>> btype(N), bidx(N);
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            if (i == j) continue;
            Vec3 v = pbc_vec_d(P, v3(c->pos[i]), v3(c->pos[j]));
            const double d2 = dot(v, v);
            uint32_t ent = nl_pair_flags(A, P, c->meta[i], c->meta[j], i, j, d2, ele_nl_cut2);
            if (ent && nl_needs_image(P, A, v3(c->pos[i]), v3(c->pos[j]))) ent |= NB_IMG;   // as k_nl_rows
            if (ent) rows[i].push_back(ent | (uint32_t)j);
            if (nl_bp_geometry(A, i, j, d2)) {
                int bptype = 0, bp3;
                if (i < j) bp3 = bp_lookup_r(host_restrict(h), A.bp_model, A.bp_min_loop, h->bp3_hash.data(), h->seq.data(), i, c->meta[i], j, c->meta[j], &bptype);
                else bp3 = bp_lookup_r(host_restrict(h), A.bp_model, A.bp_min_loop, h->bp3_hash.data(), h->seq.data(), j, c->meta[j], i, c->meta[i], &bptype);
                if (bp3 > 0) { brows[i].push_back(j); btype[i].push_back(bptype); bidx[i].push_back(bp3); }
            }
        }
    size_t K = 1, KB = 1;
    for (int i = 0; i < N; i++) { K = std::max(K, rows[i].size()); KB = std::max(KB, brows[i].size()); }
    c->K = (int)K; c->KB = (int)KB;
    c->nbr.assign((size_t)N * K, 0); c->nnbr.assign(N, 0); c->nbp_row.assign(N, 0);
    c->bprow.assign((size_t)N * KB, make_int2(0, 0));
    c->bp_type.assign((size_t)N * KB, 0); c->bp_coef.assign((size_t)N * KB, make_double2(0, 0));
    c->bp_status.assign((size_t)N * KB, 0); c->bp_status_MC.assign((size_t)N * KB, 1); c->ene_bp.assign((size_t)N * KB, 0.0);
    for (int i = 0; i < N; i++) {
        c->nnbr[i] = (int)rows[i].size();
        std::copy(rows[i].begin(), rows[i].end(), c->nbr.begin() + (size_t)i * K);
        c->nbp_row[i] = (int)brows[i].size();
        for (size_t e = 0; e < brows[i].size(); e++) {
            const int j = brows[i][e];
            int ibp;
            if (j > i) {
                ibp = i * (int)KB + (int)e;
                c->bp_type[ibp] = (signed char)btype[i][e];
                if (h->bp_model == 5) c->bp_coef[ibp] = make_double2(h->bp3_dH[bidx[i][e]], h->bp3_dS[bidx[i][e]]);
            } else {
                size_t q = std::find(brows[j].begin(), brows[j].end(), i) - brows[j].begin();
                ibp = j * (int)KB + (int)q;
            }
            c->bprow[(size_t)i * KB + e] = make_int2(j, ibp);
        }
    }
    DevSys &S = c->S;
    S.K = c->K; S.KB = c->KB;
    S.nbr = c->nbr.data(); S.nnbr = c->nnbr.data(); S.bprow = c->bprow.data(); S.nbp_row = c->nbp_row.data();
    S.bp_type = c->bp_type.data(); S.bp_coef = c->bp_coef.data(); S.bp_status = c->bp_status.data();
    S.bp_status_MC = c->bp_status_MC.data(); S.ene_bp = c->ene_bp.data();
    return 0;
}

int hc_list_sizes(HostCheck *c, int *nwca, int *nele, int *nbp)
{
    const int N = c->h->nmp;
    int a = 0, b = 0, d = 0;
    for (int i = 0; i < N; i++) {
        for (int e = 0; e < c->nnbr[i]; e++) {
            uint32_t ent = c->nbr[(size_t)i * c->K + e];
            if ((int)(ent & NB_IDX) > i) { a += (ent & NB_WCA) != 0; b += (ent & NB_ELE) != 0; }
        }
        for (int e = 0; e < c->nbp_row[i]; e++) d += c->bprow[(size_t)i * c->KB + e].x > i;
    }
    *nwca = a; *nele = b; *nbp = d;
    return 0;
}

// half lists, 1-based, sorted by (i, j); bp rows carry (i, j, type) and (dH, dS)
int hc_get_lists(HostCheck *c, int *wca, int *ele, int *bp, double *coef)
{
    const int N = c->h->nmp;
    size_t a = 0, b = 0, d = 0;
    for (int i = 0; i < N; i++) {
        for (int e = 0; e < c->nnbr[i]; e++) {
            uint32_t ent = c->nbr[(size_t)i * c->K + e];
            int j = ent & NB_IDX;
            if (j <= i) continue;
            if (ent & NB_WCA) { wca[2 * a] = i + 1; wca[2 * a + 1] = j + 1; a++; }
            if (ent & NB_ELE) { ele[2 * b] = i + 1; ele[2 * b + 1] = j + 1; b++; }
        }
        for (int e = 0; e < c->nbp_row[i]; e++) {
            int2 ent = c->bprow[(size_t)i * c->KB + e];
            if (ent.x <= i) continue;
            bp[3 * d] = i + 1; bp[3 * d + 1] = ent.x + 1; bp[3 * d + 2] = c->bp_type[ent.y];
            coef[2 * d] = c->bp_coef[ent.y].x; coef[2 * d + 1] = c->bp_coef[ent.y].y;
            d++;
        }
    }
    return 0;
}

// the product's force tasks run serially: pair + local per bead, bp per half-list candidate
int hc_force(HostCheck *c, const double *xyz, unsigned mask, double *forces)
{
    const int N = c->h->nmp;
    c->pos.resize(N);
    for (int i = 0; i < N; i++) c->pos[i] = make_double4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0);
    for (int k = 0; k < N; k++) {
        Vec3 F = pair_force<1>(c->S, c->h->P, c->rp, c->pos.data(), 0, k, 0, mask);
        F += local_force<1>(c->S, c->h->P, c->rp, c->pos.data(), 0, k, 0, mask, 0ull);
        forces[3 * k] = F.x; forces[3 * k + 1] = F.y; forces[3 * k + 2] = F.z;
    }
    if (mask & (1u << ENE_BP))
        for (int k = 0; k < N; k++)
            for (int e = 0; e < c->nbp_row[k]; e++) {
                if (c->bprow[(size_t)k * c->KB + e].x < k) continue;
                int beads[6];
                Vec3 f[6];
                const int n = bp_pair_force(c->S, c->h->P, c->rp, c->pos.data(), 0, k, e, 0, 1, beads, f);
                for (int s = 0; s < n; s++) {
                    forces[3 * beads[s]] += f[s].x; forces[3 * beads[s] + 1] += f[s].y; forces[3 * beads[s] + 2] += f[s].z;
                }
            }
    return 0;
}

// the role-split base-pair evaluation of k_md_fused (bp_force_role), roles combined serially
// with the same bead routing the device lanes use
int hc_force_bp_roles(HostCheck *c, const double *xyz, double *forces)
{
    const int N = c->h->nmp;
    const FFParams &P = c->h->P;
    c->pos.resize(N);
    for (int i = 0; i < N; i++) c->pos[i] = make_double4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0);
    for (int i = 0; i < 3 * N; i++) forces[i] = 0.0;
    for (int k = 0; k < N; k++)
        for (int e = 0; e < c->nbp_row[k]; e++) {
            const int2 ent = c->bprow[(size_t)k * c->KB + e];
            const int j = ent.x, ibp = ent.y;
            if (j < k) continue;
            const int type = c->bp_type[ibp];
            double dG = 1.0;
            if (P.bp_triplet) {
                dG = c->bp_coef[ibp].x - c->rp.tempK * c->bp_coef[ibp].y;
                if (dG >= 0.0) continue;
            }
            const BpPara &p = P.bp[type];
            const int4 mi = c->meta[k], mj = c->meta[j];
            const int beads[6] = {k, j, chain_nb(mi, -1), chain_nb(mj, -1), chain_nb(mi, +1), chain_nb(mj, +1)};
            const Vec3 x0 = v3(c->pos[k]), x1 = v3(c->pos[j]);
            const Vec3 v12 = pbc_vec_d(P, x0, x1);
            const double d1212 = dot(v12, v12), a12 = sqrt(d1212);
            if (fabs(a12 - p.bond_r) > p.cutoff_ddist) continue;
            double u = 0.0;
            Vec3 F[6];
            for (int s = 0; s < 6; s++) F[s] = mk(0.0, 0.0, 0.0);
            for (int role = 0; role < 8; role++) {
                const int tA = bp_role_A(role), tB = bp_role_B(role);
                const Vec3 xa = tA >= 0 ? v3(c->pos[beads[tA]]) : x0, xb = tB >= 0 ? v3(c->pos[beads[tB]]) : x0;
                const BpSub o = bp_force_role(P, p, role, x0, x1, xa, xb, v12, d1212, a12);
                u += o.u;
                F[0] += o.c0; F[1] += o.c1;
                if (tA >= 0) F[tA] += o.ca;
                if (tB >= 0) F[tB] += o.cb;
            }
            const double ene = P.bp_triplet ? p.U0 * dG * exp(-u) : p.U0 * exp(-u);
            if (ene > P.bp_cutoff_energy) continue;
            for (int s = 0; s < 6; s++) {
                forces[3 * beads[s]] += ene * F[s].x; forces[3 * beads[s] + 1] += ene * F[s].y; forces[3 * beads[s] + 2] += ene * F[s].z;
            }
        }
    return 0;
}

// the role-major base-pair evaluation of k_md_rep / k_md_fast (bp_role_eval, bp_roles_u, bp_roles_bead_force):
// seven role items per pair into BP_ROLE doubles, then energy and the six bead forces from them
int hc_force_bp_role_major(HostCheck *c, const double *xyz, double *forces)
{
    const int N = c->h->nmp;
    const FFParams &P = c->h->P;
    c->pos.resize(N);
    for (int i = 0; i < N; i++) c->pos[i] = make_double4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0);
    for (int i = 0; i < 3 * N; i++) forces[i] = 0.0;
    for (int k = 0; k < N; k++)
        for (int e = 0; e < c->nbp_row[k]; e++) {
            const int2 ent = c->bprow[(size_t)k * c->KB + e];
            const int j = ent.x, ibp = ent.y;
            if (j < k) continue;
            const int type = c->bp_type[ibp];
            double dG = 1.0;
            if (P.bp_triplet) {
                dG = c->bp_coef[ibp].x - c->rp.tempK * c->bp_coef[ibp].y;
                if (dG >= 0.0) continue;
            }
            const BpPara &p = P.bp[type];
            const int4 mi = c->meta[k], mj = c->meta[j];
            const int beads[6] = {k, j, chain_nb(mi, -1), chain_nb(mj, -1), chain_nb(mi, +1), chain_nb(mj, +1)};
            const Vec3 x0 = v3(c->pos[k]), x1 = v3(c->pos[j]);
            const Vec3 v12 = pbc_vec_d(P, x0, x1);
            if (fabs(sqrt(dot(v12, v12)) - p.bond_r) > p.cutoff_ddist) continue;
            double r[BP_ROLE];
            for (int role = 0; role < 7; role++) {
                const int tA = bp_role_A(role), tB = bp_role_B(role);
                const Vec3 xa = tA >= 0 ? v3(c->pos[beads[tA]]) : x0, xb = tB >= 0 ? v3(c->pos[beads[tB]]) : x0;
                bp_role_eval(P, p, role, x0, x1, xa, xb, r);
            }
            const double u = bp_roles_u(r);
            const double ene = P.bp_triplet ? p.U0 * dG * exp(-u) : p.U0 * exp(-u);
            if (ene > P.bp_cutoff_energy) continue;
            for (int t = 0; t < 6; t++) {
                const Vec3 F = bp_roles_bead_force(r, t);
                forces[3 * beads[t]] += ene * F.x; forces[3 * beads[t] + 1] += ene * F.y; forces[3 * beads[t] + 2] += ene * F.z;
            }
        }
    return 0;
}
// the fp32 pair scalar of the mixed-precision mode against the fp64 scalars, for one squared distance
double hc_pair_scalar(HostCheck *c, double d2, int wca, int ele, int mixed)
{
    const FFParams &P = c->h->P;
    if (mixed) return pair_scalar_f32(P, wca != 0, ele != 0, d2, c->rp.ele_coef, c->rp.inv_lambdaD);
    double s = 0.0;
    if (wca) s += wca_scalar(P, d2);
    if (ele) s += dh_scalar(c->rp.ele_coef, c->rp.inv_lambdaD, d2);
    return s;
}

double hc_gauss(unsigned long long seed, unsigned stream, unsigned long long step, unsigned bead, int comp)
{
    return gauss_component(seed, stream, step, bead, comp);
}
double hc_uniform(unsigned long long seed, unsigned stream, unsigned long long counter)
{
    return uniform_counter(seed, stream, counter);
}
int hc_coef(HostCheck *c, int imp, double *out)
{
    double4 v = c->coef[imp];
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w * c->rp.sqrt_tempK;
    return 0;
}
// REMD host logic (tables, acceptance, label -> parameter refresh) of rank `rank` of `nranks`
int hc_remd_setup(HostCheck *c, const int *nrep, const double *values, int nstep_exchange, int rank, int nranks)
{
    return remd_setup_host(c->h, nrep, values, nstep_exchange, rank, nranks);
}
// scale_out[nrep_proc]: product of the velocity factors the exchange applies to each local replica;
// params_out[nrep_proc][4]: tempK, ionic strength, DCF force, Rg0 after the refresh
int hc_remd_exchange(HostCheck *c, const double *E, const double *uniforms, int *natt, int *nacc, double *scale_out, double *params_out)
{
    sisgpu *h = c->h;
    std::vector<VelScale> sc;
    remd_decide(h, E, uniforms, natt, nacc, sc);
    for (int r = 0; r < h->nrep; r++) scale_out[r] = 1.0;
    for (const VelScale &v : sc) scale_out[v.irep] *= v.factor;
    for (int r = 0; r < h->nrep; r++) {
        params_out[4 * r] = h->rep[r].tempK; params_out[4 * r + 1] = h->rep[r].ionic;
        params_out[4 * r + 2] = h->rep[r].twz_force; params_out[4 * r + 3] = h->rep[r].rg0;
    }
    return 0;
}
int hc_rep_params(HostCheck *c, double *out)
{
    out[0] = c->h->rep[0].ele_coef; out[1] = c->h->rep[0].lambdaD; out[2] = c->h->rep[0].ele_cutoff; out[3] = c->h->rep[0].diele;
    return 0;
}
}
