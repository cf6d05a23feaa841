// sisgpu.cu -- host side of libsisgpu.so: the C ABI of include/sisgpu.h.
//
// Mirrors the call structure of the reference's drivers (src/job_md.F90, src/force.F90,
// src/neighbor_list.F90, src/energy_sumup.F90, src/energy_replica.F90,
// src/replica_exchange.F90) but keeps all state resident on the device and batches every
// local replica into one launch (grid.y = replica).  No CPU fallback exists: every entry point
// that computes runs CUDA kernels or fails.
#include "../../include/sisgpu.h"
#include "sis_kernels.cuh"
#include "sis_fused.cuh"
#include "sis_fast.cuh"
#include "sis_big.cuh"
#include "sis_rep.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

static thread_local char g_err[1024] = "";
static int set_err(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return 1;
}
#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return set_err("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define TRY(call)                 \
    do {                          \
        int rc_ = (call);         \
        if (rc_) return rc_;      \
    } while (0)

// ---- minimal NCCL binding (dlopen: the library must load on boxes without NCCL) ------------
struct NcclId { char internal[SISGPU_NCCL_ID_BYTES]; };   // ncclUniqueId (passed by value)
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load()
{
    if (g_nccl.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return set_err("cannot dlopen libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (int (*)(void *))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather) return set_err("NCCL symbols missing");
    return 0;
}
constexpr int NCCL_FLOAT64 = 8;   // ncclDouble in nccl.h

// ---- host-side mirror of the reference's module state ----------------------------------------
enum { REPT_TEMP = 0, REPT_TWZDCF = 1, REPT_ION = 2, REPT_RG = 3, REPT_MAX = 4 };
constexpr int MAX_REP_PER_DIM = 256, MAX_REP_DIM = 3;

struct HostRep {
    double tempK = 300.0, ionic = 0.15, twz_force = 0.0, rg0 = 0.0;
    double ele_coef = 0, lambdaD = 1, ele_cutoff = 0, diele = 0, diele_dTcoef = 0;
};

template <class T>
struct DBuf {   // device buffer
    T *p = nullptr;
    size_t n = 0;
    int alloc(size_t count)
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = count;
        if (count == 0) count = 1;
        CK(cudaMalloc(&p, count * sizeof(T)));
        CK(cudaMemset(p, 0, count * sizeof(T)));
        return 0;
    }
    int upload(const std::vector<T> &v)
    {
        TRY(alloc(v.size()));
        if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

constexpr int BAR_WORDS = 64;

struct sisgpu {
    int device = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: the bp task branch of the step
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaGraphExec_t graph_chunk = nullptr;               // `chunk` MD steps as one graph launch
    int graph_chunk_steps = 0;
    cudaGraphExec_t graph_big = nullptr;                 // `big_chunk` launches of k_md_big (large systems)
    int graph_big_steps = 0, graph_big_bpc = 0;
    int big_kernel = 1, big_chunk = 16;                  // options "big_kernel" (0: k_md_fused instead), "big_chunk" (even)
    bool big_smem_set = false;
    bool finalized = false;
    // topology
    int nmp = 0, nchains = 0;
    std::vector<int> nmp_chain, moltypes, chain_start, seq, ichain, lmp, has_charge, is_frozen;
    std::vector<double> mass;
    // force field
    FFParams P{};
    double bp_cutoff_dist = 0;
    int bp_model = 5, bp_min_loop = 3, max_bp_per_nt = 1, nstep_bp_MC = 0;
    double NN_dH[22]{}, NN_dS[22]{}, dH0 = 0, dS0 = 0, NNend_dH[7]{}, NNend_dS[7]{}, dHend0 = 0, dSend0 = 0;
    int flg_NNend = 0, have_nn = 0;
    // ct / bpseq / bpl restriction (src/read_ss.F90): sorted keys imp * nmp + jmp of the accepted pairs
    bool flg_ss = false;
    std::vector<unsigned long long> ss_keys;
    DBuf<unsigned long long> d_ss_keys;
    std::vector<int> bp3_hash;          // 4096
    std::vector<double> bp3_dH, bp3_dS; // 469 (1-based)
    int bp3_n = 0;
    // ele
    int ele_cutoff_type = 1, excl_bond = 1, excl_angle = 1;
    double ele_cutoff_inp = 0, length_per_charge = 0, tempK_ref = 0;
    // md
    double dt = 0.05, viscosity_Pas = 1e-5, nl_margin = 10.0;
    unsigned long long seed = 0;
    int flg_bp_MC = 0;
    bool lists_full = false;   // list_exv/ele/bp mode (no rebuild trigger)
    bool lists_built = false;
    int nbp_half_max = 0;      // max over replicas of the half-list bp candidate count
    int coop_blocks = 0;       // co-resident grid size of the persistent kernel
    int fast_kernel = 1;       // option "fast_kernel": 0 forces the general k_md_fused
    size_t fast_smem_set = 0, part_smem_set = 0;
    void *fast_fn_set = nullptr;
    std::vector<int> nbp_half_rep;   // half-list bp candidate count per replica (last build)
    // biases (host copies)
    std::vector<int> twz_pairs, twzfr_pairs, rest_id;
    std::vector<double> twz_dir, twzfr_k, twzfr_init, twzfr_velo, rest_rcut, rest_para;
    // replicas
    int nrep = 1;
    std::vector<HostRep> rep;
    std::vector<RepParams> rp_host;
    // REMD
    bool flg_replica = false;
    int nrep_all = 1, ndim_replica = 0, nrep_dim[REPT_MAX]{1, 1, 1, 1}, rank = 0, nranks = 1;
    bool flg_repvar[REPT_MAX]{};
    int type_array[MAX_REP_DIM]{};
    std::vector<int> pair_tb, rep2lab, lab2rep;   // pair_tb [2*ndim][nrep_all]
    std::vector<double> lab2val;                  // [REPT_MAX][nrep_all]
    double replica_values[REPT_MAX][MAX_REP_PER_DIM]{};
    int idx_pair = 1, nstep_exchange = 0;
    unsigned long long exchange_counter = 0;
    long long n_attempt = 0, n_accept = 0;
    void *nccl_comm = nullptr;
    // device
    DevSys S{};
    DBuf<int4> d_meta;
    DBuf<int> d_seq, d_hash, d_nnbr, d_nbp_row, d_bp_status, d_bp_status_MC, d_nt_excess, d_cell_of,
        d_cell_count, d_cell_start, d_cell_cursor, d_cell_beads, d_rep_active, d_twz_pairs,
        d_twzfr_pairs, d_rest_id;
    DBuf<double4> d_coef, d_pos0, d_pos1;
    DBuf<double> d_f_pair, d_f_loc, d_f_bp;
    DBuf<int> d_bp_half, d_nbp_half, d_bp_status_save;   // _save: bp_status / ene_bp around the cross-energy pass
    DBuf<double> d_ene_bp_save;
    DBuf<double> d_mass, d_bp3_dH, d_bp3_dS, d_vel, d_acc, d_move, d_force, d_ene_bp,
        d_ene_partial, d_ene_out, d_kin, d_twz_dir, d_twzfr_k, d_twzfr_init, d_twzfr_velo, d_rest_rcut,
        d_rest_para, d_gather_send, d_gather_recv;
    DBuf<uint32_t> d_nbr;
    DBuf<int2> d_bprow;
    DBuf<signed char> d_bp_type;
    DBuf<double2> d_bp_coef, d_top2_slots;
    DBuf<unsigned long long> d_mc_counter, d_prof;
    DBuf<unsigned int> d_ene_counter, d_bar;
    DBuf<int> d_cta_item0, d_cell_sorted;
    DBuf<double> d_xyz_stage;            // [nmp][3] scratch of set_state / get_state
    DBuf<GridInfo> d_grid;
    GridInfo grid_host{};                // last grid written to d_grid by the host
    bool grid_host_valid = false;
    int dev_nrebuild_seen = 0;           // DevMisc::nrebuild at the last sync
    bool last_launch_local = false;
    const char *last_step_kernel = "k_force_local + k_force_bp + k_integrate";   // what md_steps launched last
    // device-resident REMD state (exchange steps inside md_steps without host round trips)
    int remd_inkernel = 1;               // option "remd_inkernel": exchange steps inside the replica-local kernel (one launch per md_steps call)
    DBuf<double> d_xch;                  // {E[2][n][2]; arrival counter (u64)}: one allocation = one IPC handle
    void *peer_xch[REMD_MAX_PEERS]{};    // the same buffer of every rank of the node (own included), sisgpu_peer_import
    bool peers_ready = false;
    unsigned long long xch_rounds = 0;   // exchange rounds done inside the kernel so far (identical on every rank)
    int remd_device = 1;                 // option "remd_device"
    bool remd_on_device = false;         // inside an md_steps call whose exchange steps run on the device: host tables are stale
    double remd_cut_max = 0.0;           // largest ele_cutoff any label can bring to a local replica (cutoff_type 2)
    DBuf<RemdDev> d_remd;
    DBuf<int> d_pair_tb, d_rep2lab, d_lab2rep, d_has_opp;
    DBuf<double> d_lab_temp, d_vscale;
    DBuf<RepParams> d_lab_rp;
    // output staging (sisgpu_snapshot_begin / _end): two slots of device + pinned buffers, copies on stream2
    struct SnapSlot {
        float *d_xyz = nullptr, *d_e = nullptr, *h_xyz = nullptr, *h_e = nullptr;
        int *d_cnt = nullptr, *d_nrec = nullptr, *h_nrec = nullptr;
        int2 *d_ij = nullptr, *h_ij = nullptr;
        double *h_ene = nullptr;          // 11 energies + Ekin
        int cap = 0, flags = 0, irep = 0;
        bool busy = false;
        cudaEvent_t staged = nullptr, copied = nullptr;
    } snap[2];
    int snap_next = 0;
    // pending device-side timings: (start, end, accumulator), resolved at the next synchronisation
    struct Pending { cudaEvent_t a, b; double *acc; };
    std::vector<Pending> ev_pending;
    std::vector<cudaEvent_t> ev_free;
    int fast_max_cta = 0;                // option "fast_max_cta": upper bound of the k_md_fast grid (0 = one CTA per SM)
    int mc_persistent = 1;               // option "mc_persistent": Monte-Carlo base-pair limit inside the replica-local k_md_fast
    int local_kernel = 1;                // option "local_kernel": replica-local mode of k_md_fast (one CTA per replica)
    int local_ncand = -1;                // sequence-allowed base pairs of one replica (bound of the per-CTA candidates)
    int rep_kernel = 1;                  // option "rep_kernel": k_md_rep (right-sized CTA per replica, static pair tables) in the replica-local mode
    int rep_threads = 0;                 // option "rep_threads": 128 / 256 threads per replica (0: by the number of replicas)
    size_t rep_smem_set[6] = {0, 0, 0, 0, 0, 0};
    int pipeline = 1;                    // option "pipeline": device-driven list rebuilds inside md_steps
    double est_cycle_steps = 50.0;       // running estimate of MD steps between two list rebuilds
    DBuf<RepParams> d_rp, d_rp_alt;
    DBuf<unsigned char> d_sync;          // DevSyncHdr + top2[nrep][2][2]
    unsigned char *h_sync = nullptr;     // pinned mirror
    size_t sync_bytes = 0;
    DevCtl *h_ctl = nullptr;             // -> ctl[0] ("A") inside h_sync
    DevMisc *h_misc = nullptr;
    unsigned long long *h_top2 = nullptr;
    std::vector<int> h_need;             // replicas past the list margin (from h_top2)
    int K = 0, KB = 0, G = 8;   // G = lanes per bead in the pair task
    // stats / options
    long long launches = 0, nrebuild = 0, graph_launches_per_chunk = 0;
    int launch_mode = 2, chunk = 16, timing = 0;   // launch_mode 2: persistent kernels (k_md_fast / k_md_fused); 1: CUDA-graph chunks; 0: stream launches
    double t_step = 0, t_nl = 0, t_ene = 0, t_rex = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_mark[2] = {nullptr, nullptr};
    long long istep_host = 0;
};

static void snap_free(sisgpu::SnapSlot &s);

const char *sisgpu_last_error(void) { return g_err; }
int sisgpu_version(void) { return 100; }

int sisgpu_create(sisgpu_t **out, int device)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err("no CUDA device available (%s); libsisgpu has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return set_err("device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    sisgpu *h = new sisgpu();
    h->device = device;
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    h->P.integrator = 1;
    h->P.bp_cutoff_energy = 0.001;
    *out = h;
    return 0;
}

int sisgpu_destroy(sisgpu_t *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->nccl_comm);
    if (h->h_sync) cudaFreeHost(h->h_sync);
    // DBuf members are released explicitly (no destructors touching CUDA at exit)
    h->d_meta.release(); h->d_seq.release(); h->d_hash.release(); h->d_nnbr.release(); h->d_nbp_row.release();
    h->d_bp_status.release(); h->d_bp_status_MC.release(); h->d_nt_excess.release(); h->d_cell_of.release();
    h->d_cell_count.release(); h->d_cell_start.release(); h->d_cell_cursor.release(); h->d_cell_beads.release();
    h->d_rep_active.release(); h->d_twz_pairs.release(); h->d_twzfr_pairs.release();
    h->d_rest_id.release(); h->d_coef.release(); h->d_pos0.release(); h->d_pos1.release(); h->d_mass.release();
    h->d_bp3_dH.release(); h->d_bp3_dS.release(); h->d_vel.release(); h->d_acc.release(); h->d_move.release();
    h->d_force.release(); h->d_ene_bp.release(); h->d_ene_partial.release();
    h->d_ene_out.release(); h->d_kin.release(); h->d_twz_dir.release(); h->d_twzfr_k.release();
    h->d_twzfr_init.release(); h->d_twzfr_velo.release(); h->d_rest_rcut.release(); h->d_rest_para.release();
    h->d_gather_send.release(); h->d_gather_recv.release(); h->d_nbr.release(); h->d_bprow.release();
    h->d_bp_type.release(); h->d_bp_coef.release(); h->d_mc_counter.release();
    h->d_ene_counter.release(); h->d_bar.release(); h->d_top2_slots.release(); h->d_cta_item0.release(); h->d_cell_sorted.release(); h->d_grid.release(); h->d_xyz_stage.release(); h->d_ss_keys.release();
    snap_free(h->snap[0]); snap_free(h->snap[1]);
    for (auto &p : h->ev_pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    h->ev_pending.clear();
    for (auto e : h->ev_free) cudaEventDestroy(e);
    h->ev_free.clear();
    h->d_remd.release(); h->d_pair_tb.release(); h->d_rep2lab.release(); h->d_lab2rep.release(); h->d_has_opp.release();
    h->d_lab_temp.release(); h->d_vscale.release(); h->d_lab_rp.release(); for (int r = 0; r < REMD_MAX_PEERS; r++) if (h->peer_xch[r] && r != h->rank && h->peers_ready) cudaIpcCloseMemHandle(h->peer_xch[r]);
    h->d_xch.release();
    h->d_rp.release(); h->d_rp_alt.release(); h->d_sync.release();
    if (h->graph_chunk) cudaGraphExecDestroy(h->graph_chunk);
    if (h->graph_big) cudaGraphExecDestroy(h->graph_big);
    h->d_f_pair.release(); h->d_f_loc.release(); h->d_f_bp.release(); h->d_bp_half.release(); h->d_nbp_half.release();
    h->d_bp_status_save.release(); h->d_ene_bp_save.release();
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaEventDestroy(h->ev_fork);
    cudaEventDestroy(h->ev_join);
    cudaStreamDestroy(h->stream2);
    cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

// kernel arguments (DevSys, FFParams, StepArgs) are baked into the instantiated graph: anything
// that changes them drops it
static void invalidate_graphs(sisgpu *h)
{
    if (h->graph_chunk) { cudaGraphExecDestroy(h->graph_chunk); h->graph_chunk = nullptr; }
    if (h->graph_big) { cudaGraphExecDestroy(h->graph_big); h->graph_big = nullptr; }
}

// ---- set-up -------------------------------------------------------------------------------
int sisgpu_set_topology(sisgpu_t *h, int nmp, int nchains, const int *nmp_chain, const int *moltypes,
                        const int *seq, const int *has_charge, const int *is_frozen, const double *mass)
{
    if (nmp <= 0 || nchains <= 0) return set_err("empty system");
    if (nmp >= (1 << 29)) return set_err("nmp too large for 29-bit row entries");
    int tot = 0;
    for (int c = 0; c < nchains; c++) tot += nmp_chain[c];
    if (tot != nmp) return set_err("sum(nmp_chain)=%d != nmp=%d", tot, nmp);
    h->nmp = nmp; h->nchains = nchains;
    h->nmp_chain.assign(nmp_chain, nmp_chain + nchains);
    h->moltypes.assign(moltypes, moltypes + nchains);
    h->seq.assign(seq, seq + nmp);
    h->has_charge.assign(nmp, 1);
    if (has_charge) h->has_charge.assign(has_charge, has_charge + nmp);
    h->is_frozen.assign(nmp, 0);
    if (is_frozen) h->is_frozen.assign(is_frozen, is_frozen + nmp);
    h->chain_start.resize(nchains); h->ichain.resize(nmp); h->lmp.resize(nmp);
    int imp = 0;
    for (int c = 0; c < nchains; c++) {
        h->chain_start[c] = imp;
        for (int i = 0; i < nmp_chain[c]; i++, imp++) { h->ichain[imp] = c; h->lmp[imp] = i; }
    }
    h->mass.resize(nmp);
    static const double m5[5] = {328.212, 305.164, 344.212, 304.182, 320.443};   // src/job_md.F90:116-133
    for (int i = 0; i < nmp; i++) {
        if (seq[i] < 0 || seq[i] > 4) return set_err("Unknown seq type %d at imp=%d", seq[i], i + 1);
        h->mass[i] = mass ? mass[i] : m5[seq[i]];
    }
    return 0;
}

int sisgpu_set_replicas(sisgpu_t *h, int nrep_proc)
{
    if (nrep_proc < 1) return set_err("nrep_proc must be >= 1");
    h->nrep = nrep_proc;
    h->rep.assign(nrep_proc, HostRep());
    return 0;
}

int sisgpu_set_ff_bonded(sisgpu_t *h, double bond_k, double bond_r0, int angl_type, double angl_k,
                         double angl_t0, int dih_flags, double dih_k, double dih_w, double dih_p0)
{
    FFParams &P = h->P;
    P.bond_k = bond_k; P.bond_r0 = bond_r0;
    P.flg_angl_ReB = angl_type == 1;
    P.angl_k = angl_k; P.angl_t0 = angl_t0; P.angl_co0 = cos(angl_t0);
    P.flg_dih_cos = dih_flags & 1; P.flg_dih_exp = (dih_flags >> 1) & 1;
    P.dih_k = dih_k; P.dih_w = dih_w; P.dih_p0 = dih_p0;
    return 0;
}

int sisgpu_set_ff_wca(sisgpu_t *h, double sigma, double eps)
{
    FFParams &P = h->P;
    P.wca_sigma = sigma; P.wca_eps = eps;
    P.wca_sigma2 = sigma * sigma;
    P.wca_coef = 12.0 * eps / P.wca_sigma2;   // src/force_wca.F90:40-41
    return 0;
}

int sisgpu_set_ff_bp(sisgpu_t *h, int model, int min_loop, int max_bp_per_nt, int nstep_MC,
                     double cutoff_energy, const double *q)
{
    if (model != 1 && model != 3 && model != 4 && model != 5) return set_err("invalid bp model %d", model);
    h->bp_model = model; h->bp_min_loop = min_loop; h->max_bp_per_nt = max_bp_per_nt; h->nstep_bp_MC = nstep_MC;
    FFParams &P = h->P;
    P.bp_cutoff_energy = cutoff_energy;
    P.bp_triplet = (model == 4 || model == 5);
    P.nstep_bp_MC = nstep_MC;
    for (int t = 0; t < 3; t++) {
        BpPara b{};
        const double *v = q + 15 * t;
        b.U0 = v[0]; b.bond_k = v[1]; b.bond_r = v[2];
        b.angl_k1 = v[3]; b.angl_k2 = v[4]; b.angl_k3 = v[5]; b.angl_k4 = v[6];
        b.angl_theta1 = v[7]; b.angl_theta2 = v[8]; b.angl_theta3 = v[9]; b.angl_theta4 = v[10];
        b.dihd_k1 = v[11]; b.dihd_k2 = v[12]; b.dihd_phi1 = v[13]; b.dihd_phi2 = v[14];
        b.cos_phi1 = cos(b.dihd_phi1); b.sin_phi1 = sin(b.dihd_phi1); b.cos_phi2 = cos(b.dihd_phi2); b.sin_phi2 = sin(b.dihd_phi2);
        P.bp[1 + t] = b;
        BpPara s = b;   // CG / UA / UG: src/read_force_field.F90:203-207
        std::swap(s.angl_k1, s.angl_k2);
        std::swap(s.angl_k3, s.angl_k4);
        P.bp[4 + t] = s;
    }
    return 0;
}

// host view of the restriction (hostcheck, setup-time counts)
static BpRestrict host_restrict(const sisgpu *h)
{
    BpRestrict r;
    r.keys = h->ss_keys.data(); r.n_ss = h->flg_ss ? (int)h->ss_keys.size() : -1; r.nmp = h->nmp;
    return r;
}
static int upload_restrict(sisgpu *h)
{
    if (!h->finalized) return 0;
    TRY(h->d_ss_keys.upload(h->ss_keys));
    h->S.rs.keys = h->d_ss_keys.p; h->S.rs.n_ss = h->flg_ss ? (int)h->ss_keys.size() : -1; h->S.rs.nmp = h->nmp;
    h->lists_built = false;
    // bound of the base-pair candidates of one replica (replica-local mode of k_md_fast)
    if (h->flg_ss && h->nmp <= FAST_GROUPS) h->local_ncand = (int)h->ss_keys.size();
    return 0;
}
// src/read_ss.F90:120-210: the pairs of a ct / bpseq / bpl file that enter bp_map (1-based (imp, jmp),
// any order); n < 0 removes the restriction.  Call after set_topology and set_ff_bp.
int sisgpu_set_bp_restrict(sisgpu_t *h, int n, const int *pairs)
{
    if (h->nmp <= 0) return set_err("set_topology must come first");
    h->flg_ss = n >= 0;
    h->ss_keys.clear();
    for (int k = 0; k < n; k++) {
        int imp = pairs[2 * k] - 1, jmp = pairs[2 * k + 1] - 1;
        if (imp > jmp) std::swap(imp, jmp);
        if (imp < 0 || jmp >= h->nmp || imp == jmp) continue;
        const int ic = h->ichain[imp], jc = h->ichain[jmp];
        const int i = h->lmp[imp] + 1, j = h->lmp[jmp] + 1;
        if (i == 1 || i == h->nmp_chain[ic]) continue;                  // chain ends
        if (j == 1 || j == h->nmp_chain[jc]) continue;
        if (ic == jc && i + h->bp_min_loop >= j) continue;               // minimum loop length
        if (h->bp_model == 3 || h->bp_model == 4 || h->bp_model == 5)    // isolated base pair
            if (!is_complement(h->seq[imp - 1], h->seq[jmp + 1]) && !is_complement(h->seq[imp + 1], h->seq[jmp - 1])) continue;
        if (!bp_type_of(h->seq[imp], h->seq[jmp])) continue;             // no known type
        h->ss_keys.push_back((unsigned long long)imp * h->nmp + jmp);
    }
    std::sort(h->ss_keys.begin(), h->ss_keys.end());
    h->ss_keys.erase(std::unique(h->ss_keys.begin(), h->ss_keys.end()), h->ss_keys.end());
    return upload_restrict(h);
}

int sisgpu_set_ff_nn(sisgpu_t *h, const double *NN_dH, const double *NN_dS, double dH0, double dS0,
                     int flg_NNend, const double *NNend_dH, const double *NNend_dS, double dHend0, double dSend0)
{
    for (int i = 0; i < 21; i++) { h->NN_dH[i + 1] = NN_dH[i]; h->NN_dS[i + 1] = NN_dS[i]; }
    h->dH0 = dH0; h->dS0 = dS0; h->flg_NNend = flg_NNend;
    for (int i = 0; i < 6; i++) {
        h->NNend_dH[i + 1] = NNend_dH ? NNend_dH[i] : 0.0;
        h->NNend_dS[i + 1] = NNend_dS ? NNend_dS[i] : 0.0;
    }
    h->dHend0 = dHend0; h->dSend0 = dSend0; h->have_nn = 1;
    return 0;
}

int sisgpu_set_ele(sisgpu_t *h, int flg_ele, int cutoff_type, double cutoff_inp, int excl_bond, int excl_angle,
                   double length_per_charge, int temp_independent, double tempK_ref)
{
    if (flg_ele && cutoff_type != 1 && cutoff_type != 2) return set_err("Invalid ele_cutoff_type value in set_ele.");
    h->P.flg_ele = flg_ele; h->ele_cutoff_type = cutoff_type; h->ele_cutoff_inp = cutoff_inp;
    h->excl_bond = excl_bond; h->excl_angle = excl_angle; h->length_per_charge = length_per_charge;
    h->P.temp_independent = temp_independent; h->tempK_ref = tempK_ref;
    return 0;
}

int sisgpu_set_pbc(sisgpu_t *h, int flg_pbc, const double *box)
{
    h->P.flg_pbc = flg_pbc;
    for (int i = 0; i < 3; i++) {
        h->P.box[i] = box ? box[i] : 0.0;
        h->P.box_half[i] = 0.5 * h->P.box[i];
    }
    if (flg_pbc && (h->P.box[0] <= 0 || h->P.box[1] <= 0 || h->P.box[2] <= 0)) return set_err("invalid PBC box");
    invalidate_graphs(h);
    return 0;
}

int sisgpu_set_md(sisgpu_t *h, int integrator, double dt, double viscosity_Pas, double nl_margin, uint64_t seed)
{
    if (integrator != 1 && integrator != 2) return set_err("unknown integrator %d", integrator);
    h->P.integrator = integrator; h->dt = dt; h->viscosity_Pas = viscosity_Pas; h->nl_margin = nl_margin;
    h->seed = seed;
    invalidate_graphs(h);
    return 0;
}

int sisgpu_set_bias_ss(sisgpu_t *h, int flg, double force) { h->P.flg_bias_ss = flg; h->P.bias_ss_force = force; return 0; }
int sisgpu_set_stage(sisgpu_t *h, int flg, double sigma, double eps)
{
    h->P.flg_stage = flg; h->P.stage_sigma = sigma; h->P.stage_eps = eps;
    return 0;
}
int sisgpu_set_tweezers_dcf(sisgpu_t *h, int n, const int *pairs, const double *direction)
{
    h->P.ntwz_DCF = n;
    h->twz_pairs.resize(2 * n);
    for (int i = 0; i < 2 * n; i++) h->twz_pairs[i] = pairs[i] - 1;
    h->twz_dir.assign(direction, direction + 3 * n);
    return 0;
}
int sisgpu_set_tweezers_fr(sisgpu_t *h, int n, const int *pairs, const double *k, const double *init, const double *velo)
{
    h->P.ntwz_FR = n;
    h->twzfr_pairs.resize(2 * n);
    for (int i = 0; i < 2 * n; i++) h->twzfr_pairs[i] = pairs[i] - 1;
    h->twzfr_k.assign(k, k + 2 * n);
    h->twzfr_init.assign(init, init + 6 * n);
    h->twzfr_velo.assign(velo, velo + 6 * n);
    return 0;
}
int sisgpu_set_bias_rg(sisgpu_t *h, int flg, int pott, double k)
{
    h->P.flg_bias_rg = flg; h->P.bias_rg_pott = pott; h->P.bias_rg_k = k;
    return 0;
}
int sisgpu_set_restraints(sisgpu_t *h, int n, const int *id, const double *rcut, const double *para)
{
    h->P.nrest = n;
    h->rest_id.resize(3 * n);
    for (int i = 0; i < n; i++) {
        h->rest_id[3 * i] = id[3 * i] - 1; h->rest_id[3 * i + 1] = id[3 * i + 1] - 1; h->rest_id[3 * i + 2] = id[3 * i + 2];
    }
    h->rest_rcut.assign(rcut, rcut + n);
    h->rest_para.assign(para, para + 3 * n);
    return 0;
}

// set_ele: src/set_ele.F90:47-102 (the non-_HTN_CONSISTENT branch the reference builds by default)
static void host_set_ele(const sisgpu *h, HostRep &r, double tempk, double ionstr)
{
    const double MM_A = 87.740, MM_B = -0.40008, MM_C = 9.398e-4, MM_D = -1.410e-6;
    const double PI = 4.0 * atan(1.0);
    if (h->P.temp_independent == 0) {
        double Tc = tempk - 273.15;
        r.diele = MM_A + MM_B * Tc + MM_C * Tc * Tc + MM_D * Tc * Tc * Tc;
    } else {
        double Tc = h->tempK_ref - 273.15;
        r.diele = MM_A + MM_B * Tc + MM_C * Tc * Tc + MM_D * Tc * Tc * Tc;
        r.diele_dTcoef = 1.0 + h->tempK_ref / r.diele * (MM_B + 2.0 * MM_C * Tc + 3.0 * MM_D * Tc * Tc);
    }
    double lb = SIS_ELE * SIS_ELE / (4.0 * PI * SIS_EPS0 * r.diele * SIS_BOLTZ_J * tempk) * 1.0e10;
    double Zp = h->length_per_charge / lb;
    r.ele_coef = SIS_JOUL2KCAL_MOL * 1.0e10 * SIS_ELE * SIS_ELE / (4.0 * PI * SIS_EPS0 * r.diele);
    r.ele_coef = r.ele_coef * Zp * Zp;
    r.lambdaD = 1.0e10 * sqrt((1.0e-3 * SIS_EPS0 * r.diele * SIS_BOLTZ_J) / (2.0 * SIS_N_AVO * SIS_ELE * SIS_ELE)) *
                sqrt(tempk / ionstr);
}

static void fill_rp(const sisgpu *h, int irep, RepParams &R)
{
    const HostRep &r = h->rep[irep];
    memset(&R, 0, sizeof R);
    R.tempK = r.tempK;
    R.ele_coef = r.ele_coef;
    R.inv_lambdaD = r.lambdaD != 0.0 ? 1.0 / r.lambdaD : 0.0;
    R.lambdaD = r.lambdaD;
    R.ele_cutoff = r.ele_cutoff;
    R.diele_dTcoef = r.diele_dTcoef;
    R.sqrt_tempK = sqrt(r.tempK);
    R.bias_rg_0 = r.rg0;
    R.twz_force = r.twz_force;
    R.stream = h->rank * h->nrep + irep;
}

static int push_rp(sisgpu *h)
{
    if (!h->finalized) return 0;
    h->rp_host.resize(h->nrep);
    for (int i = 0; i < h->nrep; i++) fill_rp(h, i, h->rp_host[i]);
    CK(cudaMemcpyAsync(h->d_rp.p, h->rp_host.data(), h->nrep * sizeof(RepParams), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int sisgpu_set_replica_params(sisgpu_t *h, int irep, double tempK, double ionic_strength, double twz_force, double rg0)
{
    if ((int)h->rep.size() != h->nrep) h->rep.assign(h->nrep, HostRep());
    if (irep < 0 || irep >= h->nrep) return set_err("irep %d out of range", irep);
    HostRep &r = h->rep[irep];
    r.tempK = tempK; r.ionic = ionic_strength; r.twz_force = twz_force; r.rg0 = rg0;
    if (h->P.flg_ele) {
        host_set_ele(h, r, tempK, ionic_strength);
        if (h->ele_cutoff_type == 1) r.ele_cutoff = h->ele_cutoff_inp;
        else if (r.ele_cutoff == 0.0) r.ele_cutoff = h->ele_cutoff_inp * r.lambdaD;   // src/init_ele.F90:88-90
    }
    return push_rp(h);
}

// 21 nearest-neighbour stacks in NNT order (src/const_idx.F90:96-121); "XY_WZ" = 5'-XY-3' / 3'-WZ-5'
static int nn_index(int x, int y, int w, int z)
{
    static const char *names[21] = {"GC_CG", "CC_GG", "GA_CU", "CG_GC", "AC_UG", "CA_GU", "AG_UC", "UA_AU", "AU_UA",
                                    "AA_UU", "GC_UG", "CU_GG", "GG_CU", "CG_GU", "AU_UG", "GA_UU", "UG_GU", "UA_GU",
                                    "GG_UU", "GU_UG", "AG_UU"};
    static const char S[4] = {'A', 'U', 'G', 'C'};
    const char a[6] = {S[x], S[y], '_', S[w], S[z], 0};
    const char b[6] = {S[z], S[w], '_', S[y], S[x], 0};   // the same stack read from the other strand
    for (int n = 0; n < 21; n++)
        if (!strcmp(a, names[n]) || !strcmp(b, names[n])) return n + 1;
    return 0;
}

// the 468-entry triplet table of src/init_bp.F90:39-172, enumerated in the same order so that the
// bp3 indices (and with them bp_coef) agree with the reference's
static int build_bp3_table(sisgpu *h)
{
    h->bp3_hash.assign(4096, 0);
    h->bp3_dH.assign(469, 0.0);
    h->bp3_dS.assign(469, 0.0);
    auto end_type = [](int x, int y, int w, int z) -> int {   // NNENDT, src/const_idx.F90:123-132
        auto is = [](int a, int b, int p, int q) { return (a == p && b == q) || (a == q && b == p); };
        bool au = is(x, y, SEQ_A, SEQ_U), gu = is(x, y, SEQ_G, SEQ_U);
        int on = is(w, z, SEQ_A, SEQ_U) ? 0 : is(w, z, SEQ_G, SEQ_C) ? 1 : is(w, z, SEQ_G, SEQ_U) ? 2 : -1;
        if (on < 0) return 0;
        if (au) return on == 0 ? 1 : on == 1 ? 2 : 3;   // AUonAU, AUonCG, AUonGU
        if (gu) return on == 1 ? 4 : on == 0 ? 5 : 6;   // GUonCG, GUonAU, GUonGU
        return 0;
    };
    int n = 0;
    for (int x = 0; x <= 2; x++)
        for (int y = x + 1; y <= 3; y++) {
            if (!is_complement(x, y)) continue;
            for (int w = 0; w <= 3; w++)
                for (int z = 0; z <= 3; z++) {
                    const bool up = is_complement(w, z);
                    const double upH = up ? 0.5 * (h->NN_dH[nn_index(w, x, z, y)] - h->dH0) : 0.0;
                    const double upS = up ? 0.5 * (h->NN_dS[nn_index(w, x, z, y)] - h->dS0) : 0.0;
                    for (int u = 0; u <= 3; u++)
                        for (int v = 0; v <= 3; v++) {
                            const bool lo = is_complement(u, v);
                            if (!up && !lo) continue;
                            ++n;
                            h->bp3_hash[bp3_hash_key(w, x, u, z, y, v)] = n;
                            double dH = upH, dS = upS;
                            if (lo) {
                                dH = upH + 0.5 * (h->NN_dH[nn_index(x, u, y, v)] - h->dH0);
                                dS = upS + 0.5 * (h->NN_dS[nn_index(x, u, y, v)] - h->dS0);
                            }
                            dS *= 1.0e-3;
                            if (h->flg_NNend && up != lo) {
                                int t = up ? end_type(x, y, w, z) : end_type(x, y, u, v);
                                if (t) {
                                    dH += h->NNend_dH[t] - h->dHend0;
                                    dS += (h->NNend_dS[t] - h->dSend0) * 1.0e-3;
                                }
                            }
                            h->bp3_dH[n] = dH;
                            h->bp3_dS[n] = dS;
                        }
                }
        }
    h->bp3_n = n;
    if (n != 468) return set_err("bp3 table has %d entries, expected 468", n);
    return 0;
}

static int alloc_lists(sisgpu *h, int K, int KB)
{
    const size_t R = h->nrep, N = h->nmp;
    if (K != h->K) {
        h->K = K;
        TRY(h->d_nbr.alloc(R * N * K));
    }
    if (KB != h->KB) {
        h->KB = KB;
        TRY(h->d_bprow.alloc(R * N * KB));
        TRY(h->d_bp_type.alloc(R * N * KB));
        TRY(h->d_bp_coef.alloc(R * N * KB));
        TRY(h->d_bp_status.alloc(R * N * KB));
        TRY(h->d_bp_status_MC.alloc(R * N * KB));
        TRY(h->d_ene_bp.alloc(R * N * KB));
        TRY(h->d_bp_half.alloc(R * N * KB));
        TRY(h->d_bp_status_save.alloc(R * N * KB));
        TRY(h->d_ene_bp_save.alloc(R * N * KB));
    }
    invalidate_graphs(h);
    DevSys &S = h->S;
    S.K = h->K; S.KB = h->KB;
    S.bp_half = h->d_bp_half.p; S.bp_half_cap = (int)(N * h->KB);
    S.nbr = h->d_nbr.p; S.bprow = h->d_bprow.p; S.bp_type = h->d_bp_type.p; S.bp_coef = h->d_bp_coef.p;
    S.bp_status = h->d_bp_status.p; S.bp_status_MC = h->d_bp_status_MC.p; S.ene_bp = h->d_ene_bp.p;
    return 0;
}

static int alloc_cells(sisgpu *h, int ncell)
{
    if (ncell <= h->S.ncell_cap) return 0;
    const size_t R = h->nrep;
    int cap = std::max(ncell, 64);
    TRY(h->d_cell_count.alloc(R * cap));
    TRY(h->d_cell_start.alloc(R * (cap + 1)));
    TRY(h->d_cell_cursor.alloc(R * cap));
    h->S.ncell_cap = cap;
    h->S.cell_count = h->d_cell_count.p; h->S.cell_start = h->d_cell_start.p; h->S.cell_cursor = h->d_cell_cursor.p;
    return 0;
}

// md_coef / md_coef_rep of src/job_md.F90:94-108,700-767; noise0 * sqrt(T) = md_coef_rep(1)
static void build_coef(const sisgpu *h, std::vector<double4> &coef)
{
    const double PI = 4.0 * atan(1.0), radius = 10.0;
    coef.resize(h->nmp);
    for (int i = 0; i < h->nmp; i++) {
        double4 c;
        if (h->P.integrator == 1) {
            double v = h->viscosity_Pas * sqrt(1.0e3 / SIS_KCAL2JOUL) * SIS_N_AVO * 1.0e-20;
            double fric = 6.0 * PI * v * radius;
            double c1 = 0.5 * h->dt * fric / h->mass[i];
            double c2 = sqrt(1.0 / (1.0 + c1));
            c.x = (1.0 - c1) / (1.0 + c1);
            c.y = c2 * h->dt / h->mass[i];
            c.z = h->is_frozen[i] ? 0.0 : c2 * h->dt;
            c.w = 0.5 * c2 / h->mass[i] * sqrt(2.0 * fric * SIS_BOLTZ_KCAL_MOL * h->dt);
        } else {
            double v = (SIS_JOUL2KCAL_MOL * 1.0e-18) * h->viscosity_Pas;
            double fric = 6.0 * PI * v * radius;
            c.x = h->is_frozen[i] ? 0.0 : h->dt / fric;
            c.y = c.z = 0.0;
            c.w = h->is_frozen[i] ? 0.0 : sqrt(2.0 * SIS_BOLTZ_KCAL_MOL * h->dt / fric);
        }
        coef[i] = c;
    }
}

// host half of finalize_setup: everything init_bp / list_local / set_md_coef leave behind
// (src/init_bp.F90:39-172,319-346; src/job_md.F90:700-767), no device work
static int host_prepare(sisgpu *h, std::vector<int4> &meta, std::vector<double4> &coef)
{
    if (!h->nmp) return set_err("topology not set");
    if ((int)h->rep.size() != h->nrep) h->rep.assign(h->nrep, HostRep());
    const int N = h->nmp;
    FFParams &P = h->P;
    if (h->bp_model == 5) {
        if (!h->have_nn) return set_err("[NN] required in FF file (bp model 5)");
        TRY(build_bp3_table(h));
    } else {
        h->bp3_hash.assign(4096, 0); h->bp3_dH.assign(469, 0.0); h->bp3_dS.assign(469, 0.0);
    }
    // bp cutoffs: src/init_bp.F90:319-346
    if (fabs(P.bp_cutoff_energy) <= SIS_SMALL_VALUE) {
        h->bp_cutoff_dist = 18.0;
        for (int t = 1; t <= 6; t++) P.bp[t].cutoff_ddist = h->bp_cutoff_dist - 13.8;
    } else {
        h->bp_cutoff_dist = 0.0;
        for (int t = 1; t <= 6; t++) {
            P.bp[t].cutoff_ddist = sqrt(log(fabs(10.0 / P.bp_cutoff_energy)) / P.bp[t].bond_k);
            h->bp_cutoff_dist = std::max(h->bp_cutoff_dist, P.bp[t].bond_r + P.bp[t].cutoff_ddist);
        }
    }
    P.bp[0] = P.bp[1];
    meta.resize(N);
    for (int i = 0; i < N; i++) {
        int c = h->ichain[i];
        int fl = h->seq[i] & MF_SEQ_MASK;
        if (h->has_charge[i]) fl |= MF_CHARGE;
        if (h->moltypes[c] == MOLT_CIRC) fl |= MF_CIRC;
        if (h->is_frozen[i]) fl |= MF_FROZEN;
        meta[i] = make_int4(h->chain_start[c], h->nmp_chain[c], h->lmp[i], fl);
    }
    build_coef(h, coef);
    return 0;
}

int sisgpu_finalize_setup(sisgpu_t *h)
{
    std::vector<int4> meta;
    std::vector<double4> coef;
    TRY(host_prepare(h, meta, coef));
    CK(cudaSetDevice(h->device));
    const int N = h->nmp, R = h->nrep;
    TRY(h->d_meta.upload(meta));
    TRY(h->d_seq.upload(h->seq));
    snap_free(h->snap[0]); snap_free(h->snap[1]);
    h->local_ncand = -1;
    if (N <= FAST_GROUPS) {   // bound of the bp candidates k_md_fast can meet in its replica-local mode
        int cnt = 0, t = 0;
        for (int i = 0; i < N; i++)
            for (int j = i + 1; j < N; j++)
                cnt += bp_lookup_r(host_restrict(h), h->bp_model, h->bp_min_loop, h->bp3_hash.data(), h->seq.data(), i, meta[i], j, meta[j], &t) > 0;
        h->local_ncand = cnt;
    }
    TRY(h->d_coef.upload(coef));
    TRY(h->d_mass.upload(h->mass));
    TRY(h->d_hash.upload(h->bp3_hash));
    TRY(h->d_bp3_dH.upload(h->bp3_dH));
    TRY(h->d_bp3_dS.upload(h->bp3_dS));
    TRY(h->d_twz_pairs.upload(h->twz_pairs)); TRY(h->d_twz_dir.upload(h->twz_dir));
    TRY(h->d_twzfr_pairs.upload(h->twzfr_pairs)); TRY(h->d_twzfr_k.upload(h->twzfr_k));
    TRY(h->d_twzfr_init.upload(h->twzfr_init)); TRY(h->d_twzfr_velo.upload(h->twzfr_velo));
    TRY(h->d_rest_id.upload(h->rest_id)); TRY(h->d_rest_rcut.upload(h->rest_rcut)); TRY(h->d_rest_para.upload(h->rest_para));
    // state
    const size_t RN = (size_t)R * N;
    TRY(h->d_pos0.alloc(RN)); TRY(h->d_pos1.alloc(RN));
    TRY(h->d_vel.alloc(RN * 3)); TRY(h->d_acc.alloc(RN * 3)); TRY(h->d_move.alloc(RN * 3)); TRY(h->d_force.alloc(RN * 3));
    TRY(h->d_nnbr.alloc(RN)); TRY(h->d_nbp_row.alloc(RN)); TRY(h->d_nt_excess.alloc(RN));
    TRY(h->d_cell_of.alloc(RN)); TRY(h->d_cell_beads.alloc(RN)); TRY(h->d_cell_sorted.alloc(RN)); TRY(h->d_grid.alloc(1)); TRY(h->d_xyz_stage.alloc((size_t)N * 3));
    h->grid_host_valid = false;
    TRY(h->d_mc_counter.alloc(R)); TRY(h->d_rep_active.alloc(R));
    TRY(h->d_rp.alloc(R)); TRY(h->d_rp_alt.alloc(R));
    h->sync_bytes = sizeof(DevSyncHdr) + (size_t)R * 4 * sizeof(unsigned long long);
    TRY(h->d_sync.alloc(h->sync_bytes));
    TRY(h->d_kin.alloc(R));
    TRY(h->d_bar.alloc(BAR_WORDS));   // [0] bp / last-CTA counter, [32] position barrier
    TRY(h->d_top2_slots.alloc(2 * FAST_MAX_CTA));
    TRY(h->d_cta_item0.alloc(FAST_MAX_CTA + 1));
    const int nblk = (N + 63) / 64;            // integrate kernel: 64 beads per block
    const int nblk_ene = (N + 31) / 32;        // energy kernel: 8 lanes per bead
    TRY(h->d_ene_partial.alloc((size_t)R * nblk_ene * ENE_N)); TRY(h->d_ene_counter.alloc(R)); TRY(h->d_ene_out.alloc((size_t)R * ENE_N));
    TRY(h->d_f_pair.alloc(RN * 3)); TRY(h->d_f_loc.alloc(RN * 3)); TRY(h->d_f_bp.alloc(RN * 3));
    TRY(h->d_nbp_half.alloc(R));
    TRY(h->d_gather_send.alloc(2 * (size_t)R)); TRY(h->d_gather_recv.alloc(2 * (size_t)std::max(h->nrep_all, R)));
    if (h->h_sync) cudaFreeHost(h->h_sync);
    CK(cudaMallocHost(&h->h_sync, h->sync_bytes));
    memset(h->h_sync, 0, h->sync_bytes);
    h->h_ctl = &((DevSyncHdr *)h->h_sync)->ctl[0];
    h->h_misc = &((DevSyncHdr *)h->h_sync)->misc;
    h->h_top2 = (unsigned long long *)(h->h_sync + sizeof(DevSyncHdr));
    h->h_need.assign(R, 0);
    DevSys &S = h->S;
    S.nmp = N; S.nrep = R; S.nblk = nblk; S.nblk_ene = nblk_ene;
    S.f_pair = h->d_f_pair.p; S.f_loc = h->d_f_loc.p; S.f_bp = h->d_f_bp.p; S.nbp_half = h->d_nbp_half.p;
    S.meta = h->d_meta.p; S.seq = h->d_seq.p; S.coef = h->d_coef.p; S.mass = h->d_mass.p; S.hash = h->d_hash.p;
    S.bp3_dH = h->d_bp3_dH.p; S.bp3_dS = h->d_bp3_dS.p;
    S.pos[0] = h->d_pos0.p; S.pos[1] = h->d_pos1.p; S.vel = h->d_vel.p; S.acc = h->d_acc.p; S.move = h->d_move.p;
    S.force = h->d_force.p; S.nnbr = h->d_nnbr.p; S.nbp_row = h->d_nbp_row.p; S.nt_excess = h->d_nt_excess.p;
    S.mc_counter = h->d_mc_counter.p; S.cell_of = h->d_cell_of.p; S.cell_beads = h->d_cell_beads.p;
    S.cell_sorted = h->d_cell_sorted.p; S.grid = h->d_grid.p;
    S.rp = h->d_rp.p;
    S.ctl = ((DevSyncHdr *)h->d_sync.p)->ctl; S.misc = &((DevSyncHdr *)h->d_sync.p)->misc;
    S.top2 = (unsigned long long *)(h->d_sync.p + sizeof(DevSyncHdr));
    S.nl_margin = h->nl_margin;
    S.bar = h->d_bar.p;
    S.top2_slots = h->d_top2_slots.p;
    S.cta_item0 = h->d_cta_item0.p;
    S.ene_partial = h->d_ene_partial.p; S.ene_counter = h->d_ene_counter.p;
    S.ene_out = h->d_ene_out.p;
    S.twz_pairs = h->d_twz_pairs.p; S.twz_dir = h->d_twz_dir.p; S.twzfr_pairs = h->d_twzfr_pairs.p;
    S.twzfr_k = h->d_twzfr_k.p; S.twzfr_init = h->d_twzfr_init.p; S.twzfr_velo = h->d_twzfr_velo.p;
    S.rest_id = h->d_rest_id.p; S.rest_rcut = h->d_rest_rcut.p; S.rest_para = h->d_rest_para.p;
    S.ncell_cap = 0;
    h->K = h->KB = 0;
    // systems of at most 64 beads: rows that can hold every partner, so that an in-kernel rebuild
    // (k_md_fast, replica-local mode) can never overflow
    // K: rows of the C4 system reach ~130 entries once the coils have relaxed; 192 keeps the x1.2 growth
    // (a reallocation of every list) out of ordinary runs
    TRY(alloc_lists(h, 192, N <= FAST_GROUPS ? std::max(16, (N + 7) / 8 * 8) : 16));
    TRY(alloc_cells(h, 4096));
    h->finalized = true;
    TRY(upload_restrict(h));
    TRY(push_rp(h));
    return 0;
}

// ---- state --------------------------------------------------------------------------------
// one D2H copy of the step bookkeeping; h_need[r] = replica r is past the list margin
// (the test of src/job_md.F90:388 on the displacement top-2 of the last completed step)
static int sync_ctl(sisgpu *h)
{
    CK(cudaMemcpyAsync(h->h_sync, h->d_sync.p, h->sync_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const unsigned long long slot = h->h_ctl->step & 1ull;
    for (int r = 0; r < h->nrep; r++) {
        double m1, m2;
        memcpy(&m1, &h->h_top2[((size_t)r * 2 + slot) * 2], 8);
        memcpy(&m2, &h->h_top2[((size_t)r * 2 + slot) * 2 + 1], 8);
        h->h_need[r] = (sqrt(m1) + sqrt(m2) > h->S.nl_margin) ? 1 : 0;
    }
    return 0;
}

// xyz / velos / accels as the Fortran modules hold them ([nmp][3] per replica).  Copies are enqueued
// on the compute stream (pinned caller buffers make them asynchronous) and the positions are
// (un)packed on the device through a [nmp][3] scratch; one synchronisation per call.
int sisgpu_set_state(sisgpu_t *h, int irep, const double *xyz, const double *velos, const double *accels)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    if (irep < 0 || irep >= h->nrep) return set_err("irep out of range");
    CK(cudaSetDevice(h->device));
    const int N = h->nmp;
    const size_t bytes = (size_t)N * 3 * sizeof(double);
    if (xyz) {
        CK(cudaMemcpyAsync(h->d_xyz_stage.p, xyz, bytes, cudaMemcpyHostToDevice, h->stream));
        k_pos_unpack<<<(N + 255) / 256, 256, 0, h->stream>>>(h->S, h->d_xyz_stage.p, irep);
        h->launches++;
    }
    if (velos) CK(cudaMemcpyAsync(h->S.vel + (size_t)irep * N * 3, velos, bytes, cudaMemcpyHostToDevice, h->stream));
    if (accels) CK(cudaMemcpyAsync(h->S.acc + (size_t)irep * N * 3, accels, bytes, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // the caller may reuse its buffers
    return 0;
}

int sisgpu_get_state(sisgpu_t *h, int irep, double *xyz, double *velos, double *accels)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    if (irep < 0 || irep >= h->nrep) return set_err("irep out of range");
    CK(cudaSetDevice(h->device));
    const int N = h->nmp;
    const size_t bytes = (size_t)N * 3 * sizeof(double);
    if (xyz) {
        k_pos_pack<<<(N + 255) / 256, 256, 0, h->stream>>>(h->S, h->d_xyz_stage.p, irep);
        h->launches++;
        CK(cudaMemcpyAsync(xyz, h->d_xyz_stage.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    }
    if (velos) CK(cudaMemcpyAsync(velos, h->S.vel + (size_t)irep * N * 3, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (accels) CK(cudaMemcpyAsync(accels, h->S.acc + (size_t)irep * N * 3, bytes, cudaMemcpyDeviceToHost, h->stream));
    TRY(sync_ctl(h));
    return 0;
}

int sisgpu_set_step(sisgpu_t *h, int64_t istep)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    CK(cudaSetDevice(h->device));
    TRY(sync_ctl(h));
    h->h_ctl->step = (unsigned long long)istep;
    CK(cudaMemcpy(h->d_sync.p, h->h_sync, sizeof(DevCtl), cudaMemcpyHostToDevice));
    // the displacement slots are indexed by step parity: recompute the slot of the new step from xyz_move
    // (a pending rebuild trigger, src/job_md.F90:376-388, must survive) and clear the other one
    k_top2_from_move<<<h->nrep, 256, 0, h->stream>>>(h->S, (unsigned long long)istep);
    h->launches++;
    h->istep_host = istep;
    TRY(sync_ctl(h));
    CK(cudaGetLastError());
    return 0;
}
int sisgpu_get_step(sisgpu_t *h, int64_t *istep)
{
    TRY(sync_ctl(h));
    *istep = (int64_t)h->h_ctl->step;
    return 0;
}

// State of the counter-based generators for the .rst file (the reference stores 643 int32 of
// Mersenne-Twister state per stream under the PRNG / PRNGREP tags, src/write_rst.F90:69-131).  A
// counter-based stream has no hidden state: what a restart needs is the seed, the step counter (the
// Gaussian counter of every bead), the draw counter of the Monte-Carlo base-pair pruning of every local
// replica and the draw counter of the exchange stream.  state[0] = seed, [1] = istep,
// [2] = exchange counter (PRNGREP), [3] = nrep_proc, [4 .. 4 + nrep_proc) = Monte-Carlo counters (PRNG).
int sisgpu_get_rng_state(sisgpu_t *h, int64_t *state, int cap, int *n)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    const int need = 4 + h->nrep;
    if (n) *n = need;
    if (!state) return 0;
    if (cap < need) return set_err("rng state needs %d words, caller holds %d", need, cap);
    CK(cudaSetDevice(h->device));
    TRY(sync_ctl(h));
    std::vector<unsigned long long> mc(h->nrep);
    CK(cudaMemcpy(mc.data(), h->d_mc_counter.p, h->nrep * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    state[0] = (int64_t)h->seed; state[1] = (int64_t)h->h_ctl->step; state[2] = (int64_t)h->exchange_counter;
    state[3] = h->nrep;
    for (int r = 0; r < h->nrep; r++) state[4 + r] = (int64_t)mc[r];
    return 0;
}
int sisgpu_set_rng_state(sisgpu_t *h, const int64_t *state, int n)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    if (n < 4 || state[3] != h->nrep || n < 4 + h->nrep)
        return set_err("rng state of %lld replicas does not fit nrep_proc = %d", n >= 4 ? (long long)state[3] : -1ll, h->nrep);
    CK(cudaSetDevice(h->device));
    h->seed = (unsigned long long)state[0];
    invalidate_graphs(h);     // the seed is a kernel argument of the replayed graphs
    TRY(sisgpu_set_step(h, state[1]));
    h->exchange_counter = (unsigned long long)state[2];
    std::vector<unsigned long long> mc(h->nrep);
    for (int r = 0; r < h->nrep; r++) mc[r] = (unsigned long long)state[4 + r];
    CK(cudaMemcpy(h->d_mc_counter.p, mc.data(), h->nrep * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    return 0;
}
// REPLICA block of the .rst (src/write_rst.F90:41-60: nrep_all, (i, rep2lab(i))): put the label
// permutation back and refresh the parameters of the local replicas (src/read_rst.F90:405-460)
int sisgpu_remd_set_labels(sisgpu_t *h, const int *rep2lab, int idx_pair);

// ---- neighbour lists ----------------------------------------------------------------------
static void timer_begin(sisgpu *h) { if (h->timing) cudaEventRecord(h->ev0, h->stream); }
static void timer_end(sisgpu *h, double &acc)
{
    if (!h->timing) return;
    cudaEventRecord(h->ev1, h->stream);
    cudaEventSynchronize(h->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    acc += ms;
}

// device-side interval timing without a synchronisation: span_begin / span_end bracket enqueued
// work, spans_resolve adds the elapsed times to their accumulators after the next sync
static int span_begin(sisgpu *h)
{
    if (!h->timing) return -1;
    sisgpu::Pending p{};
    for (cudaEvent_t *e : {&p.a, &p.b}) {
        if (!h->ev_free.empty()) { *e = h->ev_free.back(); h->ev_free.pop_back(); }
        else cudaEventCreate(e);
    }
    p.acc = nullptr;
    cudaEventRecord(p.a, h->stream);
    h->ev_pending.push_back(p);
    return (int)h->ev_pending.size() - 1;
}
static void span_end(sisgpu *h, int id, double *acc)
{
    if (id < 0) return;
    h->ev_pending[id].acc = acc;
    cudaEventRecord(h->ev_pending[id].b, h->stream);
}
static void spans_resolve(sisgpu *h)   // call after a stream synchronisation
{
    static const bool dbg = getenv("SIS_SPAN_DEBUG") != nullptr;
    for (auto &p : h->ev_pending) {
        float ms = 0;
        if (p.acc && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) *p.acc += ms;
        if (dbg && p.acc && ms > 0.3f) fprintf(stderr, "[dev %d] step %lld span %s %.3f ms\n", h->device, (long long)h->h_ctl->step, p.acc == &h->t_nl ? "nl" : p.acc == &h->t_step ? "step" : p.acc == &h->t_rex ? "rex" : "?", ms);
        h->ev_free.push_back(p.a); h->ev_free.push_back(p.b);
    }
    h->ev_pending.clear();
}

static NlArgs nl_args(const sisgpu *h, bool all_pairs)
{
    NlArgs A{};
    A.nl_margin = h->nl_margin;
    A.wca_nl_cut2 = (h->P.wca_sigma + h->nl_margin) * (h->P.wca_sigma + h->nl_margin);
    A.bp_nl_cut2 = h->P.flg_bias_ss ? 9999999999.9 : (h->bp_cutoff_dist + h->nl_margin) * (h->bp_cutoff_dist + h->nl_margin);
    A.excl_bond = h->excl_bond; A.excl_angle = h->excl_angle;
    A.bp_model = h->bp_model; A.bp_min_loop = h->bp_min_loop;
    A.all_pairs = all_pairs;
    A.ele_cutoff_type = h->P.flg_ele ? h->ele_cutoff_type : 0;
    A.ele_cutoff_inp = h->ele_cutoff_inp;
    return A;
}

// bp half-list counts of the last build -> host (sizes the bp task grid of the stream / graph modes)
static int refresh_half_counts(sisgpu *h)
{
    const int R = h->nrep;
    std::vector<int> nh(R);
    CK(cudaMemcpyAsync(nh.data(), h->d_nbp_half.p, R * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    int mx = 0;
    for (int r = 0; r < R; r++) mx = std::max(mx, nh[r]);
    h->nbp_half_rep = nh;
    // the bp task grid is sized from this count; regrow (and re-capture) with slack
    if (mx > h->nbp_half_max || mx < h->nbp_half_max / 4) { h->nbp_half_max = mx + mx / 4 + 128; invalidate_graphs(h); }
    return 0;
}

// grid of the persistent kernels: one CTA per SM
static int ensure_coop_blocks(sisgpu *h)
{
    if (h->coop_blocks) return 0;
    int per_sm = 0, dev_sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_md_fused, FUSED_THREADS, 0));
    CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->device));
    if (per_sm < 1) return set_err("k_md_fused does not fit on an SM");
    h->coop_blocks = dev_sms;
    return 0;
}
// number of CTA bead ranges of k_md_fast (0: the system does not fit: more than 64 beads per SM)
static int fast_ncta_of(sisgpu *h)
{
    if (ensure_coop_blocks(h)) return 0;
    const long long nitems = (long long)h->nrep * h->nmp;
    const int ncta_max = std::min(std::min(h->coop_blocks, FAST_MAX_CTA), PART_MAX_ITEMS / FAST_OWN);
    // up to FAST_OWN beads per CTA; beyond ~96 on average the bead ranges can no longer be cut by pair work
    if (nitems > (long long)ncta_max * (FAST_OWN * 3 / 4)) return 0;
    // small systems: at least ~6 beads per CTA -- the grid barrier gets slower with the number of CTAs
    // (measured: one CAG47 chain 8.6 -> 7.2 us per step with 24 instead of 141 CTAs, four chains 9.2 -> 8.2)
    int ncta = (int)std::min<long long>(ncta_max, (nitems + 5) / 6);
    if (h->fast_max_cta > 0) ncta = std::min(ncta, h->fast_max_cta);
    const int ncta_min = (int)((nitems + FAST_OWN * 3 / 4 - 1) / (FAST_OWN * 3 / 4));
    return std::max(1, std::max(ncta, ncta_min));
}

static int rep_threads_of(sisgpu *h, int nlab, bool xch);
// replica-local mode of k_md_fast: one CTA owns one whole replica and rebuilds its lists itself
static bool fast_local_ok(sisgpu *h)
{
    if (!h->fast_kernel || !h->local_kernel || ensure_coop_blocks(h)) return false;
    if (h->nmp > FAST_GROUPS || h->nrep > FAST_MAX_CTA) return false;
    if (h->nrep > h->coop_blocks && !rep_threads_of(h, 0, false)) return false;   // beyond one replica per SM: k_md_rep only
    // the replica-local kernel carries dual constant-force tweezers (F-REMD); the other biases go through k_md_big / per-phase kernels
    if (h->P.flg_stage || h->P.ntwz_FR || h->P.nrest || h->P.flg_bias_ss || h->P.flg_bias_rg) return false;
    if (h->K < h->nmp || h->KB < h->nmp - 1) return false;
    return h->local_ncand >= 0 && h->local_ncand <= FAST_BPC;
}

// k_md_rep: the replica-local mode with a CTA sized to the replica (sis_rep.cuh).  Returns the threads per
// replica (128 / 256) or 0 when the kernel cannot take the run; *nlab_out = labels of the in-kernel exchange tables
static size_t rep_smem_bytes(sisgpu *h, int nlab)
{
    return rep_carve(nullptr, h->nmp, std::max(1, h->local_ncand), nlab).bytes;
}
static void *rep_kernel_fn(int nt, bool xch)
{
    if (nt == 512) return xch ? (void *)k_md_rep<512, true> : (void *)k_md_rep<512, false>;
    if (nt == 256) return xch ? (void *)k_md_rep<256, true> : (void *)k_md_rep<256, false>;
    return xch ? (void *)k_md_rep<128, true> : (void *)k_md_rep<128, false>;
}
// replicas the GPU holds co-resident with `nt` threads each (the exchange barrier needs every replica's CTA resident)
static int rep_capacity(sisgpu *h, int nt, int nlab, bool xch)
{
    const size_t smem = rep_smem_bytes(h, nlab);
    void *fn = rep_kernel_fn(nt, xch);
    const int slot = (nt == 512 ? 4 : nt == 256 ? 2 : 0) + (xch ? 1 : 0);
    if (smem > h->rep_smem_set[slot]) {
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        h->rep_smem_set[slot] = smem;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, nt, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return per_sm * h->coop_blocks;
}
static int rep_threads_of(sisgpu *h, int nlab, bool xch)
{
    if (!h->rep_kernel || !h->fast_kernel || !h->local_kernel || ensure_coop_blocks(h)) return 0;
    if (h->nmp > REP_MAXB || h->nmp < 2 || h->local_ncand < 0 || h->local_ncand > REP_MAXC) return 0;
    if (h->P.flg_stage || h->P.ntwz_FR || h->P.nrest || h->P.flg_bias_ss || h->P.flg_bias_rg) return 0;
    if (h->K < h->nmp || h->KB < h->nmp - 1) return 0;
    if (rep_smem_bytes(h, nlab) > 200 * 1024) return 0;
    // the widest CTA the replica count leaves room for: a replica's step is a latency chain, more warps shorten it
    const int want = h->rep_threads == 128 || h->rep_threads == 256 || h->rep_threads == 512 ? h->rep_threads : 0;
    for (int nt = 512; nt >= 128; nt >>= 1) {
        if (want && nt != want) continue;
        if (h->nrep <= rep_capacity(h, nt, nlab, xch)) return nt;
    }
    // more replicas than the GPU holds at once: fine in waves as long as no exchange barrier runs inside the kernel
    if (!xch && rep_capacity(h, 128, nlab, false) > 0) return want ? want : 128;
    return 0;
}

// cell grid over [lo, hi] with cells of at least rc
static void make_grid(const double lo[3], const double hi[3], double rc, int pbc, GridInfo &Gd)
{
    long long ncell = 1;
    Gd.pbc = pbc;
    for (int d = 0; d < 3; d++) {
        double ext = std::max(hi[d] - lo[d], 1e-9);
        int n = (int)floor(ext / (rc * 1.000001));
        n = std::max(1, std::min(n, 128));
        Gd.n[d] = n;
        Gd.origin[d] = lo[d];
        Gd.inv_cell[d] = n / ext;
        ncell *= n;
    }
    Gd.ncell = (int)ncell;
}
static int upload_grid(sisgpu *h, const GridInfo &Gd)
{
    if (h->grid_host_valid && memcmp(&h->grid_host, &Gd, sizeof Gd) == 0) return 0;
    h->grid_host = Gd;
    h->grid_host_valid = true;
    CK(cudaMemcpyAsync(h->d_grid.p, &h->grid_host, sizeof Gd, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

// the kernels of one list build after the replica flags and the grid are on the device
static void enqueue_nl_kernels(sisgpu *h, const NlArgs &A, bool grid_pass, bool wrap_only)
{
    const int N = h->nmp, R = h->nrep;
    DevSys &S = h->S;
    const int *act = h->d_rep_active.p;
    if (grid_pass) {
        k_nl_wrap_count<<<dim3((N + 255) / 256, R), 256, 0, h->stream>>>(S, h->P, act);
        k_nl_scan<<<R, 1024, 0, h->stream>>>(S, act);
        k_nl_fill<<<dim3((N + 255) / 256, R), 256, 0, h->stream>>>(S, act);
        k_nl_rank<<<dim3((N + 255) / 256, R), 256, 0, h->stream>>>(S, act);
        h->launches += 4;
    } else if (wrap_only) {
        k_nl_wrap_count<<<dim3((N + 255) / 256, R), 256, 0, h->stream>>>(S, h->P, act);
        h->launches += 1;
    }
    k_nl_rows<<<dim3((N + 7) / 8, R), 256, 0, h->stream>>>(S, h->P, A, act);
    h->launches += 1;
}
static void enqueue_nl_tail(sisgpu *h)
{
    const int N = h->nmp, R = h->nrep;
    k_nl_bp_slots<<<dim3((N * h->KB + 255) / 256, R), 256, 0, h->stream>>>(h->S, h->d_rep_active.p);
    const bool local = fast_local_ok(h);
    const int ncta = local ? h->nrep : (h->fast_kernel ? fast_ncta_of(h) : 0);
    const size_t part_smem = ncta > 0 ? (size_t)h->nrep * h->nmp * sizeof(int) : 0;
    if (part_smem > h->part_smem_set) {
        if (cudaFuncSetAttribute((void *)k_nl_partition, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem) == cudaSuccess)
            h->part_smem_set = part_smem;
    }
    k_nl_partition<<<1, 1024, part_smem, h->stream>>>(h->S, h->d_rep_active.p, ncta, FAST_OWN, local ? 1 : 0);
    h->launches += 2;
}

// builds the rows of the replicas flagged in `active`; all_pairs = list_exv/list_ele/list_bp mode
static int build_lists(sisgpu *h, const std::vector<int> &active, bool all_pairs)
{
    CK(cudaSetDevice(h->device));
    const int N = h->nmp, R = h->nrep;
    DevSys &S = h->S;
    timer_begin(h);
    // src/neighbor_list.F90:80-82: ele_cutoff refreshed here for cutoff_type 2
    double rc = std::max(h->P.wca_sigma, h->P.flg_bias_ss ? 0.0 : h->bp_cutoff_dist);
    bool rp_dirty = false;
    for (int r = 0; r < R; r++) {
        if (!active[r]) continue;
        if (h->P.flg_ele) {
            if (h->ele_cutoff_type == 2) {
                double c = h->ele_cutoff_inp * h->rep[r].lambdaD;
                if (c != h->rep[r].ele_cutoff) { h->rep[r].ele_cutoff = c; rp_dirty = true; }
            }
            rc = std::max(rc, h->rep[r].ele_cutoff);
        }
    }
    if (rp_dirty) TRY(push_rp(h));
    rc += h->nl_margin;
    TRY(sync_ctl(h));
    const int cur = h->h_ctl->cur;
    GridInfo Gd{};
    Gd.pbc = h->P.flg_pbc;
    const bool brute = all_pairs || h->P.flg_bias_ss;   // Bias_SS: bp cutoff is infinite
    if (!brute) {
        double lo[3], hi[3];
        if (h->P.flg_pbc) {
            for (int d = 0; d < 3; d++) { lo[d] = 0.0; hi[d] = h->P.box[d]; }
        } else {
            // bounding box of the active replicas (host reduction; this path is the explicit
            // sisgpu_neighbor_list call -- rebuilds inside md_steps use k_nl_plan)
            std::vector<double4> p(N);
            for (int d = 0; d < 3; d++) { lo[d] = 1e300; hi[d] = -1e300; }
            for (int r = 0; r < R; r++) {
                if (!active[r]) continue;
                CK(cudaMemcpy(p.data(), S.pos[cur] + (size_t)r * N, N * sizeof(double4), cudaMemcpyDeviceToHost));
                for (int i = 0; i < N; i++) {
                    lo[0] = std::min(lo[0], p[i].x); hi[0] = std::max(hi[0], p[i].x);
                    lo[1] = std::min(lo[1], p[i].y); hi[1] = std::max(hi[1], p[i].y);
                    lo[2] = std::min(lo[2], p[i].z); hi[2] = std::max(hi[2], p[i].z);
                }
            }
        }
        make_grid(lo, hi, rc, h->P.flg_pbc, Gd);
        TRY(alloc_cells(h, Gd.ncell));
    } else {
        Gd.n[0] = Gd.n[1] = Gd.n[2] = 1; Gd.ncell = 1;
    }
    TRY(upload_grid(h, Gd));
    CK(cudaMemcpyAsync(h->d_rep_active.p, active.data(), R * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    NlArgs A = nl_args(h, brute);
    A.all_pairs = all_pairs ? 1 : (brute ? 2 : 0);
    for (int attempt = 0; attempt < 8; attempt++) {
        h->h_misc->nl_overflow = 0; h->h_misc->bp_overflow = 0;
        CK(cudaMemcpyAsync(&h->S.misc->nl_overflow, &h->h_misc->nl_overflow, 2 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        // Bias_SS with PBC still wraps first (src/neighbor_list.F90:28-30): wrap kernel on the 1-cell grid
        enqueue_nl_kernels(h, A, !brute && attempt == 0, brute && h->P.flg_pbc && !all_pairs && attempt == 0);
        TRY(sync_ctl(h));
        CK(cudaGetLastError());
        if (h->h_misc->nl_overflow > h->K || h->h_misc->bp_overflow > h->KB) {
            int K = h->K, KB = h->KB;
            if (h->h_misc->nl_overflow > K) K = ((int)(h->h_misc->nl_overflow * 1.2) + 31) / 32 * 32;   // growth as the reference's x1.2
            if (h->h_misc->bp_overflow > KB) KB = ((int)(h->h_misc->bp_overflow * 1.2) + 7) / 8 * 8;
            // a capacity change relocates every replica's rows: rebuild all of them
            std::vector<int> all(R, 1);
            bool was_all = true;
            for (int r = 0; r < R; r++) was_all &= (active[r] != 0);
            TRY(alloc_lists(h, K, KB));
            if (!was_all) return build_lists(h, all, all_pairs);
            continue;
        }
        k_nl_begin_slots<<<(R + 63) / 64, 64, 0, h->stream>>>(S, h->d_rep_active.p);
        h->launches += 1;
        enqueue_nl_tail(h);
        TRY(refresh_half_counts(h));
        CK(cudaGetLastError());
        timer_end(h, h->t_nl);
        h->lists_full = all_pairs;
        h->lists_built = true;
        const double margin = all_pairs ? 1e300 : h->nl_margin;
        if (margin != h->S.nl_margin) { h->S.nl_margin = margin; invalidate_graphs(h); }
        return 0;
    }
    return set_err("neighbour list capacity did not converge");
}

// The device-driven rebuild of the MD loop: k_nl_plan decides on the device which replicas are
// past the list margin (src/job_md.F90:388), the chain rebuilds exactly those, and k_nl_partition
// re-cuts the CTA bead ranges of k_md_fast.  Nothing is read back; a chain with no replica to
// rebuild costs nine empty launches.
static int enqueue_nl_chain(sisgpu *h, int force_all, long long stop)
{
    const int R = h->nrep;
    NlPlanArgs PA{};
    double rc = std::max(h->P.wca_sigma, h->bp_cutoff_dist);
    if (h->P.flg_ele)
        for (int r = 0; r < R; r++) {
            rc = std::max(rc, h->rep[r].ele_cutoff);
            if (h->ele_cutoff_type == 2) rc = std::max(rc, h->ele_cutoff_inp * h->rep[r].lambdaD);
        }
    if (h->flg_replica) rc = std::max(rc, h->remd_cut_max);   // labels met during an enqueued batch
    PA.rc = rc + h->nl_margin;
    PA.ele_cutoff_inp = h->ele_cutoff_inp; PA.ele_cutoff_type = h->ele_cutoff_type; PA.flg_ele = h->P.flg_ele;
    PA.pbc = h->P.flg_pbc; PA.force_all = force_all; PA.stop = stop;   // stop < 0: S.misc->stop_step
    if (h->P.flg_pbc) {
        double lo[3] = {0.0, 0.0, 0.0}, hi[3] = {h->P.box[0], h->P.box[1], h->P.box[2]};
        GridInfo Gd{};
        make_grid(lo, hi, PA.rc, 1, Gd);
        TRY(alloc_cells(h, Gd.ncell));
        TRY(upload_grid(h, Gd));
    } else {
        h->grid_host_valid = false;   // k_nl_plan writes the grid
    }
    k_nl_plan<<<1, 1024, 0, h->stream>>>(h->S, PA, h->d_rep_active.p);
    h->launches += 1;
    enqueue_nl_kernels(h, nl_args(h, false), true, false);
    enqueue_nl_tail(h);
    return 0;
}

int sisgpu_neighbor_list(sisgpu_t *h, int irep)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    if (irep < 0 || irep >= h->nrep) return set_err("irep out of range");
    std::vector<int> act(h->nrep, 0);
    act[irep] = 1;
    h->nrebuild++;
    return build_lists(h, act, false);
}

int sisgpu_full_lists(sisgpu_t *h)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    std::vector<int> act(h->nrep, 1);
    // src/main.F90:150-152 path; ele_cutoff for cutoff_type 2 as set by init_ele
    return build_lists(h, act, true);
}

// half lists in the reference's loop order (ichain, jchain, i, j): src/neighbor_list.F90:99-118
struct HalfLists {
    std::vector<std::pair<int, int>> wca, ele;
    std::vector<int> bp_i, bp_j, bp_ibp;
};
static int fetch_half_lists(sisgpu *h, int irep, HalfLists &L)
{
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    const int N = h->nmp, K = h->K, KB = h->KB;
    std::vector<uint32_t> nbr((size_t)N * K);
    std::vector<int> nn(N), nb(N);
    std::vector<int2> bprow((size_t)N * KB);
    CK(cudaMemcpy(nbr.data(), h->S.nbr + (size_t)irep * N * K, nbr.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(nn.data(), h->S.nnbr + (size_t)irep * N, N * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(nb.data(), h->S.nbp_row + (size_t)irep * N, N * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(bprow.data(), h->S.bprow + (size_t)irep * N * KB, bprow.size() * sizeof(int2), cudaMemcpyDeviceToHost));
    auto key_less = [h](const std::pair<int, int> &a, const std::pair<int, int> &b) {
        int ac = h->ichain[a.first], bc = h->ichain[b.first];
        if (ac != bc) return ac < bc;
        int ajc = h->ichain[a.second], bjc = h->ichain[b.second];
        if (ajc != bjc) return ajc < bjc;
        if (a.first != b.first) return a.first < b.first;
        return a.second < b.second;
    };
    for (int i = 0; i < N; i++)
        for (int e = 0; e < nn[i]; e++) {
            uint32_t ent = nbr[(size_t)i * K + e];
            int j = ent & NB_IDX;
            if (j <= i) continue;
            if (ent & NB_WCA) L.wca.emplace_back(i, j);
            if (ent & NB_ELE) L.ele.emplace_back(i, j);
        }
    std::sort(L.wca.begin(), L.wca.end(), key_less);
    std::sort(L.ele.begin(), L.ele.end(), key_less);
    std::vector<std::pair<std::pair<int, int>, int>> bp;
    for (int i = 0; i < N; i++)
        for (int e = 0; e < nb[i]; e++) {
            int2 ent = bprow[(size_t)i * KB + e];
            if (ent.x > i) bp.push_back({{i, ent.x}, ent.y});
        }
    // src/list_bp.F90:55-57 walks (imp, jmp) in global order; neighbor_list walks chain pairs
    std::sort(bp.begin(), bp.end(), [&](const std::pair<std::pair<int, int>, int> &a, const std::pair<std::pair<int, int>, int> &b) {
        return h->lists_full ? a.first < b.first : key_less(a.first, b.first);
    });
    for (auto &b : bp) { L.bp_i.push_back(b.first.first); L.bp_j.push_back(b.first.second); L.bp_ibp.push_back(b.second); }
    return 0;
}

int sisgpu_get_list_sizes(sisgpu_t *h, int irep, int *nwca, int *nele, int *nbp)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    HalfLists L;
    TRY(fetch_half_lists(h, irep, L));
    *nwca = (int)L.wca.size(); *nele = (int)L.ele.size(); *nbp = (int)L.bp_i.size();
    return 0;
}

static int host_chain_nb(const sisgpu *h, int imp, int off)
{
    int c = h->ichain[imp], l = h->lmp[imp] + off, n = h->nmp_chain[c];
    if (h->moltypes[c] == MOLT_CIRC) { l %= n; if (l < 0) l += n; return h->chain_start[c] + l; }
    // linear chain ends never enter the bp list; mirror the reference's index arithmetic anyway
    return imp + off;
}

int sisgpu_get_lists(sisgpu_t *h, int irep, int *wca_mp, int *ele_mp, int *bp_mp, double *bp_coef)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    HalfLists L;
    TRY(fetch_half_lists(h, irep, L));
    if (wca_mp) for (size_t k = 0; k < L.wca.size(); k++) { wca_mp[2 * k] = L.wca[k].first + 1; wca_mp[2 * k + 1] = L.wca[k].second + 1; }
    if (ele_mp) for (size_t k = 0; k < L.ele.size(); k++) { ele_mp[2 * k] = L.ele[k].first + 1; ele_mp[2 * k + 1] = L.ele[k].second + 1; }
    const size_t nb = L.bp_i.size();
    if (nb && (bp_mp || bp_coef)) {
        const size_t NKB = (size_t)h->nmp * h->KB;
        std::vector<signed char> type(NKB);
        std::vector<double2> coef(NKB);
        CK(cudaMemcpy(type.data(), h->S.bp_type + irep * NKB, NKB, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(coef.data(), h->S.bp_coef + irep * NKB, NKB * sizeof(double2), cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < nb; k++) {
            int i = L.bp_i[k], j = L.bp_j[k], ibp = L.bp_ibp[k];
            if (bp_mp) {
                int *b = bp_mp + 7 * k;
                b[0] = i + 1; b[1] = j + 1;
                b[2] = host_chain_nb(h, i, -1) + 1; b[3] = host_chain_nb(h, j, -1) + 1;
                b[4] = host_chain_nb(h, i, +1) + 1; b[5] = host_chain_nb(h, j, +1) + 1;
                b[6] = type[ibp];
            }
            if (bp_coef) { bp_coef[2 * k] = coef[ibp].x; bp_coef[2 * k + 1] = coef[ibp].y; }
        }
    }
    return 0;
}

// list_local (src/list_local.F90:16-119) is implicit on the device (consecutive beads per chain
// + circRNA closures); this getter materialises it for parity checks and the Fortran shim
int sisgpu_get_local_lists(sisgpu_t *h, int *nbond, int *nangl, int *ndih, int *bond_mp, int *angl_mp, int *dih_mp)
{
    std::vector<int> bd, an, di;
    const bool has_dih = h->P.flg_dih_cos || h->P.flg_dih_exp;
    for (int c = 0; c < h->nchains; c++) {
        const int n = h->nmp_chain[c], s = h->chain_start[c];
        const bool circ = h->moltypes[c] == MOLT_CIRC;
        auto at = [&](int l) { return s + ((l % n) + n) % n + 1; };
        for (int i = 0; i + 1 < n; i++) { bd.push_back(at(i)); bd.push_back(at(i + 1)); }
        for (int i = 0; i + 2 < n; i++) for (int q = 0; q < 3; q++) an.push_back(at(i + q));
        if (has_dih) for (int i = 0; i + 3 < n; i++) for (int q = 0; q < 4; q++) di.push_back(at(i + q));
        if (circ) {
            bd.push_back(at(n - 1)); bd.push_back(at(n));
            for (int a = n - 2; a <= n - 1; a++) for (int q = 0; q < 3; q++) an.push_back(at(a + q));
            if (has_dih) for (int a = n - 3; a <= n - 1; a++) for (int q = 0; q < 4; q++) di.push_back(at(a + q));
        }
    }
    *nbond = (int)bd.size() / 2; *nangl = (int)an.size() / 3; *ndih = (int)di.size() / 4;
    if (bond_mp) std::copy(bd.begin(), bd.end(), bond_mp);
    if (angl_mp) std::copy(an.begin(), an.end(), angl_mp);
    if (dih_mp) std::copy(di.begin(), di.end(), dih_mp);
    return 0;
}

// write_bpcoef (src/write_bpcoef.F90:30-83) without the N x N bp3_map: every sequence-allowed pair imp < jmp in the
// reference's loop order with its (dH, dS * 1e-3); the driver applies coef_dG * (dH - T dS) < 0 per replica and
// formats the line.  imp / jmp 1-based; cap = capacity of the arrays; *n = pairs found (call with cap = 0 to size).
// Host-side enumeration: meant for the system sizes whose .bpcoef file is usable (it has O(N^2) lines).
int sisgpu_bpcoef(sisgpu_t *h, int cap, int *n, int *imp, int *jmp, double *dH, double *dS)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    if (h->bp_model != 5) { *n = 0; return 0; }       // the reference writes the table for model 5 only (:24)
    std::vector<int4> meta;
    std::vector<double4> coef;
    // meta as finalize_setup built it
    meta.resize(h->nmp);
    for (int i = 0; i < h->nmp; i++) {
        const int c = h->ichain[i];
        int fl = h->seq[i] & MF_SEQ_MASK;
        if (h->has_charge[i]) fl |= MF_CHARGE;
        if (h->moltypes[c] == MOLT_CIRC) fl |= MF_CIRC;
        if (h->is_frozen[i]) fl |= MF_FROZEN;
        meta[i] = make_int4(h->chain_start[c], h->nmp_chain[c], h->lmp[i], fl);
    }
    const BpRestrict rs = host_restrict(h);
    long long cnt = 0;
    for (int i = 0; i < h->nmp - 1; i++)
        for (int j = i + 1; j < h->nmp; j++) {
            int t = 0;
            const int b3 = bp_lookup_r(rs, h->bp_model, h->bp_min_loop, h->bp3_hash.data(), h->seq.data(), i, meta[i], j, meta[j], &t);
            if (b3 <= 0) continue;
            if (cnt < cap) {
                if (imp) imp[cnt] = i + 1;
                if (jmp) jmp[cnt] = j + 1;
                if (dH) dH[cnt] = h->bp3_dH[b3];
                if (dS) dS[cnt] = h->bp3_dS[b3];
            }
            cnt++;
        }
    if (cnt > 0x7fffffffll) return set_err("bpcoef: more than 2^31 pairs");
    *n = (int)cnt;
    return 0;
}

int sisgpu_bp3_table(sisgpu_t *h, int *n, int *hash4096, double *dH, double *dS)
{
    if (h->bp3_hash.empty()) return set_err("finalize_setup not called");
    *n = h->bp3_n;
    if (hash4096) std::copy(h->bp3_hash.begin(), h->bp3_hash.end(), hash4096);
    if (dH) std::copy(h->bp3_dH.begin() + 1, h->bp3_dH.end(), dH);
    if (dS) std::copy(h->bp3_dS.begin() + 1, h->bp3_dS.end(), dS);
    return 0;
}

// ---- hot path -----------------------------------------------------------------------------
int sisgpu_set_bp_mc(sisgpu_t *h, int flg) { h->flg_bp_MC = flg; return 0; }

static StepArgs step_args(const sisgpu *h, unsigned mask, int use_mask, int only_rep, int integrate)
{
    StepArgs A{};
    A.seed = h->seed;
    A.term_mask = mask; A.use_mask = use_mask; A.only_rep = only_rep; A.integrate = integrate;
    A.nblk_pair = (h->nmp + 256 / h->G - 1) / (256 / h->G);
    A.nblk_loc = (h->nmp + 31) / 32;
    A.nblk_bp = (std::max(h->nbp_half_max, 1) + 255) / 256;
    A.max_bp_per_nt = h->max_bp_per_nt;
    return A;
}

// one force evaluation (+ integration): the pair/local task kernel on `stream`, the bp task
// kernel concurrently on `stream2` (fork/join through events; inside a stream capture these
// become the two parallel branches of the step graph), then the integrate kernel
static void enqueue_step(sisgpu *h, const StepArgs &A)
{
    const int R = A.only_rep >= 0 ? 1 : h->nrep;
    const bool do_bp = A.term_mask & (1u << ENE_BP);
    if (do_bp) {
        cudaEventRecord(h->ev_fork, h->stream);
        cudaStreamWaitEvent(h->stream2, h->ev_fork, 0);
        k_force_bp<<<dim3((std::max(h->nbp_half_max, 1) + 127) / 128, R), 128, 0, h->stream2>>>(h->S, h->P, A);
        cudaEventRecord(h->ev_join, h->stream2);
        h->launches++;
    }
    dim3 grid(A.nblk_pair + A.nblk_loc, R);
    switch (h->G) {
    case 4: k_force_local<4><<<grid, 256, 0, h->stream>>>(h->S, h->P, A); break;
    case 16: k_force_local<16><<<grid, 256, 0, h->stream>>>(h->S, h->P, A); break;
    case 32: k_force_local<32><<<grid, 256, 0, h->stream>>>(h->S, h->P, A); break;
    default: k_force_local<8><<<grid, 256, 0, h->stream>>>(h->S, h->P, A); break;
    }
    if (do_bp) cudaStreamWaitEvent(h->stream, h->ev_join, 0);
    k_integrate<<<dim3(h->S.nblk, R), 256, 0, h->stream>>>(h->S, h->P, A);
    h->launches += 2;
}

// evaluation + MC pruning passes of force_bp_limit_triplet / energy_bp_limit_triplet
template <int SIDE>
static void launch_bp_eval(sisgpu *h, int only_rep, bool with_prune, double tK, int use_override, int use_mask,
                           int copy_to_mc, int deterministic, int in_md = 0)
{
    const int R = only_rep >= 0 ? 1 : h->nrep;
    if (with_prune) {
        k_excess_init<<<dim3((h->nmp + 255) / 256, R), 256, 0, h->stream>>>(h->S, only_rep, h->max_bp_per_nt, in_md);
        h->launches++;
    }
    k_bp_eval<SIDE><<<dim3((h->nmp * h->KB + 255) / 256, R), 256, 0, h->stream>>>(h->S, h->P, only_rep, tK, use_override,
                                                                                use_mask, with_prune ? 1 : 0, in_md);
    h->launches++;
    if (with_prune) {
        k_bp_prune<<<R, 32, 0, h->stream>>>(h->S, h->P, h->seed, only_rep, tK, use_override, deterministic, copy_to_mc, in_md);
        h->launches++;
    }
}

static void launch_rg(sisgpu *h, int only_rep)
{
    if (!h->P.flg_bias_rg) return;
    k_rg<<<only_rep >= 0 ? 1 : h->nrep, 256, 0, h->stream>>>(h->S, only_rep);
    h->launches++;
}

int sisgpu_force_terms(sisgpu_t *h, int irep, unsigned mask, double *forces)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    if (irep < 0 || irep >= h->nrep) return set_err("irep out of range");
    CK(cudaSetDevice(h->device));
    int use_mask = h->nstep_bp_MC > 0 ? 1 : 0;
    if ((mask & (1u << ENE_BP)) && h->flg_bp_MC) {
        // src/force.F90:44-49 -> force_bp_limit_triplet: evaluate all, prune, scatter survivors
        launch_bp_eval<0>(h, irep, true, 0.0, 0, 0, 0, 0);
        use_mask = 2;
    }
    if (mask & (1u << ENE_RG)) launch_rg(h, irep);
    enqueue_step(h, step_args(h, mask, use_mask, irep, 0));
    CK(cudaMemcpyAsync(forces, h->S.force + (size_t)irep * h->nmp * 3, 3 * h->nmp * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return 0;
}
int sisgpu_force(sisgpu_t *h, int irep, double *forces) { return sisgpu_force_terms(h, irep, 0x7FEu, forces); }

static int energy_launch(sisgpu *h, int only_rep, const RepParams *rp_alt)
{
    dim3 grid(h->S.nblk_ene, only_rep >= 0 ? 1 : h->nrep);
    k_energy<8><<<grid, 256, 0, h->stream>>>(h->S, h->P, rp_alt, only_rep);
    h->launches++;
    return 0;
}

int sisgpu_energy_sumup(sisgpu_t *h, int irep, double tempK, int recompute_bp, double *energies)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    if (irep < 0 || irep >= h->nrep) return set_err("irep out of range");
    CK(cudaSetDevice(h->device));
    timer_begin(h);
    if (recompute_bp) {
        // energy_bp / energy_bp_triplet / energy_bp_limit_triplet with flg_bp_energy = .False.
        const bool limit = h->flg_bp_MC != 0;
        launch_bp_eval<1>(h, irep, limit, tempK, 1, (!limit && h->nstep_bp_MC > 0) ? 1 : 0, 0, h->P.temp_independent);
    }
    launch_rg(h, irep);
    TRY(energy_launch(h, irep, nullptr));
    CK(cudaMemcpyAsync(energies, h->S.ene_out + (size_t)irep * ENE_N, ENE_N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TRY(sync_ctl(h));
    CK(cudaGetLastError());
    timer_end(h, h->t_ene);
    if (h->h_misc->max_force_flag & 2) return set_err("Error: Z-coordinate < 0.85 * stage_sigma in energy_stage. Abort!");
    return 0;
}

// job_dcd (src/job_dcd.F90:65-79): energy_sumup with flg_bp_energy = .False. for every frame of a
// trajectory, nrep_proc frames per launch (the replica slots hold frames; all slots must carry the
// same parameters).  xyz32: float32 [nframes][nmp][3] as read from the DCD (src/dcd.F90:584-592),
// energies: [nframes][11].  The replica coordinates are overwritten.
int sisgpu_dcd_energies(sisgpu_t *h, int nframes, const float *xyz32, double tempK, double *energies)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    if (!h->lists_built || !h->lists_full)
        return set_err("dcd_energies needs the cutoff-free lists (sisgpu_full_lists, as job = DCD does: src/main.F90:148-154)");
    CK(cudaSetDevice(h->device));
    const int R = h->nrep, N = h->nmp;
    const size_t fbytes = (size_t)N * 3 * sizeof(float);
    float *d_xyz[2] = {nullptr, nullptr}, *h_xyz[2] = {nullptr, nullptr};
    double *h_ene[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    int rc = 0;
    auto cleanup = [&]() {
        for (int b = 0; b < 2; b++) {
            if (d_xyz[b]) cudaFree(d_xyz[b]);
            if (h_xyz[b]) cudaFreeHost(h_xyz[b]);
            if (h_ene[b]) cudaFreeHost(h_ene[b]);
            if (done[b]) cudaEventDestroy(done[b]);
        }
    };
#define DCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = set_err("CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); cleanup(); return rc; } } while (0)
    for (int b = 0; b < 2; b++) {
        DCK(cudaMalloc(&d_xyz[b], fbytes * R));
        DCK(cudaMallocHost(&h_xyz[b], fbytes * R));
        DCK(cudaMallocHost(&h_ene[b], (size_t)R * ENE_N * sizeof(double)));
        DCK(cudaEventCreate(&done[b]));
    }
    TRY(sync_ctl(h));
    timer_begin(h);
    const bool limit = h->flg_bp_MC != 0;
    const int nbatch = (nframes + R - 1) / R;
    // two staging buffers: batch k + 1 is packed on the host while batch k runs; results of batch
    // k - 1 are unpacked after its event
    for (int k = 0; k <= nbatch; k++) {
        const int b = k & 1;
        if (k < nbatch) {
            const int f0 = k * R, nf = std::min(R, nframes - f0);
            if (k >= 2) DCK(cudaEventSynchronize(done[b]));   // the buffer's previous batch has been read back
            if (k >= 2) {
                const int g0 = (k - 2) * R, ng = std::min(R, nframes - g0);
                memcpy(energies + (size_t)g0 * ENE_N, h_ene[b], (size_t)ng * ENE_N * sizeof(double));
            }
            memcpy(h_xyz[b], xyz32 + (size_t)f0 * N * 3, fbytes * nf);
            DCK(cudaMemcpyAsync(d_xyz[b], h_xyz[b], fbytes * nf, cudaMemcpyHostToDevice, h->stream));
            k_dcd_load<<<dim3((N + 127) / 128, nf), 128, 0, h->stream>>>(h->S, d_xyz[b], nf);
            h->launches++;
            launch_bp_eval<1>(h, -1, limit, tempK, 1, (!limit && h->nstep_bp_MC > 0) ? 1 : 0, 0, h->P.temp_independent);
            launch_rg(h, -1);
            TRY(energy_launch(h, -1, nullptr));
            DCK(cudaMemcpyAsync(h_ene[b], h->S.ene_out, (size_t)nf * ENE_N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            DCK(cudaEventRecord(done[b], h->stream));
        }
    }
    for (int k = std::max(0, nbatch - 2); k < nbatch; k++) {
        const int b = k & 1, g0 = k * R, ng = std::min(R, nframes - g0);
        DCK(cudaEventSynchronize(done[b]));
        memcpy(energies + (size_t)g0 * ENE_N, h_ene[b], (size_t)ng * ENE_N * sizeof(double));
    }
    DCK(cudaGetLastError());
#undef DCK
    timer_end(h, h->t_ene);
    cleanup();
    TRY(sync_ctl(h));
    if (h->h_misc->max_force_flag & 2) return set_err("Error: Z-coordinate < 0.85 * stage_sigma in energy_stage. Abort!");
    return 0;
}

int sisgpu_energy_kinetic(sisgpu_t *h, int irep, double *ekin)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    CK(cudaSetDevice(h->device));
    k_kinetic<<<1, 256, 0, h->stream>>>(h->S, h->d_kin.p, irep);
    h->launches++;
    CK(cudaMemcpyAsync(ekin, h->d_kin.p + irep, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// bp_status_MC(:,:) = bp_status(:,:) after the initial energies of a run or a restart (src/job_md.F90:344-348):
// the mask the steps between two Monte-Carlo steps evaluate (src/force_bp_triplet.F90:46-48)
int sisgpu_bp_status_to_mc(sisgpu_t *h)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    CK(cudaSetDevice(h->device));
    k_copy_status_mc<<<dim3((h->nmp * h->KB + 255) / 256, h->nrep), 256, 0, h->stream>>>(h->S, -1);
    h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return 0;
}

int sisgpu_get_bp_state(sisgpu_t *h, int irep, int *status, double *ene)
{
    HalfLists L;
    TRY(fetch_half_lists(h, irep, L));
    const size_t NKB = (size_t)h->nmp * h->KB;
    std::vector<int> st(NKB);
    std::vector<double> en(NKB);
    CK(cudaMemcpy(st.data(), h->S.bp_status + irep * NKB, NKB * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(en.data(), h->S.ene_bp + irep * NKB, NKB * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < L.bp_ibp.size(); k++) {
        if (status) status[k] = st[L.bp_ibp[k]];
        if (ene) ene[k] = en[L.bp_ibp[k]];
    }
    return 0;
}

int sisgpu_get_rg(sisgpu_t *h, int irep, double *rg)
{
    RepParams R;
    CK(cudaMemcpy(&R, h->d_rp.p + irep, sizeof R, cudaMemcpyDeviceToHost));
    *rg = R.rg_val;
    return 0;
}

int sisgpu_init_velocities(sisgpu_t *h, int irep)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    CK(cudaSetDevice(h->device));
    k_init_velocities<<<(h->nmp * 3 + 255) / 256, 256, 0, h->stream>>>(h->S, h->seed, irep);
    h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return 0;
}

// ---- REMD (host mirror of var_replica / init_replica / replica_exchange) ---------------------
static int get_pair(const sisgpu *h, int label) { return h->pair_tb[(size_t)(h->idx_pair - 1) * h->nrep_all + label - 1]; }

static void label_params(const sisgpu *h, int label, const HostRep &cur, double &T, double &I, double &f, double &g0)
{
    const int n = h->nrep_all, l = label - 1;
    T = h->flg_repvar[REPT_TEMP] ? h->lab2val[REPT_TEMP * n + l] : cur.tempK;
    I = h->flg_repvar[REPT_ION] ? h->lab2val[REPT_ION * n + l] : cur.ionic;
    f = h->flg_repvar[REPT_TWZDCF] ? h->lab2val[REPT_TWZDCF * n + l] : cur.twz_force;
    g0 = h->flg_repvar[REPT_RG] ? h->lab2val[REPT_RG * n + l] : cur.rg0;
}

// parameter refresh of src/job_md.F90:538-567: set_md_coef (T), set_ele, twz forces, Rg0.
// ele_cutoff is NOT refreshed here (reference quirk: only neighbor_list does, for cutoff_type 2).
static void apply_label(sisgpu *h, int irep, int label)
{
    HostRep &r = h->rep[irep];
    double T, I, f, g0;
    label_params(h, label, r, T, I, f, g0);
    r.tempK = T; r.ionic = I; r.twz_force = f; r.rg0 = g0;
    if (h->P.flg_ele) host_set_ele(h, r, T, I);
}

// host part of sisgpu_remd_setup: tables and the parameters of the local replicas
static int remd_setup_host(sisgpu *h, const int nrep[4], const double *values, int nstep_exchange, int rank, int nranks)
{
    h->flg_replica = true; h->ndim_replica = 0; h->nrep_all = 1;
    for (int v = 0; v < REPT_MAX; v++) {
        h->nrep_dim[v] = nrep[v];
        h->flg_repvar[v] = nrep[v] > 1;
        if (nrep[v] > MAX_REP_PER_DIM) return set_err("too many replicas in one dimension");
        for (int k = 0; k < nrep[v]; k++) h->replica_values[v][k] = values[v * MAX_REP_PER_DIM + k];
        if (h->flg_repvar[v]) { h->ndim_replica++; h->nrep_all *= nrep[v]; }
    }
    if (h->ndim_replica > MAX_REP_DIM) return set_err("too many replica dimensions");
    if (nranks > h->nrep_all) return set_err("Error: the number of MPI processes should be equal to or smaller than the number of replicas.");
    if (h->nrep_all % nranks) return set_err("Error: the number of replicas has to be a multiple of the number of MPI processes.");
    if (h->nrep_all / nranks != h->nrep) return set_err("nrep_proc=%d but nrep_all/nranks=%d", h->nrep, h->nrep_all / nranks);
    h->rank = rank; h->nranks = nranks; h->nstep_exchange = nstep_exchange;
    const int n = h->nrep_all;
    h->rep2lab.resize(n); h->lab2rep.resize(n);
    h->lab2val.assign((size_t)REPT_MAX * n, 1.0e31);
    // lab2val: src/init_replica.F90:70-92
    int n_division = 1;
    for (int v = 0; v < REPT_MAX; v++) {
        if (!h->flg_repvar[v]) continue;
        const int n_division_pre = n_division;
        n_division *= h->nrep_dim[v];
        const int n_continue = n / n_division, n_skip = n / n_division_pre;
        for (int ir = 0; ir < h->nrep_dim[v]; ir++)
            for (int ido = 0; ido < n_division_pre; ido++)
                for (int is = 0; is < n_continue; is++)
                    h->lab2val[(size_t)v * n + ir * n_continue + ido * n_skip + is] = h->replica_values[v][ir];
    }
    for (int g = 0; g < n; g++) { h->rep2lab[g] = g + 1; h->lab2rep[g] = g + 1; }
    int idx = 0;
    for (int v = 0; v < REPT_MAX; v++) if (h->flg_repvar[v]) h->type_array[idx++] = v;
    // exchange_pair_tb: src/var_replica.F90:162-259 (odd/even neighbour pairs per dimension)
    const int ncol = 2 * h->ndim_replica;
    h->pair_tb.assign((size_t)ncol * n, 0);
    int col = 0;
    for (int parity = 1; parity <= 2; parity++) {
        n_division = 1;
        for (int d = 0; d < h->ndim_replica; d++, col++) {
            int *tb = h->pair_tb.data() + (size_t)col * n;
            const int nv = h->nrep_dim[h->type_array[d]];
            const int n_division_pre = n_division;
            n_division *= nv;
            const int n_continue = n / n_division;
            int iset = 0;
            for (int cyc = 0; cyc < n_division_pre; cyc++) {
                bool post_zero;
                int start;
                if (parity == 1) { post_zero = (nv % 2 == 1); start = 1; }
                else { iset += n_continue; post_zero = (nv % 2 == 0); start = 2; }   // leading block unpaired
                for (int ir = start; ir <= nv - 1; ir += 2)
                    for (int ic = 1; ic <= 2 * n_continue; ic++) {
                        iset++;
                        tb[iset - 1] = ic <= n_continue ? iset + n_continue : iset - n_continue;
                    }
                if (post_zero) iset += n_continue;
            }
        }
    }
    h->idx_pair = 1;
    for (int irep = 0; irep < h->nrep; irep++) {
        const int grep = rank * h->nrep + irep;
        apply_label(h, irep, h->rep2lab[grep]);
        HostRep &r = h->rep[irep];
        if (h->P.flg_ele) {
            if (h->ele_cutoff_type == 1) r.ele_cutoff = h->ele_cutoff_inp;
            else r.ele_cutoff = h->ele_cutoff_inp * r.lambdaD;
        }
    }
    return 0;
}

int sisgpu_remd_setup(sisgpu_t *h, const int nrep[4], const double *values, int nstep_exchange, int rank, int nranks)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    TRY(remd_setup_host(h, nrep, values, nstep_exchange, rank, nranks));
    TRY(h->d_gather_recv.alloc(2 * (size_t)h->nrep_all));
    return push_rp(h);
}

int sisgpu_nccl_unique_id(unsigned char *id)
{
    TRY(nccl_load());
    NcclId nid;
    int rc = g_nccl.GetUniqueId(&nid);
    if (rc) return set_err("ncclGetUniqueId failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    memcpy(id, nid.internal, SISGPU_NCCL_ID_BYTES);
    return 0;
}

int sisgpu_comm_init(sisgpu_t *h, int nranks, int rank, const unsigned char *id)
{
    TRY(nccl_load());
    CK(cudaSetDevice(h->device));
    NcclId nid;
    memcpy(nid.internal, id, SISGPU_NCCL_ID_BYTES);
    int rc = g_nccl.CommInitRank(&h->nccl_comm, nranks, nid, rank);
    if (rc) return set_err("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return 0;
}

int sisgpu_remd_pair_table(sisgpu_t *h, int *tb)
{
    std::copy(h->pair_tb.begin(), h->pair_tb.end(), tb);
    return 0;
}
int sisgpu_remd_state(sisgpu_t *h, int *rep2lab, int *lab2rep, double *lab2val)
{
    if (rep2lab) std::copy(h->rep2lab.begin(), h->rep2lab.end(), rep2lab);
    if (lab2rep) std::copy(h->lab2rep.begin(), h->lab2rep.end(), lab2rep);
    if (lab2val) std::copy(h->lab2val.begin(), h->lab2val.end(), lab2val);
    return 0;
}
int sisgpu_remd_set_labels(sisgpu_t *h, const int *rep2lab, int idx_pair)
{
    if (!h->flg_replica) return set_err("REMD not configured");
    const int n = h->nrep_all;
    std::vector<int> seen(n, 0);
    for (int g = 0; g < n; g++) {
        if (rep2lab[g] < 1 || rep2lab[g] > n || seen[rep2lab[g] - 1]++) return set_err("rep2lab is not a permutation of 1..%d", n);
    }
    for (int g = 0; g < n; g++) { h->rep2lab[g] = rep2lab[g]; h->lab2rep[rep2lab[g] - 1] = g + 1; }
    if (idx_pair >= 1 && idx_pair <= 2 * h->ndim_replica) h->idx_pair = idx_pair;
    for (int irep = 0; irep < h->nrep; irep++) apply_label(h, irep, h->rep2lab[h->rank * h->nrep + irep]);
    return push_rp(h);
}
int sisgpu_remd_get_idx_pair(sisgpu_t *h, int *idx_pair) { *idx_pair = h->idx_pair; return 0; }
int sisgpu_remd_stats(sisgpu_t *h, int64_t *n_attempt, int64_t *n_accept)
{
    *n_attempt = h->n_attempt; *n_accept = h->n_accept;
    return 0;
}

// The cross-energy pass re-evaluates the base pairs with the partner's parameters into bp_status / ene_bp;
// the reference saves and restores both around it (src/energy_replica.F90:64-70,157-162) so that write_bp /
// write_out of a save step that is also an exchange step see the replica's own values
static int bp_state_save(sisgpu *h, bool restore)
{
    const size_t n = (size_t)h->nrep * h->nmp * h->KB;
    if (!restore) {
        CK(cudaMemcpyAsync(h->d_bp_status_save.p, h->d_bp_status.p, n * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_ene_bp_save.p, h->d_ene_bp.p, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    } else {
        CK(cudaMemcpyAsync(h->d_bp_status.p, h->d_bp_status_save.p, n * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_ene_bp.p, h->d_ene_bp_save.p, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    return 0;
}

// energy_replica on an exchange step (src/energy_replica.F90:45-163): own-parameter energies
// (cached bp energies), then partner-parameter energies with the bp term recomputed, then the
// allgather that replaces MPI_ALLREDUCE (src/job_md.F90:531-533)
int sisgpu_energy_replica(sisgpu_t *h, double *energies, double *replica_energies)
{
    if (!h->flg_replica) return set_err("REMD not configured");
    CK(cudaSetDevice(h->device));
    timer_begin(h);
    const int R = h->nrep, n = h->nrep_all;
    std::vector<double> e_own((size_t)R * ENE_N), e_opp((size_t)R * ENE_N);
    if (h->flg_bp_MC && h->nstep_bp_MC > 0)   // src/job_md.F90:476-479: flg_bp_energy = .False. with MC
        launch_bp_eval<1>(h, -1, true, 0.0, 0, 0, 0, h->P.temp_independent);
    launch_rg(h, -1);
    TRY(energy_launch(h, -1, nullptr));
    CK(cudaMemcpyAsync(e_own.data(), h->S.ene_out, e_own.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    // partner parameter blocks
    std::vector<RepParams> alt(R);
    std::vector<int> has_opp(R, 0);
    for (int irep = 0; irep < R; irep++) {
        const int grep = h->rank * R + irep;
        const int label_own = h->rep2lab[grep], label_opp = get_pair(h, label_own);
        fill_rp(h, irep, alt[irep]);
        if (label_opp == 0 || label_opp > n) continue;
        has_opp[irep] = 1;
        HostRep saved = h->rep[irep];
        apply_label(h, irep, label_opp);
        fill_rp(h, irep, alt[irep]);
        h->rep[irep] = saved;
    }
    CK(cudaMemcpyAsync(h->d_rp_alt.p, alt.data(), R * sizeof(RepParams), cudaMemcpyHostToDevice, h->stream));
    // cross energies: bp term recomputed with the partner temperature (flg_bp_energy = .False., :72)
    {
        // swap in the alternative block for the bp evaluation (tempK per replica comes from rp)
        CK(cudaStreamSynchronize(h->stream));
        const bool limit = h->flg_bp_MC != 0;
        TRY(bp_state_save(h, false));
        RepParams *saved = h->S.rp;
        h->S.rp = h->d_rp_alt.p;
        launch_bp_eval<1>(h, -1, limit, 0.0, 0, (!limit && h->nstep_bp_MC > 0) ? 1 : 0, 0, h->P.temp_independent);
        launch_rg(h, -1);
        const int rc = energy_launch(h, -1, nullptr);
        h->S.rp = saved;                 // restored on every path
        if (rc) return rc;
    }
    CK(cudaMemcpyAsync(e_opp.data(), h->S.ene_out, e_opp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    TRY(bp_state_save(h, true));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    std::vector<double> send(2 * (size_t)R);
    for (int irep = 0; irep < R; irep++) {
        send[2 * irep] = e_own[(size_t)irep * ENE_N];
        send[2 * irep + 1] = has_opp[irep] ? e_opp[(size_t)irep * ENE_N] : 0.0;
    }
    if (h->nranks > 1) {
        if (!h->nccl_comm) return set_err("sisgpu_comm_init not called for a multi-rank REMD");
        CK(cudaMemcpyAsync(h->d_gather_send.p, send.data(), send.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        int rc = g_nccl.AllGather(h->d_gather_send.p, h->d_gather_recv.p, 2 * (size_t)R, NCCL_FLOAT64, h->nccl_comm, h->stream);
        if (rc) return set_err("ncclAllGather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
        CK(cudaMemcpyAsync(replica_energies, h->d_gather_recv.p, 2 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    } else {
        std::copy(send.begin(), send.end(), replica_energies);
    }
    if (energies) std::copy(e_own.begin(), e_own.end(), energies);
    timer_end(h, h->t_rex);
    return 0;
}

// replica_exchange (src/replica_exchange.F90:44-154): identical decisions on every rank.
// Host part: label tables, acceptance, parameter refresh of the local replicas (src/job_md.F90:538-567).
// The velocity rescaling of accepted temperature swaps (scale_velo, :175-213) is returned as
// (local replica, factor) pairs for the caller to apply on the device.
struct VelScale { int irep; double factor; };
static void remd_decide(sisgpu *h, const double *E, const double *u, int *n_attempt, int *n_accept, std::vector<VelScale> &scales)
{
    const int n = h->nrep_all, R = h->nrep;
    int nrand = 0, natt = 0, nacc = 0;
    for (int l_i = 1; l_i <= n; l_i++) {
        const int l_j = get_pair(h, l_i);
        if (l_j < l_i) continue;
        const double rnd = u[nrand++];
        const int rep_i = h->lab2rep[l_i - 1], rep_j = h->lab2rep[l_j - 1];
        double temp_m, temp_n;
        if (h->flg_repvar[REPT_TEMP]) {
            temp_m = h->lab2val[(size_t)REPT_TEMP * n + l_i - 1];
            temp_n = h->lab2val[(size_t)REPT_TEMP * n + l_j - 1];
        } else temp_m = temp_n = h->rep[0].tempK;
        const double pot_i_m = E[2 * (rep_i - 1)], pot_i_n = E[2 * (rep_i - 1) + 1];
        const double pot_j_n = E[2 * (rep_j - 1)], pot_j_m = E[2 * (rep_j - 1) + 1];
        const double delta = ((pot_j_m - pot_i_m) / temp_m - (pot_j_n - pot_i_n) / temp_n) / SIS_BOLTZ_KCAL_MOL;
        natt++;
        if (delta < 0.0 || rnd <= exp(-delta)) {
            nacc++;
            h->rep2lab[rep_i - 1] = l_j; h->rep2lab[rep_j - 1] = l_i;
            h->lab2rep[l_i - 1] = rep_j; h->lab2rep[l_j - 1] = rep_i;
            if (h->flg_repvar[REPT_TEMP]) {   // scale_velo :175-213, only for locally held replicas
                const double sf = sqrt(temp_n / temp_m);
                const int li = rep_i - 1 - h->rank * R, lj = rep_j - 1 - h->rank * R;
                if (li >= 0 && li < R) scales.push_back({li, sf});
                if (lj >= 0 && lj < R) scales.push_back({lj, 1.0 / sf});
            }
        }
    }
    h->idx_pair = (h->idx_pair == 2 * h->ndim_replica) ? 1 : h->idx_pair + 1;   // set_forward
    for (int irep = 0; irep < R; irep++) apply_label(h, irep, h->rep2lab[h->rank * R + irep]);
    h->n_attempt += natt; h->n_accept += nacc;
    if (n_attempt) *n_attempt = natt;
    if (n_accept) *n_accept = nacc;
}

int sisgpu_replica_exchange(sisgpu_t *h, const double *E, const double *uniforms, int *n_attempt, int *n_accept)
{
    if (!h->flg_replica) return set_err("REMD not configured");
    CK(cudaSetDevice(h->device));
    // the reference draws ceil(nrep_all/2) numbers up front from mts_rep (:44-48)
    const int ndraw = (h->nrep_all + 1) / 2;
    std::vector<double> u(ndraw);
    for (int i = 0; i < ndraw; i++)
        u[i] = uniforms ? uniforms[i] : uniform_counter(h->seed, 0x7fffffffu, h->exchange_counter + i);
    if (!uniforms) h->exchange_counter += ndraw;
    std::vector<VelScale> scales;
    remd_decide(h, E, u.data(), n_attempt, n_accept, scales);
    for (const VelScale &v : scales) {
        k_scale_velocities<<<(3 * h->nmp + 255) / 256, 256, 0, h->stream>>>(h->S, v.irep, v.factor);
        h->launches++;
    }
    TRY(push_rp(h));
    return 0;
}

// ---- exchange buffer of the in-kernel exchange and its peer mapping ------------------------------
static int xch_alloc(sisgpu *h)
{
    const size_t n = (size_t)h->nrep_all;
    if (h->d_xch.n != 4 * n + 4) {
        if (h->peers_ready) return set_err("the exchange buffer is already shared with the other ranks");
        TRY(h->d_xch.alloc(4 * n + 4));    // zero-filled: counter 0
        h->xch_rounds = 0;
    }
    return 0;
}
// One CUDA IPC handle per rank (64 bytes) that the driver all-gathers (MPI_ALLGATHER / torch.distributed) and
// hands to sisgpu_peer_import: afterwards the exchange steps of a multi-rank REMD run inside the persistent
// kernels, energies and arrival flags travelling as peer stores / atomics over NVLink instead of through
// ncclAllGather launches.  Ranks of ONE node only; call after sisgpu_remd_setup.
int sisgpu_peer_export(sisgpu_t *h, unsigned char *handle)
{
    if (!h->flg_replica) return set_err("REMD not configured");
    CK(cudaSetDevice(h->device));
    TRY(xch_alloc(h));
    cudaIpcMemHandle_t ih;
    CK(cudaIpcGetMemHandle(&ih, h->d_xch.p));
    static_assert(sizeof ih == SISGPU_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(handle, &ih, sizeof ih);
    return 0;
}
int sisgpu_peer_import(sisgpu_t *h, int nranks, const unsigned char *handles)
{
    if (!h->flg_replica) return set_err("REMD not configured");
    if (nranks != h->nranks) return set_err("peer_import: %d handles for %d ranks", nranks, h->nranks);
    if (nranks > REMD_MAX_PEERS) return set_err("peer_import: at most %d ranks", REMD_MAX_PEERS);
    CK(cudaSetDevice(h->device));
    TRY(xch_alloc(h));
    for (int r = 0; r < nranks; r++) {
        if (r == h->rank) { h->peer_xch[r] = h->d_xch.p; continue; }
        cudaIpcMemHandle_t ih;
        memcpy(&ih, handles + (size_t)r * SISGPU_IPC_HANDLE_BYTES, sizeof ih);
        CK(cudaIpcOpenMemHandle(&h->peer_xch[r], ih, cudaIpcMemLazyEnablePeerAccess));
    }
    h->peers_ready = true;
    return 0;
}

// ---- device-resident exchange step ------------------------------------------------------------
// Host tables -> device (the host copies are authoritative between md_steps calls: the per-call API
// sisgpu_energy_replica / sisgpu_replica_exchange works on them).
static int remd_upload(sisgpu *h)
{
    const int n = h->nrep_all, R = h->nrep;
    if (h->d_rep2lab.n != (size_t)n) {
        TRY(h->d_rep2lab.alloc(n)); TRY(h->d_lab2rep.alloc(n)); TRY(h->d_lab_temp.alloc(n)); TRY(h->d_lab_rp.alloc(n));
        TRY(h->d_has_opp.alloc(R)); TRY(h->d_vscale.alloc(R)); TRY(h->d_remd.alloc(1));
    }
    if (h->d_pair_tb.n != h->pair_tb.size()) TRY(h->d_pair_tb.alloc(h->pair_tb.size()));
    if (h->d_gather_recv.n < 2 * (size_t)n) TRY(h->d_gather_recv.alloc(2 * (size_t)n));
    TRY(xch_alloc(h));
    // parameter block of every label: what apply_label + fill_rp give a replica that holds it
    std::vector<RepParams> lab(n);
    std::vector<double> ltemp(n);
    const HostRep saved = h->rep[0];
    h->remd_cut_max = 0.0;
    for (int l = 1; l <= n; l++) {
        apply_label(h, 0, l);
        fill_rp(h, 0, lab[l - 1]);
        if (h->P.flg_ele && h->ele_cutoff_type == 2) h->remd_cut_max = std::max(h->remd_cut_max, h->ele_cutoff_inp * h->rep[0].lambdaD);
        ltemp[l - 1] = h->flg_repvar[REPT_TEMP] ? h->lab2val[(size_t)REPT_TEMP * n + l - 1] : saved.tempK;
        h->rep[0] = saved;
    }
    RemdDev D{};
    D.n = n; D.ncol = 2 * h->ndim_replica; D.idx_pair = h->idx_pair; D.R = R; D.rank = h->rank;
    D.flg_temp = h->flg_repvar[REPT_TEMP];
    D.flg_ele_var = h->P.flg_ele && (h->flg_repvar[REPT_TEMP] || h->flg_repvar[REPT_ION]);
    D.flg_twz = h->flg_repvar[REPT_TWZDCF]; D.flg_rg = h->flg_repvar[REPT_RG];
    D.seed = h->seed; D.exchange_counter = h->exchange_counter;
    D.n_attempt = h->n_attempt; D.n_accept = h->n_accept;
    D.tempK_default = h->rep[0].tempK;
    D.pair_tb = h->d_pair_tb.p; D.rep2lab = h->d_rep2lab.p; D.lab2rep = h->d_lab2rep.p;
    D.lab_temp = h->d_lab_temp.p; D.lab_rp = h->d_lab_rp.p; D.rp_alt = h->d_rp_alt.p; D.has_opp = h->d_has_opp.p;
    D.send = h->d_gather_send.p; D.recv = h->nranks > 1 ? h->d_gather_recv.p : h->d_gather_send.p; D.vscale = h->d_vscale.p;
    D.nex = h->nstep_exchange; D.xch_round0 = h->xch_rounds;
    D.nranks_x = (h->nranks > 1 && h->peers_ready) ? h->nranks : 1;
    for (int r = 0; r < D.nranks_x; r++) {
        void *base = D.nranks_x == 1 ? (void *)h->d_xch.p : h->peer_xch[r];
        D.xE_peer[D.nranks_x == 1 ? h->rank : r] = (double *)base;
        D.xbar_peer[D.nranks_x == 1 ? h->rank : r] = (unsigned long long *)((double *)base + 4 * (size_t)n);
    }
    CK(cudaMemcpyAsync(h->d_pair_tb.p, h->pair_tb.data(), h->pair_tb.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_rep2lab.p, h->rep2lab.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_lab2rep.p, h->lab2rep.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_lab_temp.p, ltemp.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_lab_rp.p, lab.data(), n * sizeof(RepParams), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_remd.p, &D, sizeof D, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // the sources are locals
    return 0;
}
// device tables -> host mirrors after the last enqueued exchange (the stream is synchronised)
static int remd_download(sisgpu *h)
{
    const int n = h->nrep_all, R = h->nrep;
    RemdDev D;
    CK(cudaMemcpy(&D, h->d_remd.p, sizeof D, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h->rep2lab.data(), h->d_rep2lab.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h->lab2rep.data(), h->d_lab2rep.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    h->idx_pair = D.idx_pair; h->exchange_counter = D.exchange_counter;
    h->n_attempt = D.n_attempt; h->n_accept = D.n_accept;
    h->xch_rounds = D.xch_round0;
    for (int irep = 0; irep < R; irep++) apply_label(h, irep, h->rep2lab[h->rank * R + irep]);
    return 0;
}
// one exchange step, enqueued: energy_replica (own + partner parameters), the all-gather that
// replaces MPI_ALLREDUCE (src/job_md.F90:531-533), replica_exchange and the parameter refresh
static int enqueue_exchange(sisgpu *h)
{
    const int R = h->nrep;
    const int sp = span_begin(h);
    // with the Monte-Carlo base-pair limit the exchange step evaluates with flg_bp_MC set and
    // flg_bp_energy = .False. (src/job_md.F90:476-479): same launches as sisgpu_energy_replica
    const bool limit = h->nstep_bp_MC > 0 || h->flg_bp_MC != 0;
    if (limit && h->nstep_bp_MC > 0) launch_bp_eval<1>(h, -1, true, 0.0, 0, 0, 0, h->P.temp_independent);
    launch_rg(h, -1);
    TRY(energy_launch(h, -1, nullptr));
    k_remd_prepare<<<(R + 63) / 64, 64, 0, h->stream>>>(h->S, h->d_remd.p);
    {
        // cross energies: bp term recomputed with the partner temperature (flg_bp_energy = .False., :72)
        TRY(bp_state_save(h, false));
        RepParams *saved = h->S.rp;
        h->S.rp = h->d_rp_alt.p;
        launch_bp_eval<1>(h, -1, limit, 0.0, 0, (!limit && h->nstep_bp_MC > 0) ? 1 : 0, 0, h->P.temp_independent);
        launch_rg(h, -1);
        const int rc = energy_launch(h, -1, nullptr);
        h->S.rp = saved;                 // restored on every path
        if (rc) return rc;
        TRY(bp_state_save(h, true));     // k_energy has consumed the partner-parameter values
    }
    k_remd_pack<<<(R + 63) / 64, 64, 0, h->stream>>>(h->S, h->d_remd.p);
    if (h->nranks > 1) {
        if (!h->nccl_comm) return set_err("sisgpu_comm_init not called for a multi-rank REMD");
        int rc = g_nccl.AllGather(h->d_gather_send.p, h->d_gather_recv.p, 2 * (size_t)R, NCCL_FLOAT64, h->nccl_comm, h->stream);
        if (rc) return set_err("ncclAllGather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    }   // a single rank reads its own send buffer (RemdDev::recv points at it, remd_upload)
    const int scale_here = (long long)R * h->nmp * 3 <= 32768 ? 1 : 0;
    k_remd_exchange<<<1, 256, 0, h->stream>>>(h->S, h->d_remd.p, scale_here);
    if (!scale_here) k_remd_scale<<<dim3((3 * h->nmp + 255) / 256, R), 256, 0, h->stream>>>(h->S, h->d_remd.p);
    h->launches += scale_here ? 3 : 4;
    span_end(h, sp, &h->t_rex);
    return 0;
}

// which persistent kernel sisgpu_md_steps launches: k_md_fast in its replica-local mode, k_md_fast
// with a grid barrier, or the general k_md_fused
static size_t fast_smem_bytes()
{
    return ((sizeof(FastShared) + 15) & ~size_t(15)) + (size_t)3 * FAST_EMAX * sizeof(uint32_t);   // rows, sort scratch / unit forces, unit tables
}
static void persistent_choice(sisgpu *h, bool *local, bool *fast_ok, int *want_fast)
{
    const int R = h->nrep;
    *local = fast_local_ok(h);
    *want_fast = *local ? R : (h->fast_kernel ? fast_ncta_of(h) : 0);
    long long ncand_total = 0;
    for (int r = 0; r < R; r++) ncand_total += h->nbp_half_rep.empty() ? h->nbp_half_max : h->nbp_half_rep[r];
    const bool biases = h->P.flg_stage || (h->P.ntwz_DCF && !*local) || h->P.ntwz_FR || h->P.nrest;
    *fast_ok = *want_fast > 0 && !biases && h->K >= 24 && h->K <= 512 && fast_smem_bytes() <= 200 * 1024 &&
               (*local || 3 * (ncand_total / *want_fast) + 16 <= FAST_BPC);
    if (!*fast_ok) *local = false;
}

// launch of one persistent kernel that runs until step `stop` or until a replica needs new lists
static int launch_persistent(sisgpu *h, long long stop)
{
    const int R = h->nrep;
    TRY(ensure_coop_blocks(h));
    StepArgs A = step_args(h, 0x7FEu, 0, -1, 1);
    const long long nitems = (long long)R * h->nmp;
    CK(cudaMemsetAsync(h->d_bar.p, 0, BAR_WORDS * sizeof(unsigned), h->stream));
    NlArgs NA = nl_args(h, false);
    void *args[] = {&h->S, &h->P, &A, &stop, &NA};
    // k_md_fast: rows / candidates / bonded tables on chip
    bool local, fast_ok;
    int want_fast;
    persistent_choice(h, &local, &fast_ok, &want_fast);
    const size_t smem_fast = fast_smem_bytes();
    if (h->nstep_bp_MC > 0 && !(fast_ok && local))
        return set_err("internal: the Monte-Carlo base-pair limit needs the replica-local kernel");
    h->last_launch_local = local;
    h->last_step_kernel = fast_ok ? (local ? "k_md_fast<1,true>" : "k_md_fast<1,false>") : "k_md_fused";
    const bool xch = h->S.remd != nullptr;
    const int nlab = xch ? h->nrep_all : 0;
    const int rep_nt = (fast_ok && local) ? rep_threads_of(h, nlab, xch) : 0;
    if (fast_ok && local && !rep_nt && R > h->coop_blocks) return set_err("internal: %d replicas need k_md_rep", R);
    if (rep_nt) {
        void *fn = rep_kernel_fn(rep_nt, xch);
        int ccap = std::max(1, h->local_ncand), nl = nlab;
        const size_t smem = rep_smem_bytes(h, nlab);
        void *rargs[] = {&h->S, &h->P, &A, &stop, &NA, &ccap, &nl};
        h->last_step_kernel = rep_nt == 128 ? "k_md_rep<128>" : rep_nt == 256 ? "k_md_rep<256>" : "k_md_rep<512>";
        if (xch || R <= rep_capacity(h, rep_nt, nlab, xch)) CK(cudaLaunchCooperativeKernel(fn, dim3(R), dim3(rep_nt), rargs, smem, h->stream));
        else CK(cudaLaunchKernel(fn, dim3(R), dim3(rep_nt), rargs, smem, h->stream));      // in waves (no barrier between the CTAs)
    } else if (fast_ok) {
        void *fn = local ? (h->S.remd ? (void *)k_md_fast<1, true, true> : (void *)k_md_fast<1, true, false>) : (void *)k_md_fast<1, false, false>;
        if (smem_fast > h->fast_smem_set || fn != h->fast_fn_set) {
            CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast));
            h->fast_smem_set = smem_fast; h->fast_fn_set = fn;
        }
        CK(cudaLaunchCooperativeKernel(fn, dim3(want_fast), dim3(FAST_THREADS), args, smem_fast, h->stream));
    } else {
        // a persistent kernel that stopped at the rebuild test may have left the base-pair
        // scatter of the step it did not execute in f_bp
        CK(cudaMemsetAsync(h->d_f_bp.p, 0, (size_t)R * h->nmp * 3 * sizeof(double), h->stream));
        const int want = (int)std::max<long long>(1, (nitems + FUSED_GROUPS - 1) / FUSED_GROUPS);
        dim3 grid(std::min(h->coop_blocks, want));
        CK(cudaLaunchCooperativeKernel((void *)k_md_fused, grid, dim3(FUSED_THREADS), args, 0, h->stream));
    }
    h->launches++;
    return 0;
}

// ---- k_md_big: one launch per MD step, replayed as graph chunks (large systems) ----------------------
// beads per CTA: rows of K entries must fit the quad map, and the grid should fill whole waves of
// 3 CTAs per SM
static int big_bpc_of(sisgpu *h)
{
    if (ensure_coop_blocks(h)) return 0;
    const int per_quads = (h->K + 3) / 4;
    int bpc = std::min(BIG_BPC, BIG_QMAP / std::max(1, per_quads));
    if (bpc < 4) return 0;
    const long long nitems = (long long)h->nrep * h->nmp;
    const long long slots = (long long)BIG_MINB * h->coop_blocks;
    const long long waves = (nitems + slots * bpc - 1) / (slots * bpc);
    const int even = (int)((nitems + slots * waves - 1) / (slots * waves));
    return std::max(4, std::min(bpc, even));
}
static bool big_ok(sisgpu *h)
{
    return h->big_kernel && h->nstep_bp_MC == 0 && !h->P.flg_bias_rg && !h->P.flg_bias_ss && big_bpc_of(h) > 0;
}
// `big_chunk` step kernels as one graph; launch idx reads S.ctl[idx & 1] and publishes S.ctl[(idx + 1) & 1],
// so an even chunk leaves the current bookkeeping in S.ctl[0] where every other kernel reads it
static int launch_big_chunk(sisgpu *h)
{
    const int bpc = big_bpc_of(h);
    const int G = std::max(2, h->big_chunk & ~1);
    const size_t smem = sizeof(BigShared);
    if (!h->big_smem_set) {
        CK(cudaFuncSetAttribute((void *)k_md_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->big_smem_set = true;
    }
    if (!h->graph_big || h->graph_big_steps != G || h->graph_big_bpc != bpc) {
        if (h->graph_big) { cudaGraphExecDestroy(h->graph_big); h->graph_big = nullptr; }
        const StepArgs A = step_args(h, 0x7FEu, 0, -1, 1);
        const long long nitems = (long long)h->nrep * h->nmp;
        const int ncta = (int)((nitems + bpc - 1) / bpc);
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < G; i++) k_md_big<<<ncta, BIG_THREADS, smem, h->stream>>>(h->S, h->P, A, i, bpc);
        CK(cudaStreamEndCapture(h->stream, &graph));
        CK(cudaGraphInstantiate(&h->graph_big, graph, 0));
        CK(cudaGraphDestroy(graph));
        h->graph_big_steps = G; h->graph_big_bpc = bpc;
    }
    CK(cudaGraphLaunch(h->graph_big, h->stream));
    h->launches += G;
    h->last_launch_local = false;
    h->last_step_kernel = "k_md_big";
    return 0;
}

// post-sync bookkeeping shared by the persistent paths
static int after_persistent_sync(sisgpu *h)
{
    h->nrebuild += h->h_misc->nrebuild - h->dev_nrebuild_seen;
    h->dev_nrebuild_seen = h->h_misc->nrebuild;
    if (h->h_misc->max_force_flag & 4) {
        // k_md_fast found more base-pair candidates or chain neighbours than its shared-memory
        // tables hold and did not run: use the general kernel from here on
        h->h_misc->max_force_flag &= ~4;
        CK(cudaMemcpy(&h->S.misc->max_force_flag, &h->h_misc->max_force_flag, sizeof(int), cudaMemcpyHostToDevice));
        h->fast_kernel = 0;
    }
    if (h->h_misc->max_force_flag & 8)
        return set_err("k_md_big: a warp's chain neighbours or base-pair survivors exceed its shared-memory tables (K = %d)", h->K);
    if (h->h_misc->max_force_flag & 16)
        return set_err("persistent kernel: a wait inside the kernel timed out (exchange barrier or work queue)");
    return 0;
}

// MD steps up to `stop` with device-driven list rebuilds: [rebuild chain, persistent kernel] cycles
// are enqueued ahead without reading anything back; one synchronisation per batch.  Kernels
// enqueued past `stop` (or past an overflow) fall through.
static int md_run_pipelined(sisgpu *h, long long stop)
{
    const int R = h->nrep;
    while ((long long)h->h_ctl->step < stop) {
        const long long cur = (long long)h->h_ctl->step;
        const int nreb_before = h->h_misc->nrebuild;
        int ncyc = (int)ceil((double)(stop - cur) / std::max(1.0, h->est_cycle_steps) * 1.2) + 1;
        ncyc = std::max(1, std::min(ncyc, 32));
        bool local, fast_ok_;
        int want_;
        persistent_choice(h, &local, &fast_ok_, &want_);
        if (h->h_misc->max_force_flag & 4) local = false;
        if (local) ncyc = 1;      // the kernel rebuilds its own lists and always reaches `stop`
        const bool big = !(fast_ok_ && !(h->h_misc->max_force_flag & 4)) && big_ok(h);
        if (big) {
            // [rebuild chain, graph of G step kernels] cycles; a step blocked by the rebuild test makes the
            // rest of its chunk fall through, the next cycle's chain rebuilds
            const int G = std::max(2, h->big_chunk & ~1);
            const double per_cycle = G * h->est_cycle_steps / (h->est_cycle_steps + 0.5 * G);
            ncyc = std::max(1, std::min((int)ceil((double)(stop - cur) / std::max(1.0, per_cycle)) + 1, 96));
            h->h_misc->stop_step = stop;
            CK(cudaMemcpyAsync(&h->S.misc->stop_step, &h->h_misc->stop_step, sizeof(long long), cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemsetAsync(h->d_bar.p, 0, BAR_WORDS * sizeof(unsigned), h->stream));
        }
        for (int c = 0; c < ncyc; c++) {
            if (!local) {
                const int sp = span_begin(h);
                TRY(enqueue_nl_chain(h, 0, stop));
                span_end(h, sp, &h->t_nl);
            }
            const int sp = span_begin(h);
            if (big) TRY(launch_big_chunk(h));
            else TRY(launch_persistent(h, stop));
            span_end(h, sp, &h->t_step);
        }
        TRY(sync_ctl(h));
        CK(cudaGetLastError());
        spans_resolve(h);
        const int nreb = h->h_misc->nrebuild - nreb_before;
        TRY(after_persistent_sync(h));
        const long long adv = (long long)h->h_ctl->step - cur;
        if (nreb > 0 && adv > 0) h->est_cycle_steps = 0.5 * h->est_cycle_steps + 0.5 * std::max(1.0, (double)adv / nreb);
        if (h->ele_cutoff_type == 2 && h->P.flg_ele && nreb > 0) {
            // k_nl_plan refreshed ele_cutoff of the rebuilt replicas (src/neighbor_list.F90:80-82)
            std::vector<RepParams> rp(R);
            CK(cudaMemcpy(rp.data(), h->d_rp.p, R * sizeof(RepParams), cudaMemcpyDeviceToHost));
            for (int r = 0; r < R; r++) h->rep[r].ele_cutoff = rp[r].ele_cutoff;
        }
        if (h->h_misc->lists_invalid) {
            // a row outgrew its capacity during a device-side rebuild: grow (x1.2, as
            // src/neighbor_list.F90:247-298) and rebuild every replica through the host path
            int K = h->K, KB = h->KB;
            if (h->h_misc->nl_overflow > K) K = ((int)(h->h_misc->nl_overflow * 1.2) + 31) / 32 * 32;
            if (h->h_misc->bp_overflow > KB) KB = ((int)(h->h_misc->bp_overflow * 1.2) + 7) / 8 * 8;
            if (getenv("SIS_SPAN_DEBUG")) fprintf(stderr, "[dev %d] step %lld lists invalid: K %d -> %d, KB %d -> %d\n", h->device, (long long)h->h_ctl->step, h->K, K, h->KB, KB);
            h->h_misc->lists_invalid = 0;
            CK(cudaMemcpy(&h->S.misc->lists_invalid, &h->h_misc->lists_invalid, sizeof(int), cudaMemcpyHostToDevice));
            // the host-driven rebuild reads (and may push) the per-replica parameters: bring the host
            // tables up to date with the exchanges done on the device so far
            if (h->remd_on_device) TRY(remd_download(h));
            TRY(alloc_lists(h, K, KB));
            std::vector<int> all(R, 1);
            TRY(build_lists(h, all, false));
            TRY(sync_ctl(h));
        } else if (nreb > 0) {
            TRY(refresh_half_counts(h));
        }
    }
    return 0;
}

// ---- the MD loop --------------------------------------------------------------------------
int sisgpu_md_steps(sisgpu_t *h, int64_t nsteps, int64_t *nrebuild_out)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    if (h->lists_full) return set_err("md_steps needs neighbour lists (call sisgpu_neighbor_list first)");
    CK(cudaSetDevice(h->device));
    if (!h->lists_built) {   // src/job_md.F90:196-198
        std::vector<int> all(h->nrep, 1);
        TRY(build_lists(h, all, false));
    }
    TRY(sync_ctl(h));
    const long long start = (long long)h->h_ctl->step, target = start + nsteps;
    long long nreb0 = h->nrebuild;
    bool force_mc = false;
    const int R = h->nrep;
    const unsigned mask_all = 0x7FEu;
    // launch_mode 2 with device-driven rebuilds (md_run_pipelined) needs no host-side trigger
    // Monte-Carlo base-pair limit: only the replica-local kernel carries it
    bool local_kernel_, fast_ok0_;
    int want0_;
    persistent_choice(h, &local_kernel_, &fast_ok0_, &want0_);
    const bool mc_ok = h->nstep_bp_MC == 0 || (h->mc_persistent && local_kernel_);
    const bool pipelined = h->launch_mode == 2 && mc_ok && !h->P.flg_bias_rg && h->pipeline && !h->P.flg_bias_ss;
    if (pipelined && h->flg_replica && h->nstep_exchange > 0 && h->remd_device && h->nrep_all <= REMD_DEV_MAX_LABELS) {
        // REMD with the exchange step on the device: MD intervals and exchange chains are enqueued
        // back to back; the host tables are refreshed once at the end
        TRY(remd_upload(h));
        h->remd_on_device = true;
        const long long nex = h->nstep_exchange;
        while ((long long)h->h_ctl->step < target) {
            const long long cur = (long long)h->h_ctl->step;
            bool local_, fast_ok_;
            int want_;
            persistent_choice(h, &local_, &fast_ok_, &want_);
            if (local_ && !(h->h_misc->max_force_flag & 4) && h->remd_inkernel && (h->nranks == 1 || h->peers_ready) &&
                h->nstep_bp_MC == 0 && h->nrep_all <= REMD_LOCAL_MAX &&
                (R <= h->coop_blocks || rep_threads_of(h, h->nrep_all, true) > 0)) {     // every replica's CTA co-resident
                // the exchange steps run inside the replica-local kernel: ONE launch up to the target
                const long long stop = std::min<long long>(target, cur + (1ll << 20));
                h->S.remd = h->d_remd.p;
                const int sp = span_begin(h);
                const int rc = launch_persistent(h, stop);
                h->S.remd = nullptr;
                if (rc) return rc;
                span_end(h, sp, &h->t_step);
                if (!h->last_launch_local) return set_err("internal: replica-local kernel not selected");
                TRY(sync_ctl(h));
                CK(cudaGetLastError());
                spans_resolve(h);
                TRY(after_persistent_sync(h));
                if (h->h_misc->max_force_flag & 16) return set_err("in-kernel exchange: a CTA did not reach the exchange barrier");
                if ((long long)h->h_ctl->step != stop) return set_err("internal: replica-local kernel stopped at step %lld, expected %lld", (long long)h->h_ctl->step, stop);
            } else if (local_ && !(h->h_misc->max_force_flag & 4)) {
                long long s = cur;
                for (int nint = 0; s < target && nint < 64; nint++) {
                    const long long stop = std::min<long long>(target, (s / nex + 1) * nex);
                    const int sp = span_begin(h);
                    TRY(launch_persistent(h, stop));
                    span_end(h, sp, &h->t_step);
                    if (!h->last_launch_local) return set_err("internal: replica-local kernel not selected");
                    if (stop % nex == 0) TRY(enqueue_exchange(h));
                    s = stop;
                }
                TRY(sync_ctl(h));
                CK(cudaGetLastError());
                spans_resolve(h);
                TRY(after_persistent_sync(h));
                if ((long long)h->h_ctl->step != s) return set_err("internal: replica-local kernel stopped at step %lld, expected %lld", (long long)h->h_ctl->step, s);
            } else {
                const long long stop = std::min<long long>(target, (cur / nex + 1) * nex);
                TRY(md_run_pipelined(h, stop));
                if (stop % nex == 0) TRY(enqueue_exchange(h));
            }
        }
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaGetLastError());
        spans_resolve(h);
        TRY(remd_download(h));
        h->remd_on_device = false;
        if (h->ele_cutoff_type == 2 && h->P.flg_ele) {   // refreshed by the device-side rebuilds
            std::vector<RepParams> rp(R);
            CK(cudaMemcpy(rp.data(), h->d_rp.p, R * sizeof(RepParams), cudaMemcpyDeviceToHost));
            for (int r = 0; r < R; r++) h->rep[r].ele_cutoff = rp[r].ele_cutoff;
        }
        TRY(refresh_half_counts(h));
        TRY(sync_ctl(h));
        // the energy evaluations of the device-side exchange steps carry the hard abort of src/energy_stage.F90:22-26
        if (h->h_misc->max_force_flag & 2) return set_err("Error: Z-coordinate < 0.85 * stage_sigma in energy_stage. Abort!");
        if (h->h_misc->max_force_flag & 1) {
            fprintf(stderr, "Warning: maxval(forces) > 100.0 during the last md_steps call (step %lld)\n", (long long)h->h_ctl->step);
            h->h_misc->max_force_flag = 0;
            CK(cudaMemcpy(&h->S.misc->max_force_flag, &h->h_misc->max_force_flag, sizeof(int), cudaMemcpyHostToDevice));
        }
        if (nrebuild_out) *nrebuild_out = h->nrebuild - nreb0;
        return 0;
    }
    while ((long long)h->h_ctl->step < target) {
        // rebuild requested by the previous position update (src/job_md.F90:388-395)
        bool any = false;
        std::vector<int> act(R, 0);
        if (!pipelined)
            for (int r = 0; r < R; r++) if (h->h_need[r]) { act[r] = 1; any = true; h->nrebuild++; }
        if (any) {
            // a persistent kernel that stopped at the rebuild test may have left the base-pair
            // scatter of the step it did not execute in f_bp
            CK(cudaMemsetAsync(h->d_f_bp.p, 0, (size_t)R * h->nmp * 3 * sizeof(double), h->stream));
            TRY(build_lists(h, act, false));
            if (h->nstep_bp_MC > 0) force_mc = true;
            TRY(sync_ctl(h));
        }
        long long cur = (long long)h->h_ctl->step;
        long long stop = target;
        if (h->flg_replica && h->nstep_exchange > 0) {
            long long next_ex = (cur / h->nstep_exchange + 1) * (long long)h->nstep_exchange;
            stop = std::min(stop, next_ex);
        }
        int nchunk = (int)std::min<long long>(h->chunk, stop - cur);
        timer_begin(h);
        const int use_mask = h->nstep_bp_MC > 0 ? 1 : 0;
        const bool uniform = h->nstep_bp_MC <= 1;   // every step has the same kernel sequence
        if (h->launch_mode == 2 && mc_ok && !h->P.flg_bias_rg) {
            if (pipelined) {
                // device-driven rebuilds: no host round trip between a rebuild and the next steps
                timer_end(h, h->t_step);   // (closes the pair opened above; the loop keeps its own events)
                TRY(md_run_pipelined(h, stop));
                timer_begin(h);
            } else {
                // persistent warp-specialised kernel: runs until `stop` or until a replica needs new lists
                TRY(launch_persistent(h, stop));
            }
        } else if (h->launch_mode >= 1 && uniform && nchunk == h->chunk) {
            if (!h->graph_chunk || h->graph_chunk_steps != nchunk) {
                invalidate_graphs(h);
                cudaGraph_t graph = nullptr;
                CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                const long long l0 = h->launches;
                for (int s = 0; s < nchunk; s++) {
                    if (h->nstep_bp_MC == 1) launch_bp_eval<0>(h, -1, true, 0.0, 0, 0, 1, 0, 1);
                    launch_rg(h, -1);
                    enqueue_step(h, step_args(h, mask_all, use_mask, -1, 1));
                }
                h->graph_launches_per_chunk = h->launches - l0;
                h->launches = l0;
                CK(cudaStreamEndCapture(h->stream, &graph));
                CK(cudaGraphInstantiate(&h->graph_chunk, graph, 0));
                CK(cudaGraphDestroy(graph));
                h->graph_chunk_steps = nchunk;
            }
            CK(cudaGraphLaunch(h->graph_chunk, h->stream));
            h->launches += h->graph_launches_per_chunk;
        } else {
            for (int s = 0; s < nchunk; s++) {
                const long long istep = cur + s + 1;
                bool mc = false;
                if (h->nstep_bp_MC > 0) mc = (istep % h->nstep_bp_MC == 0) || (s == 0 && force_mc);
                if (mc) launch_bp_eval<0>(h, -1, true, 0.0, 0, 0, 1, 0, 1);
                launch_rg(h, -1);
                enqueue_step(h, step_args(h, mask_all, use_mask, -1, 1));
            }
        }
        force_mc = false;
        TRY(sync_ctl(h));
        CK(cudaGetLastError());
        timer_end(h, h->t_step);
        TRY(after_persistent_sync(h));
        // exchange step (src/job_md.F90:474-481,514-570)
        if (h->flg_replica && h->nstep_exchange > 0 &&
            (long long)h->h_ctl->step % h->nstep_exchange == 0 && (long long)h->h_ctl->step > cur) {
            std::vector<double> re(2 * (size_t)h->nrep_all);
            const int mc_saved = h->flg_bp_MC;
            if (h->nstep_bp_MC > 0) h->flg_bp_MC = 1;   // src/job_md.F90:476-479
            TRY(sisgpu_energy_replica(h, nullptr, re.data()));
            h->flg_bp_MC = mc_saved;
            timer_begin(h);
            TRY(sisgpu_replica_exchange(h, re.data(), nullptr, nullptr, nullptr));
            timer_end(h, h->t_rex);
        }
    }
    if (h->h_misc->max_force_flag & 1) {
        fprintf(stderr, "Warning: maxval(forces) > 100.0 during the last md_steps call (step %lld)\n", (long long)h->h_ctl->step);
        h->h_misc->max_force_flag = 0;
        CK(cudaMemcpy(&h->S.misc->max_force_flag, &h->h_misc->max_force_flag, sizeof(int), cudaMemcpyHostToDevice));
    }
    if (nrebuild_out) *nrebuild_out = h->nrebuild - nreb0;
    return 0;
}

// job_check_force: src/job_check_force.F90:53-80
int sisgpu_check_force(sisgpu_t *h, double tempK, double *f_energy, double *f_force, double *maxdiff)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    if (h->nrep != 1) return set_err("check_force runs with one replica (IREP = 1)");
    const float small_f = 0.001f;
    const double small = small_f;   // `small = 0.001` is a default-real literal in the reference
    const int N = h->nmp;
    TRY(sisgpu_force(h, 0, f_force));
    std::vector<double> x(3 * (size_t)N);
    TRY(sisgpu_get_state(h, 0, x.data(), nullptr, nullptr));
    double e[ENE_N], md = 0.0;
    const int cur = h->h_ctl->cur;
    for (int imp = 0; imp < N; imp++)
        for (int k = 0; k < 3; k++) {
            double4 p = make_double4(x[3 * imp], x[3 * imp + 1], x[3 * imp + 2], 0.0), q = p;
            double *c = k == 0 ? &q.x : k == 1 ? &q.y : &q.z;
            const double save = *c;
            *c = save + small;
            CK(cudaMemcpyAsync(h->S.pos[cur] + imp, &q, sizeof q, cudaMemcpyHostToDevice, h->stream));
            TRY(sisgpu_energy_sumup(h, 0, tempK, 1, e));
            const double e_save = e[0];
            *c = save - small;
            CK(cudaMemcpyAsync(h->S.pos[cur] + imp, &q, sizeof q, cudaMemcpyHostToDevice, h->stream));
            TRY(sisgpu_energy_sumup(h, 0, tempK, 1, e));
            f_energy[3 * imp + k] = -(e_save - e[0]) / (2.0 * small);
            CK(cudaMemcpyAsync(h->S.pos[cur] + imp, &p, sizeof p, cudaMemcpyHostToDevice, h->stream));
            md = std::max(md, fabs(f_energy[3 * imp + k] - f_force[3 * imp + k]));
        }
    CK(cudaStreamSynchronize(h->stream));
    if (maxdiff) *maxdiff = md;
    return 0;
}

int sisgpu_launch_count(sisgpu_t *h, int64_t *n) { *n = h->launches; return 0; }

int sisgpu_get_cta_profile(sisgpu_t *h, int ncta, double *out)
{
    if (!h->S.prof) return set_err("option profile is off");
    std::vector<unsigned long long> p(8 * (size_t)ncta);
    CK(cudaMemcpy(p.data(), h->d_prof.p + 16, p.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < p.size(); i++) out[i] = (double)p[i];
    return 0;
}

int sisgpu_set_option(sisgpu_t *h, const char *name, double value)
{
    std::string s(name);
    if (s == "launch_mode") h->launch_mode = (int)value;
    else if (s == "chunk") { h->chunk = std::max(1, (int)value); invalidate_graphs(h); }
    else if (s == "timing") h->timing = (int)value;
    else if (s == "fast_kernel") h->fast_kernel = (int)value;
    else if (s == "big_kernel") h->big_kernel = (int)value;
    else if (s == "big_chunk") { h->big_chunk = std::max(2, (int)value); invalidate_graphs(h); }
    else if (s == "pipeline") h->pipeline = (int)value;
    else if (s == "remd_device") h->remd_device = (int)value;
    else if (s == "remd_inkernel") h->remd_inkernel = (int)value;
    else if (s == "list_capacity") {    // row capacity K (tests: a tight K makes device-side rebuilds overflow)
        if (!h->finalized) return set_err("finalize_setup not called");
        TRY(alloc_lists(h, std::max(8, (int)value), h->KB));
        h->lists_built = false;
    }
    else if (s == "inject_invalid") {   // test hook: the n-th device-side rebuild from now reports invalid lists
        if (!h->finalized) return set_err("finalize_setup not called");
        TRY(sync_ctl(h));
        const int at = h->h_misc->nrebuild + (int)value;
        CK(cudaMemcpy(&h->S.misc->inject_invalid_at, &at, sizeof(int), cudaMemcpyHostToDevice));
    }
    else if (s == "fast_max_cta") { h->fast_max_cta = (int)value; h->lists_built = false; }
    else if (s == "mc_persistent") h->mc_persistent = (int)value;
    else if (s == "precision") { h->P.mixed = value != 0 ? 1 : 0; invalidate_graphs(h); }   // 0: fp64 (default); 1: mixed (fp32 pair scalars)
    else if (s == "rep_kernel") h->rep_kernel = (int)value;
    else if (s == "rep_threads") h->rep_threads = (int)value;
    else if (s == "loc_help") h->S.dbg_no_loc_help = value == 0 ? 1 : 0;
    else if (s == "bp_groups") h->S.dbg_bp_groups = value != 0 ? 1 : 0;
    else if (s == "local_kernel") { h->local_kernel = (int)value; h->lists_built = false; }
    else if (s == "profile") {   // phase clocks of the persistent kernel (block 0): get_timers out[4..8]
        if (value != 0) { TRY(h->d_prof.alloc(16 + 8 * FAST_MAX_CTA)); h->S.prof = h->d_prof.p; }
        else h->S.prof = nullptr;
        invalidate_graphs(h);
    }
    else if (s == "threads_per_bead") {
        int g = (int)value;
        if (g != 4 && g != 8 && g != 16 && g != 32) return set_err("threads_per_bead must be 4, 8, 16 or 32");
        h->G = g;
        invalidate_graphs(h);
    } else return set_err("unknown option %s", name);
    return 0;
}

// ---- output staging ------------------------------------------------------------
// sisgpu_snapshot_begin reduces what the writers of a save step need (flags: 1 float32 DCD frame,
// 2 .bp records [ene_bp <= -kT], 4 .bpall records instead, 8 the 11 energies of energy_sumup with
// the cached bp energies + Ekin) into a staging slot on the device and starts the device-to-host
// copies on a second stream; MD steps enqueued afterwards run concurrently with the copies.
// sisgpu_snapshot_end waits for that slot only and hands the data out.
static void snap_free(sisgpu::SnapSlot &s)
{
    if (s.d_xyz) cudaFree(s.d_xyz); if (s.d_e) cudaFree(s.d_e); if (s.d_cnt) cudaFree(s.d_cnt);
    if (s.d_nrec) cudaFree(s.d_nrec); if (s.d_ij) cudaFree(s.d_ij);
    if (s.h_xyz) cudaFreeHost(s.h_xyz); if (s.h_e) cudaFreeHost(s.h_e); if (s.h_nrec) cudaFreeHost(s.h_nrec);
    if (s.h_ij) cudaFreeHost(s.h_ij); if (s.h_ene) cudaFreeHost(s.h_ene);
    if (s.staged) cudaEventDestroy(s.staged); if (s.copied) cudaEventDestroy(s.copied);
    s = sisgpu::SnapSlot();
}
static int snap_alloc(sisgpu *h, sisgpu::SnapSlot &s, int cap)
{
    const int N = h->nmp;
    if (!s.d_xyz) {
        CK(cudaMalloc(&s.d_xyz, (size_t)N * 3 * sizeof(float))); CK(cudaMallocHost(&s.h_xyz, (size_t)N * 3 * sizeof(float)));
        CK(cudaMalloc(&s.d_cnt, (size_t)(N + 1) * sizeof(int)));
        CK(cudaMalloc(&s.d_nrec, sizeof(int))); CK(cudaMallocHost(&s.h_nrec, sizeof(int)));
        CK(cudaMallocHost(&s.h_ene, 12 * sizeof(double)));
        CK(cudaEventCreateWithFlags(&s.staged, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    }
    if (cap > s.cap) {
        if (s.d_ij) { cudaFree(s.d_ij); cudaFree(s.d_e); cudaFreeHost(s.h_ij); cudaFreeHost(s.h_e); }
        CK(cudaMalloc(&s.d_ij, (size_t)cap * sizeof(int2))); CK(cudaMalloc(&s.d_e, (size_t)cap * sizeof(float)));
        CK(cudaMallocHost(&s.h_ij, (size_t)cap * sizeof(int2))); CK(cudaMallocHost(&s.h_e, (size_t)cap * sizeof(float)));
        s.cap = cap;
    }
    return 0;
}
int sisgpu_snapshot_begin(sisgpu_t *h, int irep, double tempK, int flags, int *slot_out)
{
    if (!h->finalized) return set_err("finalize_setup not called");
    if (irep < 0 || irep >= h->nrep) return set_err("irep out of range");
    CK(cudaSetDevice(h->device));
    const int si = h->snap_next;
    sisgpu::SnapSlot &s = h->snap[si];
    if (s.busy) return set_err("snapshot slot %d still pending: call sisgpu_snapshot_end first", si);
    const int N = h->nmp;
    TRY(snap_alloc(h, s, std::max(s.cap, std::max(256, N))));
    s.flags = flags; s.irep = irep;
    SnapDev D{s.d_xyz, s.d_cnt, s.d_ij, s.d_e, s.d_nrec, s.cap};
    const double kT = tempK * SIS_BOLTZ_KCAL_MOL;
    if (flags & 1) { k_snap_frame<<<(N + 255) / 256, 256, 0, h->stream>>>(h->S, D, irep); h->launches++; }
    if (flags & 6) {
        const int all = (flags & 4) ? 1 : 0;
        k_snap_bp_count<<<(N + 255) / 256, 256, 0, h->stream>>>(h->S, D, irep, kT, all);
        k_snap_bp_scan<<<1, 1024, 0, h->stream>>>(h->S, D);
        k_snap_bp_fill<<<(N + 255) / 256, 256, 0, h->stream>>>(h->S, D, irep, kT, all);
        h->launches += 3;
    }
    if (flags & 8) {
        launch_rg(h, irep);
        TRY(energy_launch(h, irep, nullptr));
        k_kinetic<<<1, 256, 0, h->stream>>>(h->S, h->d_kin.p, irep);
        h->launches++;
    }
    CK(cudaEventRecord(s.staged, h->stream));
    // copies on the second stream: the compute stream is free for the next MD steps.  The energies
    // live in buffers the next energy call rewrites, so they are copied on the compute stream.
    if (flags & 8) {
        CK(cudaMemcpyAsync(s.h_ene, h->S.ene_out + (size_t)irep * ENE_N, ENE_N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(s.h_ene + 11, h->d_kin.p + irep, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(s.staged, h->stream));
    }
    CK(cudaStreamWaitEvent(h->stream2, s.staged, 0));
    if (flags & 1) CK(cudaMemcpyAsync(s.h_xyz, s.d_xyz, (size_t)N * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream2));
    if (flags & 6) {
        CK(cudaMemcpyAsync(s.h_nrec, s.d_nrec, sizeof(int), cudaMemcpyDeviceToHost, h->stream2));
        CK(cudaMemcpyAsync(s.h_ij, s.d_ij, (size_t)s.cap * sizeof(int2), cudaMemcpyDeviceToHost, h->stream2));
        CK(cudaMemcpyAsync(s.h_e, s.d_e, (size_t)s.cap * sizeof(float), cudaMemcpyDeviceToHost, h->stream2));
    }
    CK(cudaEventRecord(s.copied, h->stream2));
    s.busy = true;
    h->snap_next ^= 1;
    if (slot_out) *slot_out = si;
    return 0;
}
int sisgpu_snapshot_end(sisgpu_t *h, int slot, float *xyz32, int *nbp, int *bp_i, int *bp_j, float *bp_e, int cap,
                        double *energies, double *ekin)
{
    if (slot < 0 || slot > 1) return set_err("snapshot slot out of range");
    sisgpu::SnapSlot &s = h->snap[slot];
    if (!s.busy) return set_err("snapshot slot %d is not pending", slot);
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(s.copied));
    s.busy = false;
    const int N = h->nmp;
    if ((s.flags & 1) && xyz32) memcpy(xyz32, s.h_xyz, (size_t)N * 3 * sizeof(float));
    if (s.flags & 6) {
        const int n = *s.h_nrec;
        if (nbp) *nbp = n;
        if (n > s.cap) return set_err("snapshot: %d base-pair records exceed the staging capacity %d", n, s.cap);
        if (bp_i && bp_j && bp_e) {
            if (n > cap) return set_err("snapshot: %d base-pair records, caller capacity %d", n, cap);
            // the reference's record order is bp_mp order: chain pair, then (imp, jmp)
            std::vector<int> idx(n);
            for (int k = 0; k < n; k++) idx[k] = k;
            std::sort(idx.begin(), idx.end(), [&](int a, int b) {
                const int2 A = s.h_ij[a], B = s.h_ij[b];
                if (!h->lists_full) {
                    const int ac = h->ichain[A.x - 1], bc = h->ichain[B.x - 1];
                    if (ac != bc) return ac < bc;
                    const int ajc = h->ichain[A.y - 1], bjc = h->ichain[B.y - 1];
                    if (ajc != bjc) return ajc < bjc;
                }
                if (A.x != B.x) return A.x < B.x;
                return A.y < B.y;
            });
            for (int k = 0; k < n; k++) { bp_i[k] = s.h_ij[idx[k]].x; bp_j[k] = s.h_ij[idx[k]].y; bp_e[k] = s.h_e[idx[k]]; }
        }
    } else if (nbp) *nbp = 0;
    if (s.flags & 8) {
        if (energies) memcpy(energies, s.h_ene, ENE_N * sizeof(double));
        if (ekin) *ekin = s.h_ene[11];
    }
    return 0;
}

const char *sisgpu_last_step_kernel(sisgpu_t *h) { return h->last_step_kernel; }

int sisgpu_mark(sisgpu_t *h, int which)
{
    if (which < 0 || which > 1) return set_err("mark index must be 0 or 1");
    CK(cudaSetDevice(h->device));
    if (!h->ev_mark[which]) CK(cudaEventCreate(&h->ev_mark[which]));
    CK(cudaEventRecord(h->ev_mark[which], h->stream));
    return 0;
}
int sisgpu_elapsed_ms(sisgpu_t *h, double *ms)
{
    if (!h->ev_mark[0] || !h->ev_mark[1]) return set_err("sisgpu_mark(0) / sisgpu_mark(1) not called");
    CK(cudaEventSynchronize(h->ev_mark[1]));
    float f = 0;
    CK(cudaEventElapsedTime(&f, h->ev_mark[0], h->ev_mark[1]));
    *ms = f;
    return 0;
}

int sisgpu_get_timers(sisgpu_t *h, double *out, int reset)
{
    out[0] = h->t_step; out[1] = h->t_nl; out[2] = h->t_ene; out[3] = h->t_rex;
    for (int i = 4; i < 16; i++) out[i] = 0;
    if (h->S.prof) {
        unsigned long long p[16];
        CK(cudaMemcpy(p, h->d_prof.p, sizeof p, cudaMemcpyDeviceToHost));
        for (int i = 0; i < 12; i++) out[4 + i] = (double)p[i];
        if (reset) CK(cudaMemset(h->d_prof.p, 0, (16 + 8 * FAST_MAX_CTA) * sizeof(unsigned long long)));
    }
    if (reset) h->t_step = h->t_nl = h->t_ene = h->t_rex = 0;
    return 0;
}
