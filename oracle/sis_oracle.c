/*
 * sis_oracle.c -- CPU restatement of the Hori-Lab/SIS per-step hot path (fp64).
 *
 * Built twice (oracle/Makefile): libsis_oracle.so is serial (the parity checker);
 * libsis_oracle_omp.so (-fopenmp) carries the reference's OpenMP threading -- one parallel
 * region in force() with per-thread force buffers and `omp do ... nowait` per term
 * (src/force.F90:27-69), `parallel do reduction(+)` in the pair / stage / kinetic energies
 * (src/energy_wca.F90:19-33, energy_ele_DH.F90:20-38, energy_stage.F90:28-37,
 * energy_kinetic.F90:23-27), `parallel do` in the bp energies (src/energy_bp_triplet.F90:36-87)
 * and in the integrator loop (src/job_md.F90:416-456); neighbor_list stays serial as in the
 * reference.  It is the CPU baseline of bench.py.
 *
 * TEST INFRASTRUCTURE ONLY -- see sis_oracle.h.  The product (sis_b200/) never
 * links or calls this.  "parity unpinned" at 1e-10 against the Fortran binary
 * (no Fortran compiler in the image); pinned against the reference's own
 * golden energies and CHECK_FORCE criterion (tests/test_oracle_*.py).
 *
 * Citations are file:line relative to /root/reference.
 */
#define _POSIX_C_SOURCE 200809L
#include "sis_oracle.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- src/const_phys.F90:5-18 ------------------------------------------------ */
static const double EPS0 = 8.8541878128e-12;
static const double ELE = 1.602176634e-19;
static const double BOLTZ_J = 1.380649e-23;
static const double N_AVO = 6.02214076e23;
static const double KCAL2JOUL = 4184.0;
#define O_PI (4.0 * atan(1.0))
#define JOUL2KCAL (1.0 / KCAL2JOUL)
#define JOUL2KCAL_MOL (JOUL2KCAL * N_AVO)
#define BOLTZ_KCAL_MOL (BOLTZ_J * JOUL2KCAL_MOL)
#define SMALL_VALUE 2.220446049250313e-16 /* epsilon(PI) */

enum { SEQ_A = 0, SEQ_U = 1, SEQ_G = 2, SEQ_C = 3, SEQ_D = 4 };
enum { MOLT_RNA = 1, MOLT_CIRC = 2 };
enum { BPT_GC = 1, BPT_AU, BPT_GU, BPT_CG, BPT_UA, BPT_UG };
enum { INT_GJF = 1, INT_BD = 2 };
enum { REPT_TEMP = 0, REPT_TWZDCF = 1, REPT_ION = 2, REPT_RG = 3, REPT_MAX = 4 };
#define MAX_REP_PER_DIM 256
#define MAX_REP_DIM 3

/* src/var_potential.F90:63-82 */
typedef struct {
    double cutoff_ddist, U0, bond_k, bond_r;
    double angl_k1, angl_k2, angl_k3, angl_k4;
    double angl_theta1, angl_theta2, angl_theta3, angl_theta4;
    double dihd_k1, dihd_k2, dihd_phi1, dihd_phi2;
} bp_para;

typedef struct {
    /* lists: src/var_potential.F90:85-112 */
    int nwca, nwca_max, nele, nele_max, nbp, nbp_max;
    int *wca_mp;     /* [nwca][2] 0-based */
    int *ele_mp;     /* [nele][2] */
    int *bp_mp;      /* [nbp][7]: i j i-1 j-1 i+1 j+1 (0-based) type(1..6) */
    double *bp_coef; /* [nbp][2] dH, dS*1e-3 */
    int *bp_status, *bp_status_MC;
    double *ene_bp;
    double *for_bp;  /* [nbp][6][3] */
    /* state */
    double *xyz, *velos, *accels, *xyz_move; /* [nmp][3] */
    /* per-replica parameters */
    double tempK, ionic_strength;
    double ele_coef, lambdaD, ele_cutoff, diele, diele_dTcoef;
    double bias_rg_0, rg, twz_force;
    double *twz_DCF_forces; /* [ntwz][3] */
    int flg_bp_energy;
    uint64_t mc_counter;
} replica;

struct sis_oracle {
    /* secondary-structure restriction of the bp map (src/read_ss.F90): sorted keys imp * nmp + jmp
     * (0-based, imp < jmp) of the pairs read_ss accepts; flg_ss = 0: all pairwise */
    int flg_ss, n_ss;
    long long *ss_keys;
    /* topology */
    int nmp, nchains;
    int *nmp_chain, *moltypes, *chain_start;
    int *seq, *ichain_mp, *lmp_mp, *has_charge, *is_frozen;
    double *mass;
    /* bonded */
    double bond_k, bond_r0, angl_k, angl_t0, dih_k, dih_w, dih_p0;
    int flg_angl_ReB, flg_dih_cos, flg_dih_exp;
    int nbond, nangl, ndih;
    int *bond_mp, *angl_mp, *dih_mp;
    /* wca */
    double wca_sigma, wca_eps;
    /* bp */
    int bp_model, bp_min_loop, max_bp_per_nt, nstep_bp_MC;
    double bp_cutoff_energy, bp_cutoff_dist;
    bp_para bp_paras[7];
    double NN_dH[22], NN_dS[22], dH0, dS0, NNend_dH[7], NNend_dS[7], dHend0, dSend0;
    int flg_NNend, have_nn;
    int bp3_hash[4096 + 1];
    double bp3_dH[469], bp3_dS[469];
    int bp3_n;
    int flg_bias_ss;
    double bias_ss_force;
    /* ele */
    int flg_ele, ele_cutoff_type, excl_bond, excl_angle, temp_independent;
    double ele_cutoff_inp, length_per_charge, tempK_ref;
    /* pbc */
    int flg_pbc;
    double box[3], box_half[3];
    /* md */
    int integrator;
    double dt, viscosity_Pas, nl_margin;
    uint64_t seed;
    int64_t istep;
    int flg_bp_MC;
    /* biases */
    int flg_stage;
    double stage_sigma, stage_eps;
    int ntwz_DCF, *twz_DCF_pairs;
    double *twz_DCF_direction;
    int ntwz_FR, *twz_FR_pairs;
    double *twz_FR_k, *twz_FR_init, *twz_FR_velo;
    int flg_bias_rg, bias_rg_pott;
    double bias_rg_k;
    int nrest, *rest_id;
    double *rest_rcut, *rest_para;
    /* replicas */
    int nrep_proc;
    replica *rep;
    /* REMD */
    int flg_replica, nrep_all, ndim_replica, nrep[REPT_MAX], flg_repvar[REPT_MAX];
    int type_array[MAX_REP_DIM];
    int *exchange_pair_tb; /* [2*ndim][nrep_all] 1-based labels, 0 = none */
    int idx_pair;          /* 1-based */
    int *rep2lab, *lab2rep;
    double *lab2val;       /* [REPT_MAX][nrep_all] (label-1) */
    double replica_values[REPT_MAX][MAX_REP_PER_DIM];
    int nstep_exchange;
    uint64_t exchange_counter;
    int finalized;
    double t_nl;           /* wall seconds spent in neighbor_list inside oracle_md_step (bench.py) */
    double *forces_t;      /* forces_t(3, nmp, 0:nthreads-1) of src/force.F90:19 */
    size_t forces_t_n;
    int *bp_excess;        /* nt_bp_excess(nmp), shared by the threads of force_bp_limit_triplet */
    int flg_save_bp;       /* energy_replica's flg_save_bp: the exchange step is also a save step */
};

static char g_err[512];
static int fail(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return 1;
}
const char *oracle_last_error(void) { return g_err; }

static void *xcalloc(size_t n, size_t s) { return calloc(n ? n : 1, s); }
static int *dup_int(const int *a, int n)
{
    int *p = (int *)xcalloc(n, sizeof(int));
    if (a) memcpy(p, a, (size_t)n * sizeof(int));
    return p;
}
static double *dup_dbl(const double *a, int n)
{
    double *p = (double *)xcalloc(n, sizeof(double));
    if (a) memcpy(p, a, (size_t)n * sizeof(double));
    return p;
}

/* Fortran x**n with integer n: repeated squaring */
static double powi(double x, int n)
{
    double r = 1.0;
    while (n) {
        if (n & 1) r *= x;
        x *= x;
        n >>= 1;
    }
    return r;
}
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(const double *a, const double *b, double *c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

/* src/pbc.F90:140-165 pbc_vec_d */
static void pbc_vec_d(const sis_oracle *o, const double *a, const double *b, double *v)
{
    for (int i = 0; i < 3; i++) v[i] = a[i] - b[i];
    if (!o->flg_pbc) return;
    for (int i = 0; i < 3; i++) {
        if (v[i] > o->box_half[i]) v[i] -= o->box[i];
        else if (v[i] < -o->box_half[i]) v[i] += o->box[i];
    }
}

/* src/pbc.F90:196-226 pbc_wrap: wraps into [0, L] */
static void pbc_wrap(const sis_oracle *o, replica *r)
{
    for (int imp = 0; imp < o->nmp; imp++)
        for (int i = 0; i < 3; i++) {
            double x = r->xyz[3 * imp + i], L = o->box[i];
            if (x > L) r->xyz[3 * imp + i] = x - L * ((int)((x - L) / L) + 1);
            else if (x < 0.0) r->xyz[3 * imp + i] = x + L * ((int)((0.0 - x) / L) + 1);
        }
}

/* ---- lifecycle -------------------------------------------------------------- */
sis_oracle *oracle_create(void)
{
    sis_oracle *o = (sis_oracle *)calloc(1, sizeof *o);
    o->nrep_proc = 1;
    o->nl_margin = 10.0;
    o->viscosity_Pas = 1e-5;
    o->integrator = INT_GJF;
    o->bp_model = 5;
    o->bp_min_loop = 3;
    o->max_bp_per_nt = 1;
    o->bp_cutoff_energy = 0.001;
    o->excl_bond = o->excl_angle = 1;
    o->ele_cutoff_type = 1;
    o->idx_pair = 1;
    return o;
}

static void free_replica(replica *r)
{
    free(r->wca_mp); free(r->ele_mp); free(r->bp_mp); free(r->bp_coef);
    free(r->bp_status); free(r->bp_status_MC); free(r->ene_bp); free(r->for_bp);
    free(r->xyz); free(r->velos); free(r->accels); free(r->xyz_move);
    free(r->twz_DCF_forces);
}

void oracle_destroy(sis_oracle *o)
{
    if (!o) return;
    if (o->rep) {
        for (int i = 0; i < o->nrep_proc; i++) free_replica(&o->rep[i]);
        free(o->rep);
    }
    free(o->ss_keys);
    free(o->nmp_chain); free(o->moltypes); free(o->chain_start); free(o->seq);
    free(o->ichain_mp); free(o->lmp_mp); free(o->has_charge); free(o->is_frozen);
    free(o->mass); free(o->bond_mp); free(o->angl_mp); free(o->dih_mp);
    free(o->twz_DCF_pairs); free(o->twz_DCF_direction); free(o->twz_FR_pairs);
    free(o->twz_FR_k); free(o->twz_FR_init); free(o->twz_FR_velo);
    free(o->rest_id); free(o->rest_rcut); free(o->rest_para);
    free(o->exchange_pair_tb); free(o->rep2lab); free(o->lab2rep); free(o->lab2val);
    free(o->forces_t); free(o->bp_excess);
    free(o);
}

/* src/read_fasta.F90:133-143: beads are numbered chain after chain */
int oracle_set_topology(sis_oracle *o, int nmp, int nchains, const int *nmp_chain,
                        const int *moltypes, const int *seq, const int *has_charge,
                        const int *is_frozen, const double *mass)
{
    int tot = 0;
    for (int c = 0; c < nchains; c++) tot += nmp_chain[c];
    if (tot != nmp) return fail("sum(nmp_chain)=%d != nmp=%d", tot, nmp);
    o->nmp = nmp;
    o->nchains = nchains;
    o->nmp_chain = dup_int(nmp_chain, nchains);
    o->moltypes = dup_int(moltypes, nchains);
    o->chain_start = (int *)xcalloc(nchains, sizeof(int));
    o->seq = dup_int(seq, nmp);
    o->ichain_mp = (int *)xcalloc(nmp, sizeof(int));
    o->lmp_mp = (int *)xcalloc(nmp, sizeof(int));
    o->has_charge = dup_int(has_charge, nmp);
    o->is_frozen = dup_int(is_frozen, nmp);
    o->mass = dup_dbl(mass, nmp);
    int imp = 0;
    for (int c = 0; c < nchains; c++) {
        o->chain_start[c] = imp;
        for (int i = 0; i < nmp_chain[c]; i++, imp++) {
            o->ichain_mp[imp] = c;
            o->lmp_mp[imp] = i; /* 0-based position in chain */
        }
    }
    if (!has_charge)
        for (int i = 0; i < nmp; i++) o->has_charge[i] = 1;
    if (!mass) /* src/job_md.F90:116-133 */
        for (int i = 0; i < nmp; i++) {
            static const double m[5] = {328.212, 305.164, 344.212, 304.182, 320.443};
            o->mass[i] = m[o->seq[i]];
        }
    return 0;
}

int oracle_set_replicas(sis_oracle *o, int nrep_proc)
{
    if (o->rep) return fail("replicas already set");
    o->nrep_proc = nrep_proc;
    return 0;
}

int oracle_set_ff_bonded(sis_oracle *o, double bond_k, double bond_r0, int angl_type,
                         double angl_k, double angl_t0, int dih_flags, double dih_k,
                         double dih_w, double dih_p0)
{
    o->bond_k = bond_k; o->bond_r0 = bond_r0;
    o->flg_angl_ReB = angl_type == 1;
    o->angl_k = angl_k; o->angl_t0 = angl_t0;
    o->flg_dih_cos = dih_flags & 1;
    o->flg_dih_exp = (dih_flags >> 1) & 1;
    o->dih_k = dih_k; o->dih_w = dih_w; o->dih_p0 = dih_p0;
    return 0;
}

int oracle_set_ff_wca(sis_oracle *o, double sigma, double eps)
{
    o->wca_sigma = sigma; o->wca_eps = eps;
    return 0;
}

/* src/read_force_field.F90:185-258 */
int oracle_set_ff_bp(sis_oracle *o, int model, int min_loop, int max_bp_per_nt, int nstep_MC,
                     double cutoff_energy, const double *p)
{
    o->bp_model = model; o->bp_min_loop = min_loop;
    o->max_bp_per_nt = max_bp_per_nt; o->nstep_bp_MC = nstep_MC;
    o->bp_cutoff_energy = cutoff_energy;
    for (int t = 0; t < 3; t++) {
        bp_para *b = &o->bp_paras[1 + t];
        const double *q = p + 15 * t;
        b->cutoff_ddist = 0.0;
        b->U0 = q[0]; b->bond_k = q[1]; b->bond_r = q[2];
        b->angl_k1 = q[3]; b->angl_k2 = q[4]; b->angl_k3 = q[5]; b->angl_k4 = q[6];
        b->angl_theta1 = q[7]; b->angl_theta2 = q[8]; b->angl_theta3 = q[9]; b->angl_theta4 = q[10];
        b->dihd_k1 = q[11]; b->dihd_k2 = q[12]; b->dihd_phi1 = q[13]; b->dihd_phi2 = q[14];
        bp_para *s = &o->bp_paras[4 + t]; /* CG, UA, UG */
        *s = *b;
        s->angl_k1 = b->angl_k2; s->angl_k2 = b->angl_k1;
        s->angl_k3 = b->angl_k4; s->angl_k4 = b->angl_k3;
    }
    return 0;
}

int oracle_set_ff_nn(sis_oracle *o, const double *NN_dH, const double *NN_dS, double dH0,
                     double dS0, int flg_NNend, const double *NNend_dH, const double *NNend_dS,
                     double dHend0, double dSend0)
{
    for (int i = 0; i < 21; i++) {
        o->NN_dH[i + 1] = NN_dH[i];
        o->NN_dS[i + 1] = NN_dS[i];
    }
    o->dH0 = dH0; o->dS0 = dS0; o->flg_NNend = flg_NNend;
    for (int i = 0; i < 6; i++) {
        o->NNend_dH[i + 1] = NNend_dH ? NNend_dH[i] : 0.0;
        o->NNend_dS[i + 1] = NNend_dS ? NNend_dS[i] : 0.0;
    }
    o->dHend0 = dHend0; o->dSend0 = dSend0;
    o->have_nn = 1;
    return 0;
}

int oracle_set_ele(sis_oracle *o, int flg_ele, int cutoff_type, double cutoff_inp, int excl_bond,
                   int excl_angle, double length_per_charge, int temp_independent,
                   double tempK_ref)
{
    o->flg_ele = flg_ele; o->ele_cutoff_type = cutoff_type; o->ele_cutoff_inp = cutoff_inp;
    o->excl_bond = excl_bond; o->excl_angle = excl_angle;
    o->length_per_charge = length_per_charge;
    o->temp_independent = temp_independent; o->tempK_ref = tempK_ref;
    return 0;
}

int oracle_set_pbc(sis_oracle *o, int flg_pbc, const double *box)
{
    o->flg_pbc = flg_pbc;
    for (int i = 0; i < 3; i++) {
        o->box[i] = box ? box[i] : 0.0;
        o->box_half[i] = 0.5 * o->box[i];
    }
    return 0;
}

int oracle_set_md(sis_oracle *o, int integrator, double dt, double viscosity_Pas,
                  double nl_margin, uint64_t seed)
{
    o->integrator = integrator; o->dt = dt; o->viscosity_Pas = viscosity_Pas;
    o->nl_margin = nl_margin; o->seed = seed;
    return 0;
}

int oracle_set_bias_ss(sis_oracle *o, int flg, double force)
{
    o->flg_bias_ss = flg; o->bias_ss_force = force;
    return 0;
}
int oracle_set_stage(sis_oracle *o, int flg, double sigma, double eps)
{
    o->flg_stage = flg; o->stage_sigma = sigma; o->stage_eps = eps;
    return 0;
}
int oracle_set_tweezers_dcf(sis_oracle *o, int n, const int *pairs, const double *direction)
{
    o->ntwz_DCF = n;
    o->twz_DCF_pairs = dup_int(pairs, 2 * n);
    o->twz_DCF_direction = dup_dbl(direction, 3 * n);
    return 0;
}
int oracle_set_tweezers_fr(sis_oracle *o, int n, const int *pairs, const double *k,
                           const double *init, const double *velo)
{
    o->ntwz_FR = n;
    o->twz_FR_pairs = dup_int(pairs, 2 * n);
    o->twz_FR_k = dup_dbl(k, 2 * n);
    o->twz_FR_init = dup_dbl(init, 6 * n);
    o->twz_FR_velo = dup_dbl(velo, 6 * n);
    return 0;
}
int oracle_set_bias_rg(sis_oracle *o, int flg, int pott, double k)
{
    o->flg_bias_rg = flg; o->bias_rg_pott = pott; o->bias_rg_k = k;
    return 0;
}
int oracle_set_restraints(sis_oracle *o, int n, const int *id, const double *rcut,
                          const double *para)
{
    o->nrest = n;
    o->rest_id = dup_int(id, 3 * n);
    o->rest_rcut = dup_dbl(rcut, n);
    o->rest_para = dup_dbl(para, 3 * n);
    return 0;
}

/* ---- set_ele: src/set_ele.F90:47-102 (non-_HTN_CONSISTENT branch) ------------ */
static void set_ele(const sis_oracle *o, replica *r, double tempk, double ionstr, double *out_lb,
                    double *out_Zp)
{
    const double MM_A = 87.740, MM_B = -0.40008, MM_C = 9.398e-4, MM_D = -1.410e-6;
    double Tc, lb, Zp;
    if (o->temp_independent == 0) {
        Tc = tempk - 273.15;
        r->diele = MM_A + MM_B * Tc + MM_C * Tc * Tc + MM_D * Tc * Tc * Tc;
    } else {
        Tc = o->tempK_ref - 273.15;
        r->diele = MM_A + MM_B * Tc + MM_C * Tc * Tc + MM_D * Tc * Tc * Tc;
        r->diele_dTcoef = 1.0 + o->tempK_ref / r->diele * (MM_B + 2.0 * MM_C * Tc + 3.0 * MM_D * Tc * Tc);
    }
    lb = ELE * ELE / (4.0 * O_PI * EPS0 * r->diele * BOLTZ_J * tempk) * 1.0e10;
    Zp = o->length_per_charge / lb;
    r->ele_coef = JOUL2KCAL_MOL * 1.0e10 * ELE * ELE / (4.0 * O_PI * EPS0 * r->diele);
    r->ele_coef = r->ele_coef * Zp * Zp;
    r->lambdaD = 1.0e10 * sqrt((1.0e-3 * EPS0 * r->diele * BOLTZ_J) / (2.0 * N_AVO * ELE * ELE)) *
                 sqrt(tempk / ionstr);
    if (out_lb) *out_lb = lb;
    if (out_Zp) *out_Zp = Zp;
}

int oracle_set_ele_calc(sis_oracle *o, double tempK, double ionstr, double out[6])
{
    replica tmp;
    memset(&tmp, 0, sizeof tmp);
    double lb, Zp;
    set_ele(o, &tmp, tempK, ionstr, &lb, &Zp);
    out[0] = tmp.diele; out[1] = lb; out[2] = Zp; out[3] = tmp.ele_coef; out[4] = tmp.lambdaD;
    out[5] = tmp.diele_dTcoef;
    return 0;
}

int oracle_get_ele(sis_oracle *o, int irep, double out[4])
{
    replica *r = &o->rep[irep];
    out[0] = r->ele_coef; out[1] = r->lambdaD; out[2] = r->ele_cutoff; out[3] = r->diele;
    return 0;
}

static void ensure_replicas(sis_oracle *o)
{
    if (o->rep) return;
    o->rep = (replica *)calloc((size_t)o->nrep_proc, sizeof(replica));
    for (int i = 0; i < o->nrep_proc; i++) {
        replica *r = &o->rep[i];
        r->xyz = dup_dbl(NULL, 3 * o->nmp);
        r->velos = dup_dbl(NULL, 3 * o->nmp);
        r->accels = dup_dbl(NULL, 3 * o->nmp);
        r->xyz_move = dup_dbl(NULL, 3 * o->nmp);
        r->tempK = 300.0;
        r->ionic_strength = 0.15;
    }
}

/* src/init_ele.F90:64-101 + src/job_md.F90:538-567 (per-replica refresh) */
int oracle_set_replica_params(sis_oracle *o, int irep, double tempK, double ionic_strength,
                              double twz_force, double rg0)
{
    ensure_replicas(o);
    if (irep < 0 || irep >= o->nrep_proc) return fail("irep out of range");
    replica *r = &o->rep[irep];
    r->tempK = tempK;
    r->ionic_strength = ionic_strength;
    r->bias_rg_0 = rg0;
    r->twz_force = twz_force;
    if (o->flg_ele) {
        set_ele(o, r, tempK, ionic_strength, NULL, NULL);
        if (o->ele_cutoff_type == 1) r->ele_cutoff = o->ele_cutoff_inp;
        else if (r->ele_cutoff == 0.0) r->ele_cutoff = o->ele_cutoff_inp * r->lambdaD; /* init_ele.F90:88-90 */
    }
    if (o->ntwz_DCF > 0) {
        if (!r->twz_DCF_forces) r->twz_DCF_forces = dup_dbl(NULL, 3 * o->ntwz_DCF);
        for (int k = 0; k < 3 * o->ntwz_DCF; k++) r->twz_DCF_forces[k] = twz_force * o->twz_DCF_direction[k];
    }
    return 0;
}

/* ---- src/const_idx.F90:343-369 is_complement -------------------------------- */
static int is_complement(int s1, int s2)
{
    if (s1 == SEQ_A) return s2 == SEQ_U;
    if (s1 == SEQ_U) return s2 == SEQ_A || s2 == SEQ_G;
    if (s1 == SEQ_G) return s2 == SEQ_C || s2 == SEQ_U;
    if (s1 == SEQ_C) return s2 == SEQ_G;
    return 0;
}

/* src/const_idx.F90:231-279,330-341: NN stack 5'-ij-3' / 3'-kl-5' -> NNT index 1..21 */
static int seqt2nnt(int i, int j, int k, int l)
{
    static const char S[4] = {'A', 'U', 'G', 'C'};
    static const struct { const char *a, *b; } T[21] = {
        {"GC_CG", NULL}, {"CC_GG", "GG_CC"}, {"GA_CU", "UC_AG"}, {"CG_GC", NULL},
        {"AC_UG", "GU_CA"}, {"CA_GU", "UG_AC"}, {"AG_UC", "CU_GA"}, {"UA_AU", NULL},
        {"AU_UA", NULL}, {"AA_UU", "UU_AA"}, {"GC_UG", "GU_CG"}, {"CU_GG", "GG_UC"},
        {"GG_CU", "UC_GG"}, {"CG_GU", "UG_GC"}, {"AU_UG", "GU_UA"}, {"GA_UU", "UU_AG"},
        {"UG_GU", NULL}, {"UA_GU", "UG_AU"}, {"GG_UU", "UU_GG"}, {"GU_UG", NULL},
        {"AG_UU", "UU_GA"}};
    char c[6] = {S[i], S[j], '_', S[k], S[l], 0};
    for (int n = 0; n < 21; n++)
        if (!strcmp(c, T[n].a) || (T[n].b && !strcmp(c, T[n].b))) return n + 1;
    return 0;
}

/* src/init_bp.F90:352-378 */
static int bp3_hash_key(int w, int x, int u, int z, int y, int v)
{
    int a, b, c, d, e, f;
    if (x < y) { a = w; b = x; c = u; d = z; e = y; f = v; }
    else { f = w; e = x; d = u; c = z; b = y; a = v; }
    return 1 + a + b * 4 + c * 16 + d * 64 + e * 256 + f * 1024;
}

enum { AUonAU = 1, AUonCG, AUonGU, GUonCG, GUonAU, GUonGU };
static int nnend_type(int x, int y, int w, int z)
{
    int au = (x == SEQ_A && y == SEQ_U) || (x == SEQ_U && y == SEQ_A);
    int gu = (x == SEQ_G && y == SEQ_U) || (x == SEQ_U && y == SEQ_G);
    int on_au = (w == SEQ_A && z == SEQ_U) || (w == SEQ_U && z == SEQ_A);
    int on_cg = (w == SEQ_G && z == SEQ_C) || (w == SEQ_C && z == SEQ_G);
    int on_gu = (w == SEQ_G && z == SEQ_U) || (w == SEQ_U && z == SEQ_G);
    if (au) return on_au ? AUonAU : on_cg ? AUonCG : on_gu ? AUonGU : 0;
    if (gu) return on_au ? GUonAU : on_cg ? GUonCG : on_gu ? GUonGU : 0;
    return 0;
}

/* src/init_bp.F90:39-172: the 468-entry triplet table */
static void init_bp3_table(sis_oracle *o)
{
    memset(o->bp3_hash, 0, sizeof o->bp3_hash);
    memset(o->bp3_dH, 0, sizeof o->bp3_dH);
    memset(o->bp3_dS, 0, sizeof o->bp3_dS);
    int h = 0;
    for (int x = 0; x <= 2; x++)
        for (int y = x + 1; y <= 3; y++) {
            if (!is_complement(x, y)) continue;
            for (int w = 0; w <= 3; w++)
                for (int z = 0; z <= 3; z++) {
                    int comp_wz = is_complement(w, z);
                    double dH = 0.0, dS = 0.0;
                    if (comp_wz) {
                        dH = 0.5 * (o->NN_dH[seqt2nnt(w, x, z, y)] - o->dH0);
                        dS = 0.5 * (o->NN_dS[seqt2nnt(w, x, z, y)] - o->dS0);
                    }
                    for (int u = 0; u <= 3; u++)
                        for (int v = 0; v <= 3; v++) {
                            int comp_uv = is_complement(u, v);
                            if (!comp_wz && !comp_uv) continue;
                            h++;
                            o->bp3_hash[bp3_hash_key(w, x, u, z, y, v)] = h;
                            if (comp_uv) {
                                o->bp3_dH[h] = dH + 0.5 * (o->NN_dH[seqt2nnt(x, u, y, v)] - o->dH0);
                                o->bp3_dS[h] = (dS + 0.5 * (o->NN_dS[seqt2nnt(x, u, y, v)] - o->dS0)) * 1.0e-3;
                            } else {
                                o->bp3_dH[h] = dH;
                                o->bp3_dS[h] = dS * 1.0e-3;
                            }
                            if (o->flg_NNend) {
                                int t = 0;
                                if (comp_wz && !comp_uv) t = nnend_type(x, y, w, z);
                                else if (!comp_wz && comp_uv) t = nnend_type(x, y, u, v);
                                if (t) {
                                    o->bp3_dH[h] += o->NNend_dH[t] - o->dHend0;
                                    o->bp3_dS[h] += (o->NNend_dS[t] - o->dSend0) * 1.0e-3;
                                }
                            }
                        }
                }
        }
    o->bp3_n = h; /* must be 468: src/init_bp.F90:170-171 */
}

/* src/init_bp.F90:184-313 evaluated on the fly for one pair (0-based imp < jmp):
 * returns bp_map (type 1..6) and bp3_map (>0 if the pair enters the bp list) */
static int ss_has(const sis_oracle *o, int imp, int jmp)
{
    const long long key = (long long)imp * o->nmp + jmp;
    int lo = 0, hi = o->n_ss;
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (o->ss_keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo < o->n_ss && o->ss_keys[lo] == key;
}
static int bp_type_seq(int a, int b)
{
    if (a == SEQ_G && b == SEQ_C) return BPT_GC;
    if (a == SEQ_C && b == SEQ_G) return BPT_CG;
    if (a == SEQ_A && b == SEQ_U) return BPT_AU;
    if (a == SEQ_U && b == SEQ_A) return BPT_UA;
    if (a == SEQ_G && b == SEQ_U) return BPT_GU;
    if (a == SEQ_U && b == SEQ_G) return BPT_UG;
    return 0;
}

static void bp_lookup(const sis_oracle *o, int imp, int jmp, int *bptype, int *bp3)
{
    *bptype = 0;
    *bp3 = 0;
    if (imp >= jmp) return;
    if (o->flg_ss) {
        /* src/init_bp.F90:207-209: only the pairs read_ss left in bp_map are visited.  For
         * bp_model /= 5, bp3_map = bp_map (:315-317) keeps every such pair in the lists, whatever
         * the loop `cycle`s over afterwards */
        if (!ss_has(o, imp, jmp)) return;
        if (o->bp_model != 5) {
            const int t = bp_type_seq(o->seq[imp], o->seq[jmp]);
            *bptype = t;
            *bp3 = t;
            return;
        }
    }
    int ichain = o->ichain_mp[imp], jchain = o->ichain_mp[jmp];
    int i = o->lmp_mp[imp] + 1, j = o->lmp_mp[jmp] + 1; /* 1-based as in the reference */
    int ni = o->nmp_chain[ichain], nj = o->nmp_chain[jchain];
    int i_pre = i - 1, i_nxt = i + 1, j_pre = j - 1, j_nxt = j + 1;
    int i_circ = o->moltypes[ichain] == MOLT_CIRC, j_circ = o->moltypes[jchain] == MOLT_CIRC;
    if (i_circ) {
        if (i == 1) i_pre = ni;
        if (i == ni) i_nxt = 1;
    } else if (i == 1 || i == ni) return;
    if (j_circ) {
        if (j == 1) j_pre = nj;
        if (j == nj) j_nxt = 1;
    } else if (j == 1 || j == nj) return;
    const int *si = o->seq + o->chain_start[ichain] - 1; /* si[1..ni] */
    const int *sj = o->seq + o->chain_start[jchain] - 1;
    int pair_ipre_jnxt = is_complement(si[i_pre], sj[j_nxt]);
    int pair_inxt_jpre = is_complement(si[i_nxt], sj[j_pre]);
    if (!i_circ && i_pre == 1) pair_ipre_jnxt = 0;
    if (!j_circ && j_nxt == nj) pair_ipre_jnxt = 0;
    if (!i_circ && i_nxt == ni) pair_inxt_jpre = 0;
    if (!j_circ && j_pre == 1) pair_inxt_jpre = 0;
    if (ichain == jchain) {
        if (i + o->bp_min_loop >= j) return;
        if (i + 1 + o->bp_min_loop >= j - 1) pair_inxt_jpre = 0;
        if (i_circ) {
            if (j + o->bp_min_loop >= i + ni) return;
            if (j + 1 + o->bp_min_loop >= i - 1 + ni) pair_ipre_jnxt = 0;
        }
    }
    if (o->bp_model == 3 || o->bp_model == 4 || o->bp_model == 5)
        if (!pair_ipre_jnxt && !pair_inxt_jpre) return;
    int a = si[i], b = sj[j], t = 0;
    if (a == SEQ_G && b == SEQ_C) t = BPT_GC;
    else if (a == SEQ_C && b == SEQ_G) t = BPT_CG;
    else if (a == SEQ_A && b == SEQ_U) t = BPT_AU;
    else if (a == SEQ_U && b == SEQ_A) t = BPT_UA;
    else if (a == SEQ_G && b == SEQ_U) t = BPT_GU;
    else if (a == SEQ_U && b == SEQ_G) t = BPT_UG;
    *bptype = t;
    if (o->bp_model == 5) {
        /* a D (=4) in the centre would index bp3_hash out of its 4**6 range in the
         * reference (undefined behaviour); the oracle defines that as "no pair". */
        if (a > 3 || b > 3) return;
        int key;
        if (!pair_ipre_jnxt) key = bp3_hash_key(SEQ_U, a, si[i_nxt], SEQ_U, b, sj[j_pre]);
        else if (!pair_inxt_jpre) key = bp3_hash_key(si[i_pre], a, SEQ_U, sj[j_nxt], b, SEQ_U);
        else key = bp3_hash_key(si[i_pre], a, si[i_nxt], sj[j_nxt], b, sj[j_pre]);
        *bp3 = (key >= 1 && key <= 4096) ? o->bp3_hash[key] : 0;
    } else {
        *bp3 = t; /* src/init_bp.F90:315-317 */
    }
}

/* src/read_ss.F90:120-210: the pairs of a ct / bpseq / bpl file that enter bp_map.  pairs = 1-based
 * (imp, jmp) in any order; n < 0 removes the restriction.  Call after set_topology and set_ff_bp. */
static int cmp_ll(const void *a, const void *b)
{
    const long long x = *(const long long *)a, y = *(const long long *)b;
    return x < y ? -1 : x > y;
}
int oracle_set_bp_restrict(sis_oracle *o, int n, const int *pairs)
{
    free(o->ss_keys);
    o->ss_keys = NULL; o->n_ss = 0; o->flg_ss = n >= 0;
    if (n <= 0) return 0;
    o->ss_keys = (long long *)malloc((size_t)n * sizeof(long long));
    for (int k = 0; k < n; k++) {
        int imp = pairs[2 * k] - 1, jmp = pairs[2 * k + 1] - 1;
        if (imp > jmp) { const int t = imp; imp = jmp; jmp = t; }
        if (imp < 0 || jmp >= o->nmp || imp == jmp) continue;
        const int ichain = o->ichain_mp[imp], jchain = o->ichain_mp[jmp];
        const int i = o->lmp_mp[imp] + 1, j = o->lmp_mp[jmp] + 1;
        if (i == 1 || i == o->nmp_chain[ichain]) continue;               /* chain ends */
        if (j == 1 || j == o->nmp_chain[jchain]) continue;
        if (ichain == jchain && i + o->bp_min_loop >= j) continue;         /* minimum loop length */
        if (o->bp_model == 3 || o->bp_model == 4 || o->bp_model == 5)      /* isolated base pair */
            if (!is_complement(o->seq[imp - 1], o->seq[jmp + 1]) && !is_complement(o->seq[imp + 1], o->seq[jmp - 1])) continue;
        if (!bp_type_seq(o->seq[imp], o->seq[jmp])) continue;              /* no known type */
        o->ss_keys[o->n_ss++] = (long long)imp * o->nmp + jmp;
    }
    qsort(o->ss_keys, (size_t)o->n_ss, sizeof(long long), cmp_ll);
    int m = 0;
    for (int k = 0; k < o->n_ss; k++)
        if (k == 0 || o->ss_keys[k] != o->ss_keys[k - 1]) o->ss_keys[m++] = o->ss_keys[k];
    o->n_ss = m;
    return 0;
}

int oracle_bp_lookup(sis_oracle *o, int imp, int jmp, int *bptype, int *bp3idx)
{
    bp_lookup(o, imp - 1, jmp - 1, bptype, bp3idx);
    return 0;
}

int oracle_bp3_table(sis_oracle *o, int *n, int *hash4096, double *dH, double *dS)
{
    *n = o->bp3_n;
    if (hash4096) memcpy(hash4096, o->bp3_hash + 1, 4096 * sizeof(int));
    if (dH) memcpy(dH, o->bp3_dH + 1, 468 * sizeof(double));
    if (dS) memcpy(dS, o->bp3_dS + 1, 468 * sizeof(double));
    return 0;
}

/* src/list_local.F90:16-119 */
static void list_local(sis_oracle *o)
{
    int has_dih = o->flg_dih_cos || o->flg_dih_exp;
    for (int n = 1; n <= 2; n++) {
        int ibond = 0, iangl = 0, idih = 0;
        for (int c = 0; c < o->nchains; c++) {
            int len = o->nmp_chain[c];
            for (int i = 1; i <= len - 1; i++) {
                int imp = o->chain_start[c] + i - 1;
                ibond++;
                if (i != len - 1) iangl++;
                if (i < len - 2) idih++;
                if (n == 2) {
                    o->bond_mp[2 * (ibond - 1)] = imp;
                    o->bond_mp[2 * (ibond - 1) + 1] = imp + 1;
                    if (i != len - 1) {
                        int *a = o->angl_mp + 3 * (iangl - 1);
                        a[0] = imp; a[1] = imp + 1; a[2] = imp + 2;
                    }
                    if (has_dih && i < len - 2) {
                        int *d = o->dih_mp + 4 * (idih - 1);
                        d[0] = imp; d[1] = imp + 1; d[2] = imp + 2; d[3] = imp + 3;
                    }
                }
            }
            if (o->moltypes[c] == MOLT_CIRC) {
                if (n == 1) { ibond += 1; iangl += 2; idih += 3; }
                else {
                    int ini = o->chain_start[c], end = ini + len - 1;
                    ibond++;
                    o->bond_mp[2 * (ibond - 1)] = end;
                    o->bond_mp[2 * (ibond - 1) + 1] = ini;
                    int *a;
                    iangl++; a = o->angl_mp + 3 * (iangl - 1);
                    a[0] = end - 1; a[1] = end; a[2] = ini;
                    iangl++; a = o->angl_mp + 3 * (iangl - 1);
                    a[0] = end; a[1] = ini; a[2] = ini + 1;
                    if (has_dih) {
                        int *d;
                        idih++; d = o->dih_mp + 4 * (idih - 1);
                        d[0] = end - 2; d[1] = end - 1; d[2] = end; d[3] = ini;
                        idih++; d = o->dih_mp + 4 * (idih - 1);
                        d[0] = end - 1; d[1] = end; d[2] = ini; d[3] = ini + 1;
                        idih++; d = o->dih_mp + 4 * (idih - 1);
                        d[0] = end; d[1] = ini; d[2] = ini + 1; d[3] = ini + 2;
                    } else idih += 3;
                }
            }
        }
        if (n == 1) {
            o->nbond = ibond; o->nangl = iangl; o->ndih = has_dih ? idih : 0;
            o->bond_mp = (int *)xcalloc(2 * (size_t)ibond, sizeof(int));
            o->angl_mp = (int *)xcalloc(3 * (size_t)iangl, sizeof(int));
            o->dih_mp = (int *)xcalloc(4 * (size_t)idih, sizeof(int));
        }
    }
}

int oracle_get_local_lists(sis_oracle *o, int *nbond, int *nangl, int *ndih, int *bond_mp,
                           int *angl_mp, int *dih_mp)
{
    *nbond = o->nbond; *nangl = o->nangl; *ndih = o->ndih;
    if (bond_mp) for (int i = 0; i < 2 * o->nbond; i++) bond_mp[i] = o->bond_mp[i] + 1;
    if (angl_mp) for (int i = 0; i < 3 * o->nangl; i++) angl_mp[i] = o->angl_mp[i] + 1;
    if (dih_mp) for (int i = 0; i < 4 * o->ndih; i++) dih_mp[i] = o->dih_mp[i] + 1;
    return 0;
}

int oracle_finalize_setup(sis_oracle *o)
{
    if (!o->nmp) return fail("topology not set");
    ensure_replicas(o);
    list_local(o);
    if (o->bp_model == 5) {
        if (!o->have_nn) return fail("bp model 5 needs NN tables");
        init_bp3_table(o);
        if (o->bp3_n != 468) return fail("bp3 table has %d entries, expected 468", o->bp3_n);
    }
    /* src/init_bp.F90:319-346 */
    if (fabs(o->bp_cutoff_energy) <= SMALL_VALUE) {
        o->bp_cutoff_dist = 18.0;
        for (int t = 1; t <= 6; t++) o->bp_paras[t].cutoff_ddist = o->bp_cutoff_dist - 13.8;
    } else {
        o->bp_cutoff_dist = 0.0;
        for (int t = 1; t <= 6; t++) {
            bp_para *b = &o->bp_paras[t];
            b->cutoff_ddist = sqrt(log(fabs(10.0 / o->bp_cutoff_energy)) / b->bond_k);
            double cutoff = b->bond_r + b->cutoff_ddist;
            if (cutoff > o->bp_cutoff_dist) o->bp_cutoff_dist = cutoff;
        }
    }
    o->finalized = 1;
    return 0;
}

/* ---- state ------------------------------------------------------------------- */
int oracle_set_state(sis_oracle *o, int irep, const double *xyz, const double *velos,
                     const double *accels)
{
    ensure_replicas(o);
    replica *r = &o->rep[irep];
    size_t n = 3 * (size_t)o->nmp * sizeof(double);
    if (xyz) memcpy(r->xyz, xyz, n);
    if (velos) memcpy(r->velos, velos, n);
    if (accels) memcpy(r->accels, accels, n);
    return 0;
}
int oracle_get_state(sis_oracle *o, int irep, double *xyz, double *velos, double *accels)
{
    replica *r = &o->rep[irep];
    size_t n = 3 * (size_t)o->nmp * sizeof(double);
    if (xyz) memcpy(xyz, r->xyz, n);
    if (velos) memcpy(velos, r->velos, n);
    if (accels) memcpy(accels, r->accels, n);
    return 0;
}
int oracle_set_step(sis_oracle *o, int64_t istep) { o->istep = istep; return 0; }

/* ---- lists ------------------------------------------------------------------- */
static void grow_bp(replica *r, int need)
{
    if (need <= r->nbp_max) return;
    int m = r->nbp_max ? r->nbp_max : 16;
    while (m < need) m = (int)(m * 1.2) + 1; /* src/neighbor_list.F90:298 growth factor */
    r->bp_mp = (int *)realloc(r->bp_mp, 7 * (size_t)m * sizeof(int));
    r->bp_coef = (double *)realloc(r->bp_coef, 2 * (size_t)m * sizeof(double));
    r->bp_status = (int *)realloc(r->bp_status, (size_t)m * sizeof(int));
    r->bp_status_MC = (int *)realloc(r->bp_status_MC, (size_t)m * sizeof(int));
    r->ene_bp = (double *)realloc(r->ene_bp, (size_t)m * sizeof(double));
    r->for_bp = (double *)realloc(r->for_bp, 18 * (size_t)m * sizeof(double));
    for (int i = r->nbp_max; i < m; i++) {
        r->bp_status[i] = 0; r->bp_status_MC[i] = 0; r->ene_bp[i] = 0.0;
    }
    r->nbp_max = m;
}
static void push_pair(int **list, int *n, int *nmax, int a, int b)
{
    if (*n >= *nmax) {
        *nmax = *nmax ? (int)(*nmax * 1.2) + 1 : 1024;
        *list = (int *)realloc(*list, 2 * (size_t)*nmax * sizeof(int));
    }
    (*list)[2 * *n] = a;
    (*list)[2 * *n + 1] = b;
    (*n)++;
}

/* src/neighbor_list.F90:28-233 */
int oracle_neighbor_list(sis_oracle *o, int irep)
{
    replica *r = &o->rep[irep];
    if (o->flg_pbc) pbc_wrap(o, r);
    double ele_nl_cut2 = 0.0;
    if (o->flg_ele) {
        if (o->ele_cutoff_type == 2) r->ele_cutoff = o->ele_cutoff_inp * r->lambdaD; /* :80-82 */
        ele_nl_cut2 = (r->ele_cutoff + o->nl_margin) * (r->ele_cutoff + o->nl_margin);
    }
    double wca_nl_cut2 = (o->wca_sigma + o->nl_margin) * (o->wca_sigma + o->nl_margin);
    double bp_nl_cut2 = o->flg_bias_ss ? 9999999999.9
                                       : (o->bp_cutoff_dist + o->nl_margin) * (o->bp_cutoff_dist + o->nl_margin);
    int iwca = 0, iele = 0, ibp = 0;
    /* the reference zeroes bp state on (re)allocation only; a rebuild leaves
     * bp_status / bp_status_MC entries in place by index (src/neighbor_list.F90:42-63) */
    for (int ichain = 0; ichain < o->nchains; ichain++)
        for (int jchain = ichain; jchain < o->nchains; jchain++)
            for (int i = 1; i <= o->nmp_chain[ichain]; i++) {
                int imp = o->chain_start[ichain] + i - 1;
                int iseq = o->seq[imp];
                int j_start = (ichain == jchain) ? i + 1 : 1;
                for (int j = j_start; j <= o->nmp_chain[jchain]; j++) {
                    int jmp = o->chain_start[jchain] + j - 1;
                    int jseq = o->seq[jmp];
                    double v[3];
                    pbc_vec_d(o, r->xyz + 3 * imp, r->xyz + 3 * jmp, v);
                    double d2 = dot3(v, v);
                    /* WCA :122-148 */
                    if (iseq != SEQ_D && jseq != SEQ_D && d2 <= wca_nl_cut2) {
                        int add = 1;
                        if (ichain == jchain) {
                            if (i + 2 >= j) add = 0;
                            else if (o->moltypes[ichain] == MOLT_CIRC)
                                if (j + 2 >= i + o->nmp_chain[ichain]) add = 0;
                        }
                        if (add) push_pair(&r->wca_mp, &iwca, &r->nwca_max, imp, jmp);
                    }
                    /* electrostatics :150-180 */
                    if (o->flg_ele && d2 <= ele_nl_cut2 && o->has_charge[imp] && o->has_charge[jmp]) {
                        int add = 1;
                        if (ichain == jchain) {
                            int len = o->nmp_chain[ichain];
                            if (j == i + 1 && o->excl_bond) add = 0;
                            if (j == i + 2 && o->excl_angle) add = 0;
                            if (o->moltypes[ichain] == MOLT_CIRC) {
                                if (i == 1) {
                                    if (j == len && o->excl_bond) add = 0;
                                    if (j == len - 1 && o->excl_angle) add = 0;
                                } else if (i == 2) {
                                    if (j == len && o->excl_angle) add = 0;
                                }
                            }
                        }
                        if (add) push_pair(&r->ele_mp, &iele, &r->nele_max, imp, jmp);
                    }
                    /* base pairs :182-223 */
                    if (d2 <= bp_nl_cut2) {
                        int bptype, bp3;
                        bp_lookup(o, imp, jmp, &bptype, &bp3);
                        if (bp3 > 0) {
                            grow_bp(r, ibp + 1);
                            int *b = r->bp_mp + 7 * ibp;
                            int ni = o->nmp_chain[ichain], nj = o->nmp_chain[jchain];
                            b[0] = imp; b[1] = jmp;
                            b[2] = (i == 1) ? imp + ni - 1 : imp - 1;
                            b[3] = (j == 1) ? jmp + nj - 1 : jmp - 1;
                            b[4] = (i == ni) ? imp + 1 - ni : imp + 1;
                            /* :212 tests `j == nmp_chain(ichain)`; for chains of unequal length that
                             * indexes a bead outside chain j (out of bounds in the worst case).
                             * Identical to the intended nmp_chain(jchain) whenever the chain
                             * lengths agree or the chains are linear of equal length (every
                             * check/ input and C1-C5).  Oracle and GPU use the intended form;
                             * DESIGN.md lists this as the one deliberate deviation. */
                            b[5] = (j == nj) ? jmp + 1 - nj : jmp + 1;
                            b[6] = bptype;
                            if (o->bp_model == 5) {
                                r->bp_coef[2 * ibp] = o->bp3_dH[bp3];
                                r->bp_coef[2 * ibp + 1] = o->bp3_dS[bp3];
                            } else {
                                r->bp_coef[2 * ibp] = 0.0;
                                r->bp_coef[2 * ibp + 1] = 0.0;
                            }
                            ibp++;
                        }
                    }
                }
            }
    r->nwca = iwca; r->nbp = ibp;
    if (o->flg_ele) r->nele = iele;
    return 0;
}

/* src/list_exv.F90:19-64, src/list_ele.F90:18-65, src/list_bp.F90:18-93 */
int oracle_full_lists(sis_oracle *o)
{
    for (int irep = 0; irep < o->nrep_proc; irep++) {
        replica *r = &o->rep[irep];
        int iwca = 0, iele = 0, ibp = 0;
        for (int ichain = 0; ichain < o->nchains; ichain++)
            for (int jchain = ichain; jchain < o->nchains; jchain++)
                for (int i = 1; i <= o->nmp_chain[ichain]; i++) {
                    int imp = o->chain_start[ichain] + i - 1;
                    if (o->seq[imp] == SEQ_D) continue;
                    int j_start = (ichain == jchain) ? i + 3 : 1;
                    for (int j = j_start; j <= o->nmp_chain[jchain]; j++) {
                        int jmp = o->chain_start[jchain] + j - 1;
                        if (o->seq[jmp] == SEQ_D) continue;
                        push_pair(&r->wca_mp, &iwca, &r->nwca_max, imp, jmp);
                    }
                }
        r->nwca = iwca;
        if (o->flg_ele) {
            for (int ichain = 0; ichain < o->nchains; ichain++)
                for (int jchain = ichain; jchain < o->nchains; jchain++)
                    for (int i = 1; i <= o->nmp_chain[ichain]; i++) {
                        int imp = o->chain_start[ichain] + i - 1;
                        if (!o->has_charge[imp]) continue;
                        int j_start = (ichain == jchain) ? i + 1 : 1;
                        for (int j = j_start; j <= o->nmp_chain[jchain]; j++) {
                            /* as in the reference: tested for inter-chain pairs too */
                            if (j == i + 1 && o->excl_bond) continue;
                            if (j == i + 2 && o->excl_angle) continue;
                            int jmp = o->chain_start[jchain] + j - 1;
                            if (!o->has_charge[jmp]) continue;
                            push_pair(&r->ele_mp, &iele, &r->nele_max, imp, jmp);
                        }
                    }
            r->nele = iele;
            if (o->ele_cutoff_type == 2 && r->ele_cutoff == 0.0) r->ele_cutoff = o->ele_cutoff_inp * r->lambdaD;
        }
        for (int imp = 0; imp < o->nmp; imp++)
            for (int jmp = imp + o->bp_min_loop + 1; jmp < o->nmp; jmp++) {
                int bptype, bp3;
                bp_lookup(o, imp, jmp, &bptype, &bp3);
                if (bp3 == 0) continue;
                grow_bp(r, ibp + 1);
                int *b = r->bp_mp + 7 * ibp;
                int ichain = o->ichain_mp[imp], jchain = o->ichain_mp[jmp];
                int ni = o->nmp_chain[ichain], nj = o->nmp_chain[jchain];
                int i_circ = o->moltypes[ichain] == MOLT_CIRC, j_circ = o->moltypes[jchain] == MOLT_CIRC;
                b[0] = imp; b[1] = jmp;
                b[2] = (i_circ && o->lmp_mp[imp] == 0) ? imp + ni - 1 : imp - 1;
                b[3] = (j_circ && o->lmp_mp[jmp] == 0) ? jmp + nj - 1 : jmp - 1;
                b[4] = (i_circ && o->lmp_mp[imp] == ni - 1) ? imp - ni + 1 : imp + 1;
                b[5] = (j_circ && o->lmp_mp[jmp] == nj - 1) ? jmp - nj + 1 : jmp + 1;
                b[6] = bptype;
                r->bp_coef[2 * ibp] = (o->bp_model == 5) ? o->bp3_dH[bp3] : 0.0;
                r->bp_coef[2 * ibp + 1] = (o->bp_model == 5) ? o->bp3_dS[bp3] : 0.0;
                ibp++;
            }
        r->nbp = ibp;
        for (int k = 0; k < ibp; k++) { r->bp_status[k] = 0; r->bp_status_MC[k] = 0; r->ene_bp[k] = 0.0; }
    }
    return 0;
}

int oracle_get_list_sizes(sis_oracle *o, int irep, int *nwca, int *nele, int *nbp)
{
    replica *r = &o->rep[irep];
    *nwca = r->nwca; *nele = o->flg_ele ? r->nele : 0; *nbp = r->nbp;
    return 0;
}

int oracle_get_lists(sis_oracle *o, int irep, int *wca_mp, int *ele_mp, int *bp_mp, double *bp_coef)
{
    replica *r = &o->rep[irep];
    if (wca_mp) for (int k = 0; k < 2 * r->nwca; k++) wca_mp[k] = r->wca_mp[k] + 1;
    if (ele_mp && o->flg_ele) for (int k = 0; k < 2 * r->nele; k++) ele_mp[k] = r->ele_mp[k] + 1;
    if (bp_mp)
        for (int k = 0; k < r->nbp; k++) {
            for (int c = 0; c < 6; c++) bp_mp[7 * k + c] = r->bp_mp[7 * k + c] + 1;
            bp_mp[7 * k + 6] = r->bp_mp[7 * k + 6];
        }
    if (bp_coef) memcpy(bp_coef, r->bp_coef, 2 * (size_t)r->nbp * sizeof(double));
    return 0;
}

/* ---- bonded forces ----------------------------------------------------------- */
/* src/force_bond.F90:34-56 */
static void force_bond(const sis_oracle *o, const replica *r, double *F)
{
#pragma omp for nowait
    for (int ibd = 0; ibd < o->nbond; ibd++) {
        int i1 = o->bond_mp[2 * ibd], i2 = o->bond_mp[2 * ibd + 1];
        double v[3];
        pbc_vec_d(o, r->xyz + 3 * i1, r->xyz + 3 * i2, v);
        double d = sqrt(dot3(v, v));
        double delta = d - o->bond_r0;
        for (int k = 0; k < 3; k++) {
            double f = o->bond_k * delta / d * v[k];
            F[3 * i1 + k] -= f;
            F[3 * i2 + k] += f;
        }
    }
}
/* src/force_angl.F90:36-82 */
static void force_angl(const sis_oracle *o, const replica *r, double *F)
{
#pragma omp for nowait
    for (int ia = 0; ia < o->nangl; ia++) {
        int i1 = o->angl_mp[3 * ia], i2 = o->angl_mp[3 * ia + 1], i3 = o->angl_mp[3 * ia + 2];
        double v21[3], v32[3];
        pbc_vec_d(o, r->xyz + 3 * i2, r->xyz + 3 * i1, v21);
        pbc_vec_d(o, r->xyz + 3 * i3, r->xyz + 3 * i2, v32);
        double d21 = dot3(v21, v21), d32 = dot3(v32, v32), d2132 = dot3(v32, v21);
        double cosine = -d2132 / sqrt(d21 * d32);
        if (fabs(cosine) > 1.0) cosine = copysign(1.0, cosine);
        double delta = acos(cosine) - o->angl_t0;
        double t = d21 * d32 - d2132 * d2132;
        if (t <= 1.0) t = 1.0;
        double f = o->angl_k * delta / sqrt(t);
        for (int k = 0; k < 3; k++) {
            double f21 = f * (v21[k] * (d2132 / d21) - v32[k]);
            double f32 = f * (v32[k] * (d2132 / d32) - v21[k]);
            F[3 * i1 + k] -= f21;
            F[3 * i2 + k] += f21 - f32;
            F[3 * i3 + k] += f32;
        }
    }
}
/* src/force_angl_ReB.F90:36-73 */
static void force_angl_ReB(const sis_oracle *o, const replica *r, double *F)
{
    double co0 = cos(o->angl_t0);
#pragma omp for nowait
    for (int ia = 0; ia < o->nangl; ia++) {
        int i1 = o->angl_mp[3 * ia], i2 = o->angl_mp[3 * ia + 1], i3 = o->angl_mp[3 * ia + 2];
        double v12[3], v32[3];
        pbc_vec_d(o, r->xyz + 3 * i1, r->xyz + 3 * i2, v12);
        pbc_vec_d(o, r->xyz + 3 * i3, r->xyz + 3 * i2, v32);
        double d12 = dot3(v12, v12), d32 = dot3(v32, v32);
        double a1232 = sqrt(d12 * d32);
        double cosine = dot3(v12, v32) / a1232;
        double om = 1.0 - cosine * cosine;
        double f = -o->angl_k * (cosine - co0) * (1.0 - cosine * co0) / (om * om);
        for (int k = 0; k < 3; k++) {
            double f1 = f * (v32[k] / a1232 - cosine * v12[k] / d12);
            double f3 = f * (v12[k] / a1232 - cosine * v32[k] / d32);
            F[3 * i1 + k] += f1;
            F[3 * i2 + k] -= (f1 + f3);
            F[3 * i3 + k] += f3;
        }
    }
}
/* src/force_dih_cos.F90:21-59 and src/force_dih_exp.F90:36-94 share the geometry */
static void force_dih(const sis_oracle *o, const replica *r, double *F, int exp_form)
{
    double coef1 = o->dih_k * o->dih_w, coef2 = -0.5 * o->dih_w;
#pragma omp for nowait
    for (int id = 0; id < o->ndih; id++) {
        const int *q = o->dih_mp + 4 * id;
        double vij[3], vkj[3], vkl[3], m[3], n[3];
        pbc_vec_d(o, r->xyz + 3 * q[0], r->xyz + 3 * q[1], vij);
        pbc_vec_d(o, r->xyz + 3 * q[2], r->xyz + 3 * q[1], vkj);
        pbc_vec_d(o, r->xyz + 3 * q[2], r->xyz + 3 * q[3], vkl);
        cross3(vij, vkj, m);
        cross3(vkj, vkl, n);
        double akj2 = dot3(vkj, vkj), akj = sqrt(akj2);
        double dih = atan2(akj * dot3(vij, n), dot3(m, n));
        double pre;
        if (exp_form) {
            double dd = dih - o->dih_p0;
            pre = coef1 * dd * exp(coef2 * dd * dd) * akj;
        } else pre = -o->dih_k * sin(dih + o->dih_p0) * akj;
        double mm = dot3(m, m), nn = dot3(n, n);
        double a = dot3(vij, vkj) / akj2, b = dot3(vkl, vkj) / akj2;
        for (int k = 0; k < 3; k++) {
            double fi = -pre / mm * m[k];
            double fl = pre / nn * n[k];
            double fj = (-1.0 + a) * fi + (-b) * fl;
            double fk = (-a) * fi + (-1.0 + b) * fl;
            F[3 * q[0] + k] += fi;
            F[3 * q[1] + k] += fj;
            F[3 * q[2] + k] += fk;
            F[3 * q[3] + k] += fl;
        }
    }
}

/* ---- non-bonded -------------------------------------------------------------- */
/* src/force_wca.F90:34-66 */
static void force_wca(const sis_oracle *o, const replica *r, double *F)
{
    double s2 = o->wca_sigma * o->wca_sigma;
    double coef = 12.0 * o->wca_eps / s2;
#pragma omp for nowait
    for (int k = 0; k < r->nwca; k++) {
        int i1 = r->wca_mp[2 * k], i2 = r->wca_mp[2 * k + 1];
        double v[3];
        pbc_vec_d(o, r->xyz + 3 * i1, r->xyz + 3 * i2, v);
        double d2 = dot3(v, v);
        if (d2 >= s2) continue;
        d2 = s2 / d2;
        double c = coef * (powi(d2, 7) - powi(d2, 4));
        for (int a = 0; a < 3; a++) {
            F[3 * i1 + a] += c * v[a];
            F[3 * i2 + a] -= c * v[a];
        }
    }
}
/* src/force_ele_DH.F90:39-81 */
static void force_ele_DH(const sis_oracle *o, const replica *r, double *F)
{
    double cutoff2 = r->ele_cutoff * r->ele_cutoff;
    double rcdist = 1.0 / r->lambdaD;
#pragma omp for nowait
    for (int k = 0; k < r->nele; k++) {
        int i1 = r->ele_mp[2 * k], i2 = r->ele_mp[2 * k + 1];
        double v21[3];
        pbc_vec_d(o, r->xyz + 3 * i2, r->xyz + 3 * i1, v21);
        double dist2 = dot3(v21, v21);
        if (dist2 > cutoff2) continue;
        double dist1 = sqrt(dist2), rdist1 = 1.0 / dist1;
        double g = r->ele_coef * rdist1 * rdist1 * (rdist1 + rcdist) * exp(-dist1 * rcdist);
        for (int a = 0; a < 3; a++) {
            F[3 * i1 + a] -= g * v21[a];
            F[3 * i2 + a] += g * v21[a];
        }
    }
}

/* ---- base pairs -------------------------------------------------------------- */
/* Force-side evaluation shared by src/force_bp.F90:50-204, force_bp_triplet.F90:62-240 and
 * pass 1 of force_bp_limit_triplet.F90:85-257.  Returns 0 if the distance gate fails.
 * f[6][3] = d(u)/dr contributions before scaling by the energy; *u_out = exponent. */
static int bp_eval_force(const sis_oracle *o, const replica *r, const int *b, const bp_para *p,
                         double f[6][3], double *u_out, double *a12_out, double v12_out[3])
{
    const double *x1 = r->xyz + 3 * b[0], *x2 = r->xyz + 3 * b[1], *x3 = r->xyz + 3 * b[2];
    const double *x4 = r->xyz + 3 * b[3], *x5 = r->xyz + 3 * b[4], *x6 = r->xyz + 3 * b[5];
    double v12[3], v13[3], v42[3], v15[3], v62[3], m[3], n[3], fi[3], fk[3], fl[3];
    pbc_vec_d(o, x1, x2, v12);
    double d1212 = dot3(v12, v12), a12 = sqrt(d1212);
    double d = a12 - p->bond_r;
    *a12_out = a12;
    memcpy(v12_out, v12, sizeof v12);
    if (fabs(d) > p->cutoff_ddist) return 0;
    double u = p->bond_k * d * d;
    memset(f, 0, 18 * sizeof(double));
    for (int k = 0; k < 3; k++) {
        fi[k] = (2.0 * p->bond_k * d / a12) * v12[k];
        f[0][k] = fi[k];
        f[1][k] = -fi[k];
    }
    pbc_vec_d(o, x1, x3, v13);
    pbc_vec_d(o, x1, x5, v15);
    pbc_vec_d(o, x4, x2, v42);
    pbc_vec_d(o, x6, x2, v62);
    double d1313 = dot3(v13, v13), d4242 = dot3(v42, v42), d1213 = dot3(v12, v13);
    double d1242 = dot3(v12, v42), d1215 = dot3(v12, v15), d1515 = dot3(v15, v15);
    double d1262 = dot3(v12, v62), d6262 = dot3(v62, v62);
    double r1213 = d1213 / d1212, r1242 = d1242 / d1212, r1215 = d1215 / d1212, r1262 = d1262 / d1212;
    double cosine, pre;
    /* angle 3-1=2 */
    cosine = d1213 / (sqrt(d1313) * a12);
    d = acos(cosine) - p->angl_theta2;
    pre = 2.0 * p->angl_k2 * d / sqrt(d1313 * d1212 - d1213 * d1213);
    u += p->angl_k2 * d * d;
    for (int k = 0; k < 3; k++) {
        fi[k] = pre * (v12[k] - (d1213 / d1313 * v13[k]));
        fk[k] = pre * (v13[k] - (r1213 * v12[k]));
        f[0][k] += -fi[k] - fk[k];
        f[1][k] += fk[k];
        f[2][k] = fi[k];
    }
    /* angle 1=2-4 */
    cosine = d1242 / (a12 * sqrt(d4242));
    d = acos(cosine) - p->angl_theta1;
    u += p->angl_k1 * d * d;
    pre = 2.0 * p->angl_k1 * d / sqrt(d1212 * d4242 - d1242 * d1242);
    for (int k = 0; k < 3; k++) {
        fi[k] = -pre * (v42[k] - (r1242 * v12[k]));
        fk[k] = -pre * (v12[k] - (d1242 / d4242 * v42[k]));
        f[0][k] += fi[k];
        f[1][k] += -fi[k] - fk[k];
        f[3][k] = fk[k];
    }
    /* angle 5-1=2 */
    cosine = d1215 / (sqrt(d1515) * a12);
    d = acos(cosine) - p->angl_theta4;
    u += p->angl_k4 * d * d;
    pre = 2.0 * p->angl_k4 * d / sqrt(d1515 * d1212 - d1215 * d1215);
    for (int k = 0; k < 3; k++) {
        fi[k] = pre * (v12[k] - (d1215 / d1515 * v15[k]));
        fk[k] = pre * (v15[k] - (r1215 * v12[k]));
        f[0][k] += -fi[k] - fk[k];
        f[1][k] += fk[k];
        f[4][k] = fi[k];
    }
    /* angle 1=2-6 */
    cosine = d1262 / (a12 * sqrt(d6262));
    d = acos(cosine) - p->angl_theta3;
    u += p->angl_k3 * d * d;
    pre = 2.0 * p->angl_k3 * d / sqrt(d1212 * d6262 - d1262 * d1262);
    for (int k = 0; k < 3; k++) {
        fi[k] = -pre * (v62[k] - (r1262 * v12[k]));
        fk[k] = -pre * (v12[k] - (d1262 / d6262 * v62[k]));
        f[0][k] += fi[k];
        f[1][k] += -fi[k] - fk[k];
        f[5][k] = fk[k];
    }
    /* dihedral 4-2=1-3 */
    cross3(v42, v12, m);
    cross3(v12, v13, n);
    double dih = atan2(dot3(v42, n) * a12, dot3(m, n));
    d = dih + p->dihd_phi1;
    u += p->dihd_k1 * (1.0 + cos(d));
    pre = -p->dihd_k1 * sin(d) * a12;
    double mm = dot3(m, m), nn = dot3(n, n);
    for (int k = 0; k < 3; k++) {
        fi[k] = pre / mm * m[k];
        fl[k] = -pre / nn * n[k];
        f[3][k] += fi[k];
        f[1][k] += (-1.0 + r1242) * fi[k] - r1213 * fl[k];
        f[0][k] += (-1.0 + r1213) * fl[k] - r1242 * fi[k];
        f[2][k] += fl[k];
    }
    /* dihedral 6-2=1-5 */
    cross3(v62, v12, m);
    cross3(v12, v15, n);
    dih = atan2(dot3(v62, n) * a12, dot3(m, n));
    d = dih + p->dihd_phi2;
    u += p->dihd_k2 * (1.0 + cos(d));
    pre = -p->dihd_k2 * sin(d) * a12;
    mm = dot3(m, m); nn = dot3(n, n);
    for (int k = 0; k < 3; k++) {
        fi[k] = pre / mm * m[k];
        fl[k] = -pre / nn * n[k];
        f[5][k] += fi[k];
        f[1][k] += (-1.0 + r1262) * fi[k] - r1215 * fl[k];
        f[0][k] += (-1.0 + r1215) * fl[k] - r1262 * fi[k];
        f[4][k] += fl[k];
    }
    *u_out = u;
    return 1;
}

/* energy-side helpers: src/energy_bp.F90:76-118 (clamped acos) */
static double mp_angle(const sis_oracle *o, const replica *r, int i1, int i2, int i3)
{
    double v12[3], v32[3];
    pbc_vec_d(o, r->xyz + 3 * i1, r->xyz + 3 * i2, v12);
    pbc_vec_d(o, r->xyz + 3 * i3, r->xyz + 3 * i2, v32);
    double co = dot3(v32, v12) / sqrt(dot3(v12, v12) * dot3(v32, v32));
    if (co > 1.0) co = 1.0;
    else if (co < -1.0) co = -1.0;
    return acos(co);
}
static double mp_dihedral(const sis_oracle *o, const replica *r, int i1, int i2, int i3, int i4)
{
    double v12[3], v32[3], v34[3], m[3], n[3];
    pbc_vec_d(o, r->xyz + 3 * i1, r->xyz + 3 * i2, v12);
    pbc_vec_d(o, r->xyz + 3 * i3, r->xyz + 3 * i2, v32);
    pbc_vec_d(o, r->xyz + 3 * i3, r->xyz + 3 * i4, v34);
    cross3(v12, v32, m);
    cross3(v32, v34, n);
    return atan2(dot3(v12, n) * sqrt(dot3(v32, v32)), dot3(m, n));
}
/* exponent u of src/energy_bp.F90:40-62 / energy_bp_triplet.F90:55-77; returns 0 if gated */
static int bp_eval_energy(const sis_oracle *o, const replica *r, const int *b, const bp_para *p,
                          double *u_out)
{
    double v[3];
    pbc_vec_d(o, r->xyz + 3 * b[0], r->xyz + 3 * b[1], v);
    double d = sqrt(dot3(v, v)) - p->bond_r;
    if (fabs(d) > p->cutoff_ddist) return 0;
    double u = p->bond_k * d * d, t;
    t = mp_angle(o, r, b[0], b[1], b[3]); u += p->angl_k1 * (t - p->angl_theta1) * (t - p->angl_theta1);
    t = mp_angle(o, r, b[2], b[0], b[1]); u += p->angl_k2 * (t - p->angl_theta2) * (t - p->angl_theta2);
    t = mp_angle(o, r, b[0], b[1], b[5]); u += p->angl_k3 * (t - p->angl_theta3) * (t - p->angl_theta3);
    t = mp_angle(o, r, b[4], b[0], b[1]); u += p->angl_k4 * (t - p->angl_theta4) * (t - p->angl_theta4);
    t = mp_dihedral(o, r, b[2], b[0], b[1], b[3]); u += p->dihd_k1 * (1.0 + cos(t + p->dihd_phi1));
    t = mp_dihedral(o, r, b[4], b[0], b[1], b[5]); u += p->dihd_k2 * (1.0 + cos(t + p->dihd_phi2));
    *u_out = u;
    return 1;
}

static int is_triplet(const sis_oracle *o) { return o->bp_model == 4 || o->bp_model == 5; }

/* counter-based uniform for the MC pruning / exchange (replaces mt_stream; streams differ) */
double oracle_uniform(uint64_t seed, uint32_t stream, uint64_t counter);

/* pass 2 of src/force_bp_limit_triplet.F90:260-328 (also energy_bp_limit_triplet.F90:108-190) */
static void bp_mc_prune(const sis_oracle *o, replica *r, int irep, int *nt_bp_excess, double beta,
                        int deterministic)
{
    int nmp = o->nmp;
    int *ntlist = (int *)xcalloc(nmp, sizeof(int));
    int *bp_seq = (int *)xcalloc(r->nbp, sizeof(int));
    for (;;) {
        int nnt = 0;
        for (int imp = 0; imp < nmp; imp++)
            if (nt_bp_excess[imp] > 0) ntlist[nnt++] = imp;
        if (nnt == 0) break;
        double rnd = oracle_uniform(o->seed, 0x40000000u + (uint32_t)irep, r->mc_counter++);
        int pick = (int)ceil(rnd * nnt);
        if (pick < 1) pick = 1;
        if (pick > nnt) pick = nnt;
        int nt_delete = ntlist[pick - 1];
        int nseq = 0;
        for (int ibp = 0; ibp < r->nbp; ibp++)
            if (r->bp_status[ibp] && (r->bp_mp[7 * ibp] == nt_delete || r->bp_mp[7 * ibp + 1] == nt_delete))
                bp_seq[nseq++] = ibp;
        int ibp_delete;
        if (!deterministic) {
            for (int i = 0; i < nseq; i++) {
                rnd = oracle_uniform(o->seed, 0x40000000u + (uint32_t)irep, r->mc_counter++);
                int sw = (int)ceil(rnd * nseq);
                if (sw < 1) sw = 1;
                if (sw > nseq) sw = nseq;
                int t = bp_seq[i]; bp_seq[i] = bp_seq[sw - 1]; bp_seq[sw - 1] = t;
            }
            ibp_delete = bp_seq[0];
            for (int i = 1; i < nseq; i++) {
                int jbp = bp_seq[i];
                double ratio = exp((r->ene_bp[jbp] - r->ene_bp[ibp_delete]) * beta);
                rnd = oracle_uniform(o->seed, 0x40000000u + (uint32_t)irep, r->mc_counter++);
                if (rnd < ratio) ibp_delete = jbp;
            }
        } else { /* temp_independent branch: energy_bp_limit_triplet.F90:167-176 */
            ibp_delete = bp_seq[0];
            for (int i = 1; i < nseq; i++)
                if (r->ene_bp[bp_seq[i]] > r->ene_bp[ibp_delete]) ibp_delete = bp_seq[i];
        }
        r->bp_status[ibp_delete] = 0;
        nt_bp_excess[r->bp_mp[7 * ibp_delete]]--;
        nt_bp_excess[r->bp_mp[7 * ibp_delete + 1]]--;
    }
    free(ntlist);
    free(bp_seq);
}

/* src/force_bp.F90, force_bp_triplet.F90, force_bp_limit_triplet.F90 */
static void force_bp_all(const sis_oracle *o, replica *r, int irep, double *F)
{
    /* called by every thread of the parallel region of oracle_force_terms (orphaned worksharing,
     * as the reference's force_bp* routines are): master resets + barrier
     * (src/force_bp_triplet.F90:49-55, force_bp_limit_triplet.F90:70-78), `omp do` over the list,
     * serial Monte-Carlo pruning on the master (:259-333), `omp do` scatter of the survivors */
    int triplet = is_triplet(o), limit = o->flg_bp_MC;
    double tK = r->tempK, beta = 1.0 / (BOLTZ_KCAL_MOL * tK);
    int *excess = o->bp_excess;
#pragma omp master
    {
        for (int k = 0; k < r->nbp; k++) { r->bp_status[k] = 0; r->ene_bp[k] = 0.0; }
        if (limit) {
            memset(r->for_bp, 0, 18 * (size_t)r->nbp * sizeof(double));
            for (int i = 0; i < o->nmp; i++) excess[i] = -o->max_bp_per_nt;
        }
    }
#pragma omp barrier
#pragma omp for nowait
    for (int ibp = 0; ibp < r->nbp; ibp++) {
        if (!limit && o->nstep_bp_MC > 0 && !r->bp_status_MC[ibp]) continue;
        const int *b = r->bp_mp + 7 * ibp;
        const bp_para *p = &o->bp_paras[b[6]];
        double dG = 1.0;
        if (triplet) {
            dG = r->bp_coef[2 * ibp] - tK * r->bp_coef[2 * ibp + 1];
            if (dG >= 0.0) continue;
        }
        double f[6][3], u, a12, v12[3];
        if (!bp_eval_force(o, r, b, p, f, &u, &a12, v12)) {
            if (triplet && o->flg_bias_ss) /* force_bp_triplet.F90:85-93 */
                for (int k = 0; k < 3; k++) {
                    double t = o->bias_ss_force / a12 * v12[k];
                    F[3 * b[0] + k] -= t;
                    F[3 * b[1] + k] += t;
                }
            continue;
        }
        double ene = triplet ? p->U0 * dG * exp(-u) : p->U0 * exp(-u);
        if (ene <= o->bp_cutoff_energy) {
            r->bp_status[ibp] = 1;
            r->ene_bp[ibp] = ene;
            if (limit) {
                for (int s = 0; s < 6; s++)
                    for (int k = 0; k < 3; k++) r->for_bp[18 * ibp + 3 * s + k] = ene * f[s][k];
#pragma omp atomic
                excess[b[0]]++;
#pragma omp atomic
                excess[b[1]]++;
            } else
                for (int s = 0; s < 6; s++)
                    for (int k = 0; k < 3; k++) F[3 * b[s] + k] += ene * f[s][k];
        }
    }
    if (limit) {
#pragma omp barrier
#pragma omp master
        bp_mc_prune(o, r, irep, excess, beta, 0);
#pragma omp barrier
#pragma omp for nowait
        for (int ibp = 0; ibp < r->nbp; ibp++) {
            if (!r->bp_status[ibp]) continue;
            const int *b = r->bp_mp + 7 * ibp;
            for (int s = 0; s < 6; s++)
                for (int k = 0; k < 3; k++) F[3 * b[s] + k] += r->for_bp[18 * ibp + 3 * s + k];
        }
    }
#pragma omp master
    r->flg_bp_energy = 1;
}

/* src/energy_bp.F90, energy_bp_triplet.F90, energy_bp_limit_triplet.F90 */
static double energy_bp_all(const sis_oracle *o, replica *r, int irep, double tempK_in)
{
    int triplet = is_triplet(o), limit = o->flg_bp_MC;
    if (!r->flg_bp_energy) {
        double tK = (o->temp_independent == 0) ? tempK_in : 0.0;
        double beta = (o->temp_independent == 0) ? 1.0 / (BOLTZ_KCAL_MOL * tK) : 0.0;
        int *excess = NULL;
        if (limit) {
            excess = (int *)xcalloc(o->nmp, sizeof(int));
            for (int i = 0; i < o->nmp; i++) excess[i] = -o->max_bp_per_nt;
        }
        for (int k = 0; k < r->nbp; k++) { r->bp_status[k] = 0; r->ene_bp[k] = 0.0; }
#pragma omp parallel for
        for (int ibp = 0; ibp < r->nbp; ibp++) {
            if (!limit && o->nstep_bp_MC > 0 && !r->bp_status_MC[ibp]) continue;
            const int *b = r->bp_mp + 7 * ibp;
            const bp_para *p = &o->bp_paras[b[6]];
            double dG = 1.0, u;
            if (triplet) {
                dG = r->bp_coef[2 * ibp] - tK * r->bp_coef[2 * ibp + 1];
                if (dG >= 0.0) continue;
            }
            if (!bp_eval_energy(o, r, b, p, &u)) continue;
            u = triplet ? p->U0 * dG * exp(-u) : p->U0 * exp(-u);
            if (u <= o->bp_cutoff_energy) {
                r->ene_bp[ibp] = u;
                r->bp_status[ibp] = 1;
                if (limit) {
#pragma omp atomic
                    excess[b[0]]++;
#pragma omp atomic
                    excess[b[1]]++;
                }
            }
        }
        if (limit) {
            bp_mc_prune(o, r, irep, excess, beta, o->temp_independent != 0);
            free(excess);
        }
    }
    double e = 0.0;
    for (int k = 0; k < r->nbp; k++)
        if (!limit || r->bp_status[k]) e += r->ene_bp[k]; /* limit: masked sum (:195) */
    return e;
}

/* ---- biases ------------------------------------------------------------------ */
/* src/force_stage.F90:17-37 */
static void force_stage(const sis_oracle *o, const replica *r, double *F)
{
    double coef = 24.0 * o->stage_eps / o->stage_sigma;
#pragma omp for nowait
    for (int imp = 0; imp < o->nmp; imp++) {
        double z = r->xyz[3 * imp + 2];
        if (z < SMALL_VALUE) z = SMALL_VALUE;
        z = o->stage_sigma / z;
        F[3 * imp + 2] += coef * (2 * powi(z, 13) - powi(z, 7));
    }
}
/* src/force_tweezers.F90:27-61 */
static void force_tweezers(const sis_oracle *o, const replica *r, double *F)
{
    for (int p = 0; p < o->ntwz_DCF; p++) {
        int i1 = o->twz_DCF_pairs[2 * p] - 1, i2 = o->twz_DCF_pairs[2 * p + 1] - 1;
        for (int k = 0; k < 3; k++) {
            F[3 * i1 + k] += r->twz_DCF_forces[3 * p + k];
            F[3 * i2 + k] -= r->twz_DCF_forces[3 * p + k];
        }
    }
    for (int p = 0; p < o->ntwz_FR; p++) {
        int i1 = o->twz_FR_pairs[2 * p] - 1, i2 = o->twz_FR_pairs[2 * p + 1] - 1;
        double p1[3], p2[3], v1[3], v2[3];
        for (int k = 0; k < 3; k++) {
            p1[k] = o->twz_FR_init[6 * p + k] + o->twz_FR_velo[6 * p + k] * (double)o->istep;
            p2[k] = o->twz_FR_init[6 * p + 3 + k] + o->twz_FR_velo[6 * p + 3 + k] * (double)o->istep;
        }
        pbc_vec_d(o, p1, r->xyz + 3 * i1, v1);
        pbc_vec_d(o, p2, r->xyz + 3 * i2, v2);
        for (int k = 0; k < 3; k++) {
            F[3 * i1 + k] += o->twz_FR_k[2 * p] * v1[k];
            F[3 * i2 + k] += o->twz_FR_k[2 * p + 1] * v2[k];
        }
    }
}
static double calc_rg(const sis_oracle *o, const replica *r, double cm[3])
{
    cm[0] = cm[1] = cm[2] = 0.0;
    for (int imp = 0; imp < o->nmp; imp++)
        for (int k = 0; k < 3; k++) cm[k] += r->xyz[3 * imp + k];
    for (int k = 0; k < 3; k++) cm[k] /= (double)o->nmp;
    double s2 = 0.0;
    for (int imp = 0; imp < o->nmp; imp++) {
        double v[3];
        for (int k = 0; k < 3; k++) v[k] = r->xyz[3 * imp + k] - cm[k];
        s2 += dot3(v, v);
    }
    return sqrt(s2 / (double)o->nmp);
}
/* src/force_rg.F90:18-39 */
static void force_rg(const sis_oracle *o, const replica *r, double *F)
{
    double cm[3];
    double rg = calc_rg(o, r, cm);
    if (o->bias_rg_pott == 2 && rg <= r->bias_rg_0) return;
    double factor = -o->bias_rg_k * (rg - r->bias_rg_0) / (rg * o->nmp);
    for (int imp = 0; imp < o->nmp; imp++)
        for (int k = 0; k < 3; k++) F[3 * imp + k] += factor * (r->xyz[3 * imp + k] - cm[k]);
}
/* src/force_restraint.F90:19-41 */
static void force_restraint(const sis_oracle *o, const replica *r, double *F)
{
    for (int ir = 0; ir < o->nrest; ir++) {
        int imp = o->rest_id[3 * ir] - 1, jmp = o->rest_id[3 * ir + 1] - 1, itype = o->rest_id[3 * ir + 2];
        double v[3];
        pbc_vec_d(o, r->xyz + 3 * imp, r->xyz + 3 * jmp, v);
        double rr = sqrt(dot3(v, v));
        if (rr >= o->rest_rcut[ir]) continue;
        double eps = o->rest_para[3 * ir], d = o->rest_para[3 * ir + 1], s = o->rest_para[3 * ir + 2];
        double th = tanh((rr - d) / s);
        double coef = 0.5 * eps / s * (1.0 - th * th) / rr;
        for (int k = 0; k < 3; k++) {
            F[3 * imp + k] -= coef * v[k];
            if (itype == 1) F[3 * jmp + k] += coef * v[k];
        }
    }
}

/* ---- force dispatcher: src/force.F90:27-75 ----------------------------------- */
int oracle_set_bp_mc(sis_oracle *o, int flg) { o->flg_bp_MC = flg; return 0; }

int oracle_force_terms(sis_oracle *o, int irep, unsigned mask, double *F)
{
    replica *r = &o->rep[irep];
    const size_t n3 = 3 * (size_t)o->nmp;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    if (o->forces_t_n < n3 * (size_t)nthreads) {   /* forces_t(3, nmp, 0:nthreads-1), src/force.F90:19 */
        free(o->forces_t);
        o->forces_t = (double *)xcalloc(n3 * (size_t)nthreads, sizeof(double));
        o->forces_t_n = n3 * (size_t)nthreads;
    }
    if (!o->bp_excess) o->bp_excess = (int *)xcalloc(o->nmp, sizeof(int));
    double *forces_t = o->forces_t;
#pragma omp parallel
    {
        int tn = 0;
#ifdef _OPENMP
        tn = omp_get_thread_num();
#endif
        double *Ft = forces_t + n3 * (size_t)tn;
        memset(Ft, 0, n3 * sizeof(double));
        if (mask & (1u << O_ENE_BOND)) force_bond(o, r, Ft);
        if (mask & (1u << O_ENE_ANGL)) {
            if (o->flg_angl_ReB) force_angl_ReB(o, r, Ft);
            else force_angl(o, r, Ft);
        }
        if (mask & (1u << O_ENE_DIHE)) {
            if (o->flg_dih_cos) force_dih(o, r, Ft, 0);
            if (o->flg_dih_exp) force_dih(o, r, Ft, 1);
        }
        if (mask & (1u << O_ENE_BP)) force_bp_all(o, r, irep, Ft);
        if (mask & (1u << O_ENE_EXV)) force_wca(o, r, Ft);
        if ((mask & (1u << O_ENE_ELE)) && o->flg_ele) force_ele_DH(o, r, Ft);
        if ((mask & (1u << O_ENE_STAGE)) && o->flg_stage) force_stage(o, r, Ft);
    }
    /* serial reduction over the threads, src/force.F90:66-69 */
    for (int tn = 1; tn < nthreads; tn++)
        for (size_t i = 0; i < n3; i++) forces_t[i] += forces_t[n3 * (size_t)tn + i];
    memcpy(F, forces_t, n3 * sizeof(double));
    /* serial biases, src/force.F90:71-75 */
    if ((mask & (1u << O_ENE_TWZ)) && (o->ntwz_DCF + o->ntwz_FR) > 0) force_tweezers(o, r, F);
    if ((mask & (1u << O_ENE_RG)) && o->flg_bias_rg) force_rg(o, r, F);
    if ((mask & (1u << O_ENE_REST)) && o->nrest > 0) force_restraint(o, r, F);
    return 0;
}
int oracle_force(sis_oracle *o, int irep, double *F) { return oracle_force_terms(o, irep, 0x7FEu, F); }

/* ---- energies ---------------------------------------------------------------- */
static double energy_bond(const sis_oracle *o, const replica *r)
{ /* src/energy_bond.F90:17-28 */
    double e = 0.0;
    for (int ibd = 0; ibd < o->nbond; ibd++) {
        double v[3];
        pbc_vec_d(o, r->xyz + 3 * o->bond_mp[2 * ibd], r->xyz + 3 * o->bond_mp[2 * ibd + 1], v);
        double d = sqrt(dot3(v, v));
        e += 0.5 * o->bond_k * (d - o->bond_r0) * (d - o->bond_r0);
    }
    return e;
}
static double energy_angl(const sis_oracle *o, const replica *r)
{ /* src/energy_angl.F90:18-41, energy_angl_ReB.F90:18-29 */
    double e = 0.0, co0 = cos(o->angl_t0);
    for (int ia = 0; ia < o->nangl; ia++) {
        const int *a = o->angl_mp + 3 * ia;
        double v12[3], v32[3];
        pbc_vec_d(o, r->xyz + 3 * a[0], r->xyz + 3 * a[1], v12);
        pbc_vec_d(o, r->xyz + 3 * a[2], r->xyz + 3 * a[1], v32);
        double t = dot3(v32, v12) / sqrt(dot3(v12, v12) * dot3(v32, v32));
        if (o->flg_angl_ReB) e += 0.5 * o->angl_k * (t - co0) * (t - co0) / (1.0 - t * t);
        else {
            if (t > 1.0) t = 1.0;
            else if (t < -1.0) t = -1.0;
            t = acos(t);
            e += 0.5 * o->angl_k * (t - o->angl_t0) * (t - o->angl_t0);
        }
    }
    return e;
}
static double energy_dih(const sis_oracle *o, const replica *r)
{ /* src/energy_dih_cos.F90:18-39, energy_dih_exp.F90:20-41 */
    double e = 0.0, coef = -0.5 * o->dih_w;
    for (int id = 0; id < o->ndih; id++) {
        const int *q = o->dih_mp + 4 * id;
        double v12[3], v23[3], v34[3], c1[3], c2[3];
        pbc_vec_d(o, r->xyz + 3 * q[1], r->xyz + 3 * q[0], v12);
        pbc_vec_d(o, r->xyz + 3 * q[2], r->xyz + 3 * q[1], v23);
        pbc_vec_d(o, r->xyz + 3 * q[3], r->xyz + 3 * q[2], v34);
        cross3(v12, v23, c1);
        cross3(v23, v34, c2);
        double dih = atan2(sqrt(dot3(v23, v23)) * dot3(v12, c2), dot3(c1, c2));
        if (o->flg_dih_cos) e += o->dih_k * (1.0 + cos(dih + o->dih_p0));
        if (o->flg_dih_exp) e -= o->dih_k * exp(coef * (dih - o->dih_p0) * (dih - o->dih_p0));
    }
    return e;
}
static double energy_wca(const sis_oracle *o, const replica *r)
{ /* src/energy_wca.F90:19-35 */
    double e = 0.0;
#pragma omp parallel for reduction(+ : e)
    for (int k = 0; k < r->nwca; k++) {
        double v[3];
        pbc_vec_d(o, r->xyz + 3 * r->wca_mp[2 * k], r->xyz + 3 * r->wca_mp[2 * k + 1], v);
        double d = sqrt(dot3(v, v));
        if (d >= o->wca_sigma) continue;
        d = o->wca_sigma / d;
        e += o->wca_eps * (powi(d, 12) - 2 * powi(d, 6) + 1.0);
    }
    return e;
}
static double energy_ele_DH(const sis_oracle *o, const replica *r)
{ /* src/energy_ele_DH.F90:20-40 */
    double e = 0.0, rcdist = 1.0 / r->lambdaD;
#pragma omp parallel for reduction(+ : e)
    for (int k = 0; k < r->nele; k++) {
        double v[3];
        pbc_vec_d(o, r->xyz + 3 * r->ele_mp[2 * k], r->xyz + 3 * r->ele_mp[2 * k + 1], v);
        double dist = sqrt(dot3(v, v));
        if (dist > r->ele_cutoff) continue;
        if (o->temp_independent == 0) e += r->ele_coef / dist * exp(-dist * rcdist);
        else {
            double rk = dist * rcdist;
            e += r->ele_coef / dist * exp(-rk) * (-(1.0 + 0.5 * rk) * r->diele_dTcoef);
        }
    }
    return e;
}
static int energy_stage(const sis_oracle *o, const replica *r, double *e_out)
{ /* src/energy_stage.F90:22-39 */
    double e = 0.0;
    for (int imp = 0; imp < o->nmp; imp++)
        if (r->xyz[3 * imp + 2] < 0.85 * o->stage_sigma)
            return fail("Z-coordinate < 0.85 * stage_sigma in energy_stage");
#pragma omp parallel for reduction(+ : e)
    for (int imp = 0; imp < o->nmp; imp++) {
        double z = o->stage_sigma / r->xyz[3 * imp + 2];
        e += 4 * o->stage_eps * (powi(z, 12) - powi(z, 6));
    }
    *e_out = e;
    return 0;
}
static double energy_tweezers(const sis_oracle *o, const replica *r)
{ /* src/energy_tweezers.F90:21-31 (DCF only) */
    double e = 0.0;
    for (int p = 0; p < o->ntwz_DCF; p++) {
        double v[3];
        pbc_vec_d(o, r->xyz + 3 * (o->twz_DCF_pairs[2 * p] - 1), r->xyz + 3 * (o->twz_DCF_pairs[2 * p + 1] - 1), v);
        e -= dot3(v, r->twz_DCF_forces + 3 * p);
    }
    return e;
}
static double energy_rg(const sis_oracle *o, replica *r)
{ /* src/energy_rg.F90:18-35 */
    double cm[3];
    r->rg = calc_rg(o, r, cm);
    if (o->bias_rg_pott == 2 && r->rg <= r->bias_rg_0) return 0.0;
    return 0.5 * o->bias_rg_k * (r->rg - r->bias_rg_0) * (r->rg - r->bias_rg_0);
}
static double energy_restraint(const sis_oracle *o, const replica *r)
{ /* src/energy_restraint.F90:19-33 */
    double e = 0.0;
    for (int ir = 0; ir < o->nrest; ir++) {
        double v[3];
        pbc_vec_d(o, r->xyz + 3 * (o->rest_id[3 * ir] - 1), r->xyz + 3 * (o->rest_id[3 * ir + 1] - 1), v);
        double rr = sqrt(dot3(v, v));
        if (rr >= o->rest_rcut[ir]) continue;
        e -= 0.5 * o->rest_para[3 * ir] * (1.0 - tanh((rr - o->rest_para[3 * ir + 1]) / o->rest_para[3 * ir + 2]));
    }
    return e;
}

/* src/energy_sumup.F90:15-54.  recompute_bp != 0 sets flg_bp_energy = .False. first
 * (job_md.F90:289, energy_replica.F90:72, job_check_force.F90:55). */
int oracle_energy_sumup(sis_oracle *o, int irep, double tempK, int recompute_bp, double *e)
{
    replica *r = &o->rep[irep];
    if (recompute_bp) r->flg_bp_energy = 0;
    for (int k = 0; k < O_ENE_N; k++) e[k] = 0.0;
    e[O_ENE_BOND] = energy_bond(o, r);
    e[O_ENE_ANGL] = energy_angl(o, r);
    if (o->flg_dih_cos || o->flg_dih_exp) e[O_ENE_DIHE] = energy_dih(o, r);
    e[O_ENE_BP] = energy_bp_all(o, r, irep, tempK);
    e[O_ENE_EXV] = energy_wca(o, r);
    if (o->flg_ele) e[O_ENE_ELE] = energy_ele_DH(o, r);
    if (o->flg_stage && energy_stage(o, r, &e[O_ENE_STAGE])) return 1;
    if (o->ntwz_DCF + o->ntwz_FR > 0) e[O_ENE_TWZ] = energy_tweezers(o, r);
    if (o->flg_bias_rg) e[O_ENE_RG] = energy_rg(o, r);
    if (o->nrest > 0) e[O_ENE_REST] = energy_restraint(o, r);
    double s = 0.0;
    for (int k = 1; k < O_ENE_N; k++) s += e[k];
    e[O_ENE_TOTAL] = s;
    return 0;
}

/* src/energy_kinetic.F90:17-29 */
int oracle_energy_kinetic(sis_oracle *o, int irep, double *ekin)
{
    replica *r = &o->rep[irep];
    double e = 0.0;
#pragma omp parallel for reduction(+ : e)
    for (int imp = 0; imp < o->nmp; imp++) {
        if (o->is_frozen[imp]) r->velos[3 * imp] = r->velos[3 * imp + 1] = r->velos[3 * imp + 2] = 0.0;
        e += o->mass[imp] * dot3(r->velos + 3 * imp, r->velos + 3 * imp);
    }
    *ekin = 0.5 * e;
    return 0;
}

int oracle_get_bp_state(sis_oracle *o, int irep, int *status, double *ene)
{
    replica *r = &o->rep[irep];
    if (status) memcpy(status, r->bp_status, (size_t)r->nbp * sizeof(int));
    if (ene) memcpy(ene, r->ene_bp, (size_t)r->nbp * sizeof(double));
    return 0;
}
int oracle_get_rg(sis_oracle *o, int irep, double *rg) { *rg = o->rep[irep].rg; return 0; }

/* ---- counter-based RNG (replaces mt_stream + rnd_boxmuller; shared spec with the GPU
 * build: Philox4x32-10, key = (seed_lo, seed_hi + stream), counter = (step_lo, step_hi,
 * bead, draw); uniforms are 53-bit in (0,1); normals by Box-Muller) ------------------ */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static double u53(uint32_t lo, uint32_t hi)
{
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
void oracle_gauss(uint64_t seed, uint32_t stream, uint64_t step, uint32_t bead, double g[3])
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) + stream};
    uint32_t ctr[4] = {(uint32_t)step, (uint32_t)(step >> 32), bead, 0}, o4[4];
    philox4x32_10(ctr, key, o4);
    double u1 = u53(o4[0], o4[1]), u2 = u53(o4[2], o4[3]);
    double rad = sqrt(-2.0 * log(u1));
    g[0] = rad * cos(2.0 * O_PI * u2);
    g[1] = rad * sin(2.0 * O_PI * u2);
    ctr[3] = 1;
    philox4x32_10(ctr, key, o4);
    u1 = u53(o4[0], o4[1]); u2 = u53(o4[2], o4[3]);
    g[2] = sqrt(-2.0 * log(u1)) * cos(2.0 * O_PI * u2);
}
double oracle_uniform(uint64_t seed, uint32_t stream, uint64_t counter)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) + stream};
    uint32_t ctr[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), 0, 2}, o4[4];
    philox4x32_10(ctr, key, o4);
    return u53(o4[0], o4[1]);
}

/* ---- dynamics ---------------------------------------------------------------- */
/* src/job_md.F90:94-108,700-767: coef = {md_coef(1), md_coef(2), md_coef(3), md_coef_rep(1)} */
static void md_coef(const sis_oracle *o, const replica *r, int imp, double c[4])
{
    double radius = 10.0, fric;
    if (o->integrator == INT_GJF) {
        double v = o->viscosity_Pas * sqrt(1.0e3 / KCAL2JOUL) * N_AVO * 1.0e-20;
        fric = 6.0 * O_PI * v * radius;
        double c1 = 0.5 * o->dt * fric / o->mass[imp];
        double c2 = sqrt(1.0 / (1.0 + c1));
        c[0] = (1.0 - c1) / (1.0 + c1);
        c[1] = c2 * o->dt / o->mass[imp];
        c[2] = o->is_frozen[imp] ? 0.0 : c2 * o->dt;
        c[3] = 0.5 * c2 / o->mass[imp] * sqrt(2.0 * fric * BOLTZ_KCAL_MOL * r->tempK * o->dt);
    } else {
        double v = (JOUL2KCAL_MOL * 1.0e-18) * o->viscosity_Pas;
        fric = 6.0 * O_PI * v * radius;
        c[0] = o->is_frozen[imp] ? 0.0 : o->dt / fric;
        c[1] = c[2] = 0.0;
        c[3] = o->is_frozen[imp] ? 0.0 : sqrt(2.0 * BOLTZ_KCAL_MOL * r->tempK * o->dt / fric);
    }
}
int oracle_md_coef(sis_oracle *o, int irep, int imp, double coef[4])
{
    md_coef(o, &o->rep[irep], imp, coef);
    return 0;
}

/* src/job_md.F90:157-176 with the counter-based generator (step = 2^63 + {0,1}) */
int oracle_init_velocities(sis_oracle *o, int irep)
{
    replica *r = &o->rep[irep];
    for (int imp = 0; imp < o->nmp; imp++) {
        double g[3], c[4];
        md_coef(o, r, imp, c);
        double c1 = sqrt(r->tempK * BOLTZ_KCAL_MOL / o->mass[imp]);
        oracle_gauss(o->seed, (uint32_t)irep, 0x8000000000000000ull, (uint32_t)imp, g);
        for (int k = 0; k < 3; k++) r->velos[3 * imp + k] = c1 * g[k];
        oracle_gauss(o->seed, (uint32_t)irep, 0x8000000000000001ull, (uint32_t)imp, g);
        for (int k = 0; k < 3; k++) r->accels[3 * imp + k] = c[3] * g[k];
    }
    return 0;
}

/* one pass of the irep loop body, src/job_md.F90:374-460; o->istep already incremented */
int oracle_md_step(sis_oracle *o, int irep, const double *gauss, int *rebuilt)
{
    replica *r = &o->rep[irep];
    int nmp = o->nmp;
    double d2max = 0.0, d2max_2nd = 0.0;
    for (int imp = 0; imp < nmp; imp++) {
        double d2 = dot3(r->xyz_move + 3 * imp, r->xyz_move + 3 * imp);
        if (d2 > d2max) { d2max_2nd = d2max; d2max = d2; }
        else if (d2 > d2max_2nd) d2max_2nd = d2;
    }
    int flg_MC_save = o->flg_bp_MC;
    if (rebuilt) *rebuilt = 0;
    if (sqrt(d2max) + sqrt(d2max_2nd) > o->nl_margin) {
        struct timespec ts0, ts1;
        clock_gettime(CLOCK_MONOTONIC, &ts0);
        oracle_neighbor_list(o, irep);
        clock_gettime(CLOCK_MONOTONIC, &ts1);
        o->t_nl += (double)(ts1.tv_sec - ts0.tv_sec) + 1e-9 * (double)(ts1.tv_nsec - ts0.tv_nsec);
        memset(r->xyz_move, 0, 3 * (size_t)nmp * sizeof(double));
        if (o->nstep_bp_MC > 0) o->flg_bp_MC = 1;
        if (rebuilt) *rebuilt = 1;
    }
    double *F = (double *)xcalloc(3 * (size_t)nmp, sizeof(double));
    oracle_force(o, irep, F);
    if (o->flg_bp_MC) memcpy(r->bp_status_MC, r->bp_status, (size_t)r->nbp * sizeof(int));
    (void)flg_MC_save; /* the reference leaves flg_bp_MC set for the remaining replicas too */
    /* src/job_md.F90:416-456 `omp parallel do`; the counter-based Gaussians are drawn inside the
     * loop (the reference draws its 3N Mersenne-Twister normals serially first, :409-413) */
#pragma omp parallel for
    for (int imp = 0; imp < nmp; imp++) {
        double g[3], c[4];
        if (gauss) memcpy(g, gauss + 3 * imp, sizeof g);
        else oracle_gauss(o->seed, (uint32_t)irep, (uint64_t)o->istep, (uint32_t)imp, g);
        md_coef(o, r, imp, c);
        for (int k = 0; k < 3; k++) {
            double dxyz;
            if (o->integrator == INT_GJF) {
                double accels_pre = c[3] * g[k];
                r->velos[3 * imp + k] = c[0] * r->velos[3 * imp + k] + c[1] * F[3 * imp + k] +
                                        (r->accels[3 * imp + k] + accels_pre);
                r->accels[3 * imp + k] = accels_pre;
                dxyz = c[2] * r->velos[3 * imp + k];
            } else
                dxyz = c[0] * F[3 * imp + k] + c[3] * g[k];
            r->xyz[3 * imp + k] += dxyz;
            r->xyz_move[3 * imp + k] += dxyz;
        }
    }
    free(F);
    return 0;
}

/* threads the OpenMP regions use (1 in the serial build) and the wall time spent in the serial
 * neighbor_list since the last reset: what bench.py prints beside the CPU baseline */
int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
double oracle_nl_seconds(sis_oracle *o, int reset)
{
    double t = o->t_nl;
    if (reset) o->t_nl = 0.0;
    return t;
}

/* src/job_md.F90:360-460 for all local replicas (no exchange, no output) */
int oracle_md_run(sis_oracle *o, int nsteps, int64_t *nrebuild)
{
    int64_t nreb = 0;
    for (int s = 0; s < nsteps; s++) {
        o->istep++;
        o->flg_bp_MC = 0;
        if (o->nstep_bp_MC > 0 && o->istep % o->nstep_bp_MC == 0) o->flg_bp_MC = 1;
        for (int irep = 0; irep < o->nrep_proc; irep++) {
            int reb;
            if (oracle_md_step(o, irep, NULL, &reb)) return 1;
            nreb += reb;
        }
    }
    if (nrebuild) *nrebuild = nreb;
    return 0;
}

/* ---- REMD -------------------------------------------------------------------- */
/* src/init_replica.F90:36-98 + src/var_replica.F90:112-259 (single process: nrep_proc = nrep_all) */
int oracle_remd_setup(sis_oracle *o, const int nrep[4], const double *values, int nstep_exchange)
{
    o->flg_replica = 1;
    o->ndim_replica = 0;
    o->nrep_all = 1;
    for (int v = 0; v < REPT_MAX; v++) {
        o->nrep[v] = nrep[v];
        o->flg_repvar[v] = nrep[v] > 1;
        for (int k = 0; k < nrep[v] && k < MAX_REP_PER_DIM; k++)
            o->replica_values[v][k] = values[v * MAX_REP_PER_DIM + k];
        if (o->flg_repvar[v]) { o->ndim_replica++; o->nrep_all *= nrep[v]; }
    }
    if (o->ndim_replica > MAX_REP_DIM) return fail("too many replica dimensions");
    if (o->nrep_all != o->nrep_proc) return fail("oracle REMD is single-process: nrep_proc must equal nrep_all");
    int n = o->nrep_all;
    o->nstep_exchange = nstep_exchange;
    o->rep2lab = (int *)xcalloc(n, sizeof(int));
    o->lab2rep = (int *)xcalloc(n, sizeof(int));
    o->lab2val = (double *)xcalloc((size_t)REPT_MAX * n, sizeof(double));
    /* lab2val: src/init_replica.F90:70-92 */
    int n_division = 1;
    for (int v = 0; v < REPT_MAX; v++) {
        if (!o->flg_repvar[v]) continue;
        int n_division_pre = n_division;
        n_division *= o->nrep[v];
        for (int irep = 1; irep <= o->nrep[v]; irep++) {
            int n_continue = n / n_division, n_skip = n / n_division_pre;
            for (int ido = 1; ido <= n_division_pre; ido++)
                for (int is = 1; is <= n_continue; is++) {
                    int label = (irep - 1) * n_continue + (ido - 1) * n_skip + is;
                    o->lab2val[v * n + label - 1] = o->replica_values[v][irep - 1];
                }
        }
    }
    for (int g = 0; g < n; g++) { o->rep2lab[g] = g + 1; o->lab2rep[g] = g + 1; }
    /* type_array: src/var_replica.F90:139-160 */
    int idx = 0;
    for (int v = 0; v < REPT_MAX; v++)
        if (o->flg_repvar[v]) o->type_array[idx++] = v;
    /* exchange_pair_tb: src/var_replica.F90:162-259 */
    int ncol = 2 * o->ndim_replica;
    o->exchange_pair_tb = (int *)xcalloc((size_t)ncol * n, sizeof(int));
    int icounter = 0;
    for (int iparity = 1; iparity <= 2; iparity++) {
        n_division = 1;
        for (int idimn = 0; idimn < o->ndim_replica; idimn++) {
            int *tb = o->exchange_pair_tb + (size_t)icounter * n;
            icounter++;
            int nv = o->nrep[o->type_array[idimn]];
            int n_division_pre = n_division;
            n_division *= nv;
            int n_continue = n / n_division;
            int iset = 0;
            for (int icycle = 1; icycle <= n_division_pre; icycle++) {
                int flg_post_zero = 0, start;
                if (iparity == 1) {
                    if (nv % 2 == 1) flg_post_zero = 1;
                    start = 1;
                } else {
                    for (int ic = 1; ic <= n_continue; ic++) tb[iset++] = 0;
                    if (nv % 2 == 0) flg_post_zero = 1;
                    start = 2;
                }
                for (int ir = start; ir <= nv - 1; ir += 2)
                    for (int ic = 1; ic <= n_continue * 2; ic++) {
                        iset++;
                        tb[iset - 1] = (ic <= n_continue) ? iset + n_continue : iset - n_continue;
                    }
                if (flg_post_zero)
                    for (int ic = 1; ic <= n_continue; ic++) tb[iset++] = 0;
            }
        }
    }
    o->idx_pair = 1;
    /* replica parameters from the labels */
    for (int g = 0; g < n; g++) {
        int lab = o->rep2lab[g] - 1;
        replica *r = &o->rep[g];
        double T = o->flg_repvar[REPT_TEMP] ? o->lab2val[REPT_TEMP * n + lab] : r->tempK;
        double I = o->flg_repvar[REPT_ION] ? o->lab2val[REPT_ION * n + lab] : r->ionic_strength;
        double f = o->flg_repvar[REPT_TWZDCF] ? o->lab2val[REPT_TWZDCF * n + lab] : r->twz_force;
        double g0 = o->flg_repvar[REPT_RG] ? o->lab2val[REPT_RG * n + lab] : r->bias_rg_0;
        oracle_set_replica_params(o, g, T, I, f, g0);
    }
    return 0;
}

int oracle_remd_pair_table(sis_oracle *o, int *tb)
{
    memcpy(tb, o->exchange_pair_tb, (size_t)2 * o->ndim_replica * o->nrep_all * sizeof(int));
    return 0;
}
int oracle_remd_state(sis_oracle *o, int *rep2lab, int *lab2rep, double *lab2val)
{
    int n = o->nrep_all;
    if (rep2lab) memcpy(rep2lab, o->rep2lab, n * sizeof(int));
    if (lab2rep) memcpy(lab2rep, o->lab2rep, n * sizeof(int));
    if (lab2val) memcpy(lab2val, o->lab2val, (size_t)REPT_MAX * n * sizeof(double));
    return 0;
}
static int get_pair(const sis_oracle *o, int label) /* 1-based */
{
    return o->exchange_pair_tb[(size_t)(o->idx_pair - 1) * o->nrep_all + label - 1];
}
static void apply_label_params(sis_oracle *o, int irep, int label)
{
    int n = o->nrep_all, lab = label - 1;
    replica *r = &o->rep[irep];
    double T = o->flg_repvar[REPT_TEMP] ? o->lab2val[REPT_TEMP * n + lab] : r->tempK;
    double I = o->flg_repvar[REPT_ION] ? o->lab2val[REPT_ION * n + lab] : r->ionic_strength;
    if (o->flg_repvar[REPT_TWZDCF]) {
        r->twz_force = o->lab2val[REPT_TWZDCF * n + lab];
        for (int k = 0; k < 3 * o->ntwz_DCF; k++) r->twz_DCF_forces[k] = r->twz_force * o->twz_DCF_direction[k];
    }
    if (o->flg_repvar[REPT_RG]) r->bias_rg_0 = o->lab2val[REPT_RG * n + lab];
    if (o->flg_ele) set_ele(o, r, T, I, NULL, NULL); /* ele_cutoff is NOT refreshed here (quirk 3) */
    r->tempK = T;
    r->ionic_strength = I;
}

/* src/energy_replica.F90:45-163 on an exchange step.  energies [nrep][11],
 * replica_energies [nrep_all][2] */
/* flg_save_bp of src/energy_replica.F90:16: set when the exchange step is also a save step */
int oracle_set_save_bp(sis_oracle *o, int flg) { o->flg_save_bp = flg; return 0; }
/* bp_status_MC(:,:) = bp_status(:,:) after the initial energies, src/job_md.F90:344-348 */
int oracle_bp_status_to_mc(sis_oracle *o)
{
    for (int irep = 0; irep < o->nrep_proc; irep++) {
        replica *r = &o->rep[irep];
        memcpy(r->bp_status_MC, r->bp_status, (size_t)r->nbp * sizeof(int));
    }
    return 0;
}

int oracle_energy_replica(sis_oracle *o, double *energies, double *replica_energies)
{
    int n = o->nrep_all;
    for (int irep = 0; irep < n; irep++)
        if (oracle_energy_sumup(o, irep, o->rep[irep].tempK, 0, energies + O_ENE_N * irep)) return 1;
    /* :64-70 save bp information to output in write_bp */
    int **st_tmp = NULL;
    double **en_tmp = NULL;
    if (o->flg_save_bp) {
        st_tmp = (int **)xcalloc(n, sizeof(int *));
        en_tmp = (double **)xcalloc(n, sizeof(double *));
        for (int irep = 0; irep < n; irep++) {
            st_tmp[irep] = dup_int(o->rep[irep].bp_status, o->rep[irep].nbp);
            en_tmp[irep] = dup_dbl(o->rep[irep].ene_bp, o->rep[irep].nbp);
        }
    }
    for (int irep = 0; irep < n; irep++) {
        replica *r = &o->rep[irep];
        r->flg_bp_energy = 0; /* :72 */
        replica_energies[2 * irep] = energies[O_ENE_N * irep + O_ENE_TOTAL];
        int label_own = o->rep2lab[irep], label_opp = get_pair(o, label_own);
        if (label_opp == 0 || label_opp > n) { replica_energies[2 * irep + 1] = 0.0; continue; }
        double e[O_ENE_N];
        double T_own = r->tempK, I_own = r->ionic_strength;
        apply_label_params(o, irep, label_opp);
        double T_opp = r->tempK;
        if (oracle_energy_sumup(o, irep, T_opp, 0, e)) return 1;
        replica_energies[2 * irep + 1] = e[O_ENE_TOTAL];
        r->tempK = T_own; r->ionic_strength = I_own;
        apply_label_params(o, irep, label_own);
    }
    if (o->flg_save_bp) { /* :157-162 */
        for (int irep = 0; irep < n; irep++) {
            memcpy(o->rep[irep].bp_status, st_tmp[irep], (size_t)o->rep[irep].nbp * sizeof(int));
            memcpy(o->rep[irep].ene_bp, en_tmp[irep], (size_t)o->rep[irep].nbp * sizeof(double));
            free(st_tmp[irep]); free(en_tmp[irep]);
        }
        free(st_tmp); free(en_tmp);
    }
    return 0;
}

int oracle_replica_exchange(sis_oracle *o, const double *E, const double *uniforms, int *n_attempt,
                            int *n_accept)
{
    int n = o->nrep_all, nrand = 0, natt = 0, nacc = 0;
    for (int l_i = 1; l_i <= n; l_i++) {
        int l_j = get_pair(o, l_i);
        if (l_j < l_i) continue;
        double u = uniforms[nrand++];
        int rep_i = o->lab2rep[l_i - 1], rep_j = o->lab2rep[l_j - 1];
        double temp_m, temp_n;
        if (o->flg_repvar[REPT_TEMP]) {
            temp_m = o->lab2val[REPT_TEMP * n + l_i - 1];
            temp_n = o->lab2val[REPT_TEMP * n + l_j - 1];
        } else temp_m = temp_n = o->rep[0].tempK;
        double pot_i_m = E[2 * (rep_i - 1)], pot_i_n = E[2 * (rep_i - 1) + 1];
        double pot_j_n = E[2 * (rep_j - 1)], pot_j_m = E[2 * (rep_j - 1) + 1];
        double delta = ((pot_j_m - pot_i_m) / temp_m - (pot_j_n - pot_i_n) / temp_n) / BOLTZ_KCAL_MOL;
        natt++;
        if (delta < 0.0 || u <= exp(-delta)) {
            nacc++;
            o->rep2lab[rep_i - 1] = l_j; o->rep2lab[rep_j - 1] = l_i;
            o->lab2rep[l_i - 1] = rep_j; o->lab2rep[l_j - 1] = rep_i;
            if (o->flg_repvar[REPT_TEMP]) { /* scale_velo :175-213 */
                double sf = sqrt(temp_n / temp_m);
                for (int k = 0; k < 3 * o->nmp; k++) {
                    o->rep[rep_i - 1].velos[k] *= sf;
                    o->rep[rep_j - 1].velos[k] /= sf;
                }
            }
        }
    }
    /* set_forward: src/var_replica.F90:112-126 */
    o->idx_pair = (o->idx_pair == 2 * o->ndim_replica) ? 1 : o->idx_pair + 1;
    /* parameter refresh: src/job_md.F90:538-567 */
    for (int irep = 0; irep < n; irep++) apply_label_params(o, irep, o->rep2lab[irep]);
    if (n_attempt) *n_attempt = natt;
    if (n_accept) *n_accept = nacc;
    return 0;
}

/* ---- job_check_force: src/job_check_force.F90:53-80 --------------------------- */
int oracle_check_force(sis_oracle *o, double tempK, double *f_energy, double *f_force,
                       double *maxdiff)
{
    const double small = 0.001f; /* `small = 0.001` is a default-real literal in the reference */
    replica *r = &o->rep[0];
    double e[O_ENE_N], md = 0.0;
    if (oracle_force(o, 0, f_force)) return 1;
    for (int imp = 0; imp < o->nmp; imp++)
        for (int k = 0; k < 3; k++) {
            double save = r->xyz[3 * imp + k];
            r->xyz[3 * imp + k] = save + small;
            if (oracle_energy_sumup(o, 0, tempK, 1, e)) return 1;
            double e_save = e[O_ENE_TOTAL];
            r->xyz[3 * imp + k] = save - small;
            if (oracle_energy_sumup(o, 0, tempK, 1, e)) return 1;
            f_energy[3 * imp + k] = -(e_save - e[O_ENE_TOTAL]) / (2.0 * small);
            r->xyz[3 * imp + k] = save;
            double d = fabs(f_energy[3 * imp + k] - f_force[3 * imp + k]);
            if (d > md) md = d;
        }
    if (maxdiff) *maxdiff = md;
    return 0;
}
