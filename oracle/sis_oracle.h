/*
 * sis_oracle.h -- CPU restatement of the Hori-Lab/SIS per-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sis_b200/ (the product) may include,
 * link or dlopen this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * timed CPU baseline.
 *
 * Parity status: "parity unpinned" at the 1e-10 level -- the reference is
 * Fortran and no Fortran compiler exists in this image, so oracle/_ref cannot
 * be built.  The oracle is pinned against the reference's own artefacts:
 *   - check/test_DCD_CAG47/CAG47.dcd + CAG47.energy.out (bond/angle/WCA to
 *     <=2e-4, DH/bp to <0.5 %), see tests/test_oracle_golden.py
 *   - the CHECK_FORCE criterion |f_energy - f_force| <= 1e-3 of
 *     src/job_check_force.F90:21,78 on the check/ structures
 *   - structural invariants (468 bp3 entries, sum F = 0, PBC-shift invariance)
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * it follows.  All arithmetic is fp64, loops are in the reference's order; the
 * serial build (libsis_oracle.so) is the checker, the -fopenmp build
 * (libsis_oracle_omp.so) threads the loops the reference threads and is the CPU baseline.  Indices are 0-based inside; list getters return the
 * reference's 1-based values.
 */
#ifndef SIS_ORACLE_H
#define SIS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sis_oracle sis_oracle;

/* energies(0:10): src/const_idx.F90:7-21 */
enum { O_ENE_TOTAL = 0, O_ENE_BOND, O_ENE_ANGL, O_ENE_DIHE, O_ENE_BP, O_ENE_EXV,
       O_ENE_ELE, O_ENE_STAGE, O_ENE_TWZ, O_ENE_RG, O_ENE_REST, O_ENE_N };

sis_oracle *oracle_create(void);
void oracle_destroy(sis_oracle *o);

/* topology: src/var_top.F90:9-28; seq codes A0 U1 G2 C3 D4; moltypes RNA1 CIRC2 */
int oracle_set_topology(sis_oracle *o, int nmp, int nchains, const int *nmp_chain,
                        const int *moltypes, const int *seq, const int *has_charge,
                        const int *is_frozen, const double *mass);
int oracle_set_replicas(sis_oracle *o, int nrep_proc);
/* angl_type 0 harmonic / 1 ReB; dih flags: bit0 cos, bit1 exp */
int oracle_set_ff_bonded(sis_oracle *o, double bond_k, double bond_r0, int angl_type,
                         double angl_k, double angl_t0, int dih_flags, double dih_k,
                         double dih_w, double dih_p0);
int oracle_set_ff_wca(sis_oracle *o, double sigma, double eps);
/* bp_paras: [3][15] for GC, AU, GU as in the .ff file:
 * U0 bond_k bond_r angl_k1..4 angl_theta1..4 dihd_k1..2 dihd_phi1..2;
 * CG/UA/UG derived with the k1<->k2, k3<->k4 swap (src/read_force_field.F90:203-207) */
int oracle_set_ff_bp(sis_oracle *o, int model, int min_loop, int max_bp_per_nt,
                     int nstep_MC, double cutoff_energy, const double *bp_paras);
int oracle_set_ff_nn(sis_oracle *o, const double *NN_dH, const double *NN_dS, double dH0,
                     double dS0, int flg_NNend, const double *NNend_dH,
                     const double *NNend_dS, double dHend0, double dSend0);
int oracle_set_ele(sis_oracle *o, int flg_ele, int cutoff_type, double cutoff_inp,
                   int excl_bond, int excl_angle, double length_per_charge,
                   int temp_independent, double tempK_ref);
int oracle_set_pbc(sis_oracle *o, int flg_pbc, const double *box);
int oracle_set_md(sis_oracle *o, int integrator, double dt, double viscosity_Pas,
                  double nl_margin, uint64_t seed);
int oracle_set_bias_ss(sis_oracle *o, int flg, double force);
int oracle_set_stage(sis_oracle *o, int flg, double sigma, double eps);
int oracle_set_tweezers_dcf(sis_oracle *o, int n, const int *pairs /*1-based [n][2]*/,
                            const double *direction /*[n][3]*/);
int oracle_set_tweezers_fr(sis_oracle *o, int n, const int *pairs, const double *k /*[n][2]*/,
                           const double *init /*[n][2][3]*/, const double *velo /*[n][2][3]*/);
int oracle_set_bias_rg(sis_oracle *o, int flg, int pott /*1 harmonic 2 flat-bottomed*/, double k);
int oracle_set_restraints(sis_oracle *o, int n, const int *id /*[n][3] 1-based i,j,type*/,
                          const double *rcut, const double *para /*[n][3] eps,d,s*/);
/* per replica: temperature, ionic strength, DCF force magnitude, Rg0 */
int oracle_set_replica_params(sis_oracle *o, int irep, double tempK, double ionic_strength,
                              double twz_force, double rg0);
int oracle_finalize_setup(sis_oracle *o);

/* state: Fortran layout xyz(3,nmp) */
int oracle_set_state(sis_oracle *o, int irep, const double *xyz, const double *velos,
                     const double *accels);
int oracle_get_state(sis_oracle *o, int irep, double *xyz, double *velos, double *accels);
int oracle_set_step(sis_oracle *o, int64_t istep);

/* lists */
int oracle_neighbor_list(sis_oracle *o, int irep);          /* src/neighbor_list.F90 */
int oracle_full_lists(sis_oracle *o);                       /* src/list_exv/ele/bp.F90 */
int oracle_get_list_sizes(sis_oracle *o, int irep, int *nwca, int *nele, int *nbp);
int oracle_get_lists(sis_oracle *o, int irep, int *wca_mp, int *ele_mp, int *bp_mp,
                     double *bp_coef);
int oracle_get_local_lists(sis_oracle *o, int *nbond, int *nangl, int *ndih, int *bond_mp,
                           int *angl_mp, int *dih_mp);
int oracle_bp3_table(sis_oracle *o, int *n, int *hash4096, double *dH, double *dS);
/* on-the-fly bp_map / bp3_map value of src/init_bp.F90:184-313 for a 1-based pair */
int oracle_bp_lookup(sis_oracle *o, int imp, int jmp, int *bptype, int *bp3idx);
/* src/read_ss.F90 + src/init_bp.F90:178-209: restrict the bp map to the pairs of a ct/bpseq/bpl file */
int oracle_set_bp_restrict(sis_oracle *o, int n, const int *pairs /*1-based [n][2]*/);

/* hot path */
int oracle_set_bp_mc(sis_oracle *o, int flg_bp_MC);         /* job_md.F90:365-368 */
int oracle_force(sis_oracle *o, int irep, double *forces);  /* src/force.F90 */
/* term mask bit k = energies index k (1..10) */
int oracle_force_terms(sis_oracle *o, int irep, unsigned mask, double *forces);
int oracle_energy_sumup(sis_oracle *o, int irep, double tempK, int recompute_bp,
                        double *energies);                  /* src/energy_sumup.F90 */
int oracle_energy_kinetic(sis_oracle *o, int irep, double *ekin);
int oracle_get_bp_state(sis_oracle *o, int irep, int *status, double *ene);
int oracle_get_rg(sis_oracle *o, int irep, double *rg);

/* dynamics: one step of job_md.F90:374-460 for replica irep.  gauss = 3*nmp
 * normal deviates in (xyz, bead) order, or NULL to use the counter-based
 * generator shared with the GPU build (oracle_gauss). */
int oracle_md_step(sis_oracle *o, int irep, const double *gauss, int *rebuilt);
int oracle_md_run(sis_oracle *o, int nsteps, int64_t *nrebuild);
/* src/energy_replica.F90:16,64-70,157-162 (flg_save_bp) and src/job_md.F90:344-348 (bp_status_MC = bp_status) */
int oracle_set_save_bp(sis_oracle *o, int flg);
int oracle_bp_status_to_mc(sis_oracle *o);
/* bench.py: OpenMP threads in use (1 = serial build); wall seconds inside neighbor_list */
int oracle_num_threads(void);
double oracle_nl_seconds(sis_oracle *o, int reset);
int oracle_init_velocities(sis_oracle *o, int irep);        /* job_md.F90:157-176 */
void oracle_gauss(uint64_t seed, uint32_t stream, uint64_t step, uint32_t bead, double g[3]);
double oracle_uniform(uint64_t seed, uint32_t stream, uint64_t counter);
int oracle_md_coef(sis_oracle *o, int irep, int imp, double coef[4]);

/* set_ele: src/set_ele.F90:47-102 */
int oracle_set_ele_calc(sis_oracle *o, double tempK, double ionstr, double out[6]);
int oracle_get_ele(sis_oracle *o, int irep, double out[4]);

/* REMD: src/init_replica.F90, src/var_replica.F90, src/energy_replica.F90,
 * src/replica_exchange.F90 */
int oracle_remd_setup(sis_oracle *o, const int nrep[4], const double *values /*[4][256]*/,
                      int nstep_exchange);
int oracle_remd_pair_table(sis_oracle *o, int *tb /*[nrep_all][2*ndim]*/);
int oracle_remd_state(sis_oracle *o, int *rep2lab, int *lab2rep, double *lab2val);
int oracle_energy_replica(sis_oracle *o, double *energies, double *replica_energies);
int oracle_replica_exchange(sis_oracle *o, const double *replica_energies,
                            const double *uniforms, int *n_attempt, int *n_accept);

/* job_check_force: src/job_check_force.F90:46-82 (perturbation applied by the caller) */
int oracle_check_force(sis_oracle *o, double tempK, double *f_energy, double *f_force,
                       double *maxdiff);

const char *oracle_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
