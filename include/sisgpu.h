/*
 * sisgpu.h -- C ABI of the B200-native SIS hot path (libsisgpu.so).
 *
 * The reference (Hori-Lab/SIS, Fortran 2008) has no plugin/FFI boundary: the
 * seam is a set of external subroutines that talk through module globals
 * (SURVEY.md section 8b).  This header is the boundary the Fortran driver binds
 * through ISO_C_BINDING (sis_b200/fortran/sisgpu_mod.F90).
 * Each entry point names the reference routine / call site it replaces
 * (file:line relative to the reference tree).
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success, non-zero on
 *    error (message via sisgpu_last_error()); the Fortran shim turns non-zero
 *    into `call sis_abort()` (src/sis_abort.F90:12-19).
 *  - arrays are the Fortran module arrays passed by reference: column-major,
 *    so xyz(3,nmp) arrives as [nmp][3] doubles.  Bead / pair indices crossing
 *    the boundary are 1-based exactly as in the reference.
 *  - `irep` is the 1-based in-process replica index of the reference.
 *  - the library owns all device memory; the caller owns every host array.
 *  - callers are single-threaded per handle (as the reference's drivers are).
 *  - all arithmetic is fp64 (PREC = REAL64, src/const.F90:10).
 */
#ifndef SISGPU_H
#define SISGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sisgpu sisgpu_t;

#define SISGPU_ENE_N 11          /* energies(0:10), src/const_idx.F90:7-21 */
#define SISGPU_NCCL_ID_BYTES 128
#define SISGPU_IPC_HANDLE_BYTES 64

const char *sisgpu_last_error(void);
int sisgpu_version(void);

/* ---- life cycle ------------------------------------------------------------ */
/* replaces nothing in the reference; rank -> GPU mapping is new (myrank mod ngpu) */
int sisgpu_create(sisgpu_t **out, int device);
int sisgpu_destroy(sisgpu_t *h);

/* ---- set-up: the outputs of the reference's init_* routines ------------------ */
/* var_top (src/var_top.F90:9-28) after init_sequence/read_fasta; seq codes A0 U1 G2 C3 D4
 * (src/const_idx.F90:31-40) per bead in bead order; moltypes RNA 1 / circRNA 2;
 * has_charge from init_ele (src/init_ele.F90:45-62); is_frozen from init_freeze;
 * mass from src/job_md.F90:116-133 (NULL = use those hard-coded masses) */
int sisgpu_set_topology(sisgpu_t *h, int nmp, int nchains, const int *nmp_chain,
                        const int *moltypes, const int *seq, const int *has_charge,
                        const int *is_frozen, const double *mass);
/* nrep_proc of src/init_replica.F90:56-65 */
int sisgpu_set_replicas(sisgpu_t *h, int nrep_proc);
/* var_potential bonded scalars read by src/read_force_field.F90:68-131.
 * angl_type 0 = [potential.angle], 1 = [potential.angle_ReB];
 * dih_flags bit0 = [potential.dihedral] (cos), bit1 = [potential.dihedral_exp] */
int sisgpu_set_ff_bonded(sisgpu_t *h, double bond_k, double bond_r0, int angl_type,
                         double angl_k, double angl_t0, int dih_flags, double dih_k,
                         double dih_w, double dih_p0);
int sisgpu_set_ff_wca(sisgpu_t *h, double sigma, double eps);
/* [potential.basepair] + [Basepair]: bp_paras[3][15] for GC, AU, GU in .ff order
 * (U0 bond_k bond_r angl_k1..4 angl_theta1..4 dihd_k1..2 dihd_phi1..2); the CG/UA/UG
 * sets are derived as in src/read_force_field.F90:203-207 */
int sisgpu_set_ff_bp(sisgpu_t *h, int model, int min_loop, int max_bp_per_nt, int nstep_MC,
                     double cutoff_energy, const double *bp_paras);
/* [NN] tables (21 stacks in NNT order, src/const_idx.F90:96-121) from which the 468-entry
 * bp3 table of src/init_bp.F90:39-172 is rebuilt; replaces the N x N bp_map / bp3_map */
/* read_ss (src/read_ss.F90) + the restricted branch of init_bp (src/init_bp.F90:178-209): only the
 * pairs of a ct / bpseq / bpl file may form base pairs.  pairs = 1-based (imp, jmp) [n][2] in any
 * order; the filters of read_ss (chain ends, minimum loop, isolated pairs, unknown types) are applied
 * here.  n < 0 removes the restriction.  Call after set_topology and set_ff_bp */
int sisgpu_set_bp_restrict(sisgpu_t *h, int n, const int *pairs);
int sisgpu_set_ff_nn(sisgpu_t *h, const double *NN_dH, const double *NN_dS, double dH0,
                     double dS0, int flg_NNend, const double *NNend_dH,
                     const double *NNend_dS, double dHend0, double dSend0);
/* [Electrostatic] (src/read_input.F90:690-770) */
int sisgpu_set_ele(sisgpu_t *h, int flg_ele, int cutoff_type, double cutoff_inp,
                   int excl_bond, int excl_angle, double length_per_charge,
                   int temp_independent, double tempK_ref);
/* src/pbc.F90:5-9; also the target of pbc_resize (src/job_md.F90:576-594) */
int sisgpu_set_pbc(sisgpu_t *h, int flg_pbc, const double *box);
/* [MD]: integrator 1 = GJF-2GJ, 2 = Ermak-McCammon (src/const_idx.F90:51-57) */
int sisgpu_set_md(sisgpu_t *h, int integrator, double dt, double viscosity_Pas,
                  double nl_margin, uint64_t seed);
int sisgpu_set_bias_ss(sisgpu_t *h, int flg, double force);
int sisgpu_set_stage(sisgpu_t *h, int flg, double sigma, double eps);
int sisgpu_set_tweezers_dcf(sisgpu_t *h, int n, const int *pairs, const double *direction);
int sisgpu_set_tweezers_fr(sisgpu_t *h, int n, const int *pairs, const double *k,
                           const double *init, const double *velo);
int sisgpu_set_bias_rg(sisgpu_t *h, int flg, int pott, double k);
int sisgpu_set_restraints(sisgpu_t *h, int n, const int *id, const double *rcut,
                          const double *para);
/* per-replica parameters (0-based irep here mirrors the oracle; the Fortran shim
 * subtracts 1): temperature, ionic strength, DCF force magnitude, Rg0.  Performs
 * set_ele (src/set_ele.F90:47-102) and set_md_coef (src/job_md.F90:700-767). */
int sisgpu_set_replica_params(sisgpu_t *h, int irep, double tempK, double ionic_strength,
                              double twz_force, double rg0);
/* list_local (src/list_local.F90), init_bp tables (src/init_bp.F90), device allocation */
int sisgpu_finalize_setup(sisgpu_t *h);

/* ---- state ------------------------------------------------------------------ */
/* xyz / velos / accels (3,nmp,irep) of src/var_state.F90:26-28; NULL = skip */
int sisgpu_set_state(sisgpu_t *h, int irep, const double *xyz, const double *velos,
                     const double *accels);
int sisgpu_get_state(sisgpu_t *h, int irep, double *xyz, double *velos, double *accels);
int sisgpu_set_step(sisgpu_t *h, int64_t istep);
int sisgpu_get_step(sisgpu_t *h, int64_t *istep);
/* Restart payload of the counter-based generators.  The reference writes 643 int32 of
 * Mersenne-Twister state per stream under the PRNG / PRNGREP tags (src/write_rst.F90:69-131) and reads
 * them back in src/read_rst.F90:405-600; the device generators are counter-based, so the payload is
 * state[0] = seed, [1] = istep (the Gaussian counter), [2] = draw counter of the exchange stream
 * (PRNGREP), [3] = nrep_proc, [4 ..] = draw counters of the Monte-Carlo base-pair pruning per local
 * replica (PRNG).  get with state == NULL returns the word count in *n.  Together with XYZ / VELO /
 * ACCEL (sisgpu_get_state), STEP and REPLICA this makes `stop -> restart` continue the trajectory
 * (src/job_md.F90:642-681). */
int sisgpu_get_rng_state(sisgpu_t *h, int64_t *state, int cap, int *n);
int sisgpu_set_rng_state(sisgpu_t *h, const int64_t *state, int n);

/* ---- lists ------------------------------------------------------------------ */
/* call neighbor_list(irep): src/job_md.F90:197,389,584 -> src/neighbor_list.F90 */
int sisgpu_neighbor_list(sisgpu_t *h, int irep);
/* call list_bp / list_exv / list_ele: src/main.F90:150-152 */
int sisgpu_full_lists(sisgpu_t *h);
int sisgpu_get_list_sizes(sisgpu_t *h, int irep, int *nwca, int *nele, int *nbp);
/* wca_mp(2,nwca) ele_mp(2,nele) bp_mp(7,nbp) bp_coef(2,nbp) of src/var_potential.F90:85-112,
 * 1-based, in the reference's (i, j) loop order */
int sisgpu_get_lists(sisgpu_t *h, int irep, int *wca_mp, int *ele_mp, int *bp_mp,
                     double *bp_coef);
int sisgpu_get_local_lists(sisgpu_t *h, int *nbond, int *nangl, int *ndih, int *bond_mp,
                           int *angl_mp, int *dih_mp);
int sisgpu_bp3_table(sisgpu_t *h, int *n, int *hash4096, double *dH, double *dS);
/* write_bpcoef (src/write_bpcoef.F90:30-83) without init_bp's N x N bp3_map: every sequence-allowed pair imp < jmp
 * (1-based, the reference's loop order) with its dH and dS * 1e-3; the driver applies coef_dG * (dH - T dS) < 0 per
 * replica and formats the line.  cap = 0 sizes the arrays (*n).  Bead ids are int32 here: the int16 of .bp
 * (src/var_io.F90:25) is a property of the file format the driver writes */
int sisgpu_bpcoef(sisgpu_t *h, int cap, int *n, int *imp, int *jmp, double *dH, double *dS);

/* ---- hot path ----------------------------------------------------------------*/
/* flg_bp_MC of src/job_md.F90:365-368 for the next force / energy call */
int sisgpu_set_bp_mc(sisgpu_t *h, int flg_bp_MC);
/* call force(irep, forces): src/job_md.F90:397, src/job_check_force.F90:53 */
int sisgpu_force(sisgpu_t *h, int irep, double *forces);
/* per-term force (the reference's DUMPFORCE probe); mask bit k = energies index k */
int sisgpu_force_terms(sisgpu_t *h, int irep, unsigned mask, double *forces);
/* call energy_sumup(irep, tempK, energies): src/job_md.F90:290, src/energy_replica.F90:54,121,
 * src/job_check_force.F90:62,66.  recompute_bp != 0 <=> flg_bp_energy = .False. */
int sisgpu_energy_sumup(sisgpu_t *h, int irep, double tempK, int recompute_bp,
                        double *energies);
/* call energy_kinetic: src/job_md.F90:291,501 */
int sisgpu_energy_kinetic(sisgpu_t *h, int irep, double *ekin);
/* job_dcd batch mode (src/job_dcd.F90:65-79): energy_sumup with flg_bp_energy = .False. for every
 * frame of a trajectory, nrep_proc frames per launch through the replica slots (same parameters
 * in every slot; lists from sisgpu_full_lists as src/main.F90:148-154).  xyz32 = float32
 * [nframes][nmp][3] as stored in the DCD (src/dcd.F90:584-592); energies = [nframes][11] */
int sisgpu_dcd_energies(sisgpu_t *h, int nframes, const float *xyz32, double tempK, double *energies);
/* bp_status_MC(:,:) = bp_status(:,:), the line after the initial energies of job_md (src/job_md.F90:344-348):
 * call it after the initial sisgpu_energy_sumup (flg_bp_MC set, recompute_bp) of a run or of a restart when
 * nstep_MC > 0; the steps up to the first Monte-Carlo step evaluate exactly these pairs */
int sisgpu_bp_status_to_mc(sisgpu_t *h);
/* bp_status / ene_bp for write_bp (src/write_bp.F90:36-49) */
int sisgpu_get_bp_state(sisgpu_t *h, int irep, int *status, double *ene);
int sisgpu_get_rg(sisgpu_t *h, int irep, double *rg);
/* initial velocities / beta(0): src/job_md.F90:157-176 (counter-based generator) */
int sisgpu_init_velocities(sisgpu_t *h, int irep);
/* the loop body src/job_md.F90:360-460 for all local replicas, nsteps times, resident on
 * the device; with REMD configured it also runs energy_replica + the allgather +
 * replica_exchange + parameter refresh (src/job_md.F90:474-481,514-570) every
 * nstep_exchange steps.  nrebuild (may be NULL) returns the neighbour-list rebuild count */
int sisgpu_md_steps(sisgpu_t *h, int64_t nsteps, int64_t *nrebuild);
/* job_check_force (src/job_check_force.F90:53-80): forces + central finite difference of
 * energy_sumup for every coordinate, evaluated on the device */
int sisgpu_check_force(sisgpu_t *h, double tempK, double *f_energy, double *f_force,
                       double *maxdiff);
/* launch statistics: kernels launched by this handle since creation */
int sisgpu_launch_count(sisgpu_t *h, int64_t *n);
/* options: name = "launch_mode" (0 stream launches, 1 CUDA-graph chunks, 2 persistent kernels =
 * default), "fast_kernel" (0: never use k_md_fast), "local_kernel" (0: no replica-local mode of
 * k_md_fast), "mc_persistent" (0: Monte-Carlo steps through the per-phase kernels), "pipeline" (0:
 * host-driven list rebuilds), "remd_device" (0: host-driven exchange steps), "fast_max_cta",
 * "list_capacity" (row capacity K), "chunk", "threads_per_bead", "timing", "profile" */
int sisgpu_set_option(sisgpu_t *h, const char *name, double value);
/* CUDA-event time (ms) accumulated per kernel family since the last reset:
 * out[0] step kernels, out[1] neighbour list, out[2] energy, out[3] exchange;
 * out[4..7] SM clocks block 0 of the persistent kernel spent in force phase, barrier, integrate
 * phase, barrier, out[8] steps counted, out[9..11] base-pair warps (gate, evaluate, arrive),
 * out[12..13] bonded-table warps (tables, Gaussians) (option "profile"); out must hold 16 doubles */
int sisgpu_get_timers(sisgpu_t *h, double *out, int reset);
/* Output staging for a save step (src/job_md.F90:487-509).  _begin reduces on the
 * device what the writers need -- flags: 1 = float32 DCD frame (src/dcd.F90:584-592), 2 = .bp
 * records (bp_status and ene_bp <= -kT, src/write_bp.F90:36-49), 4 = .bpall records instead
 * (:76-88), 8 = the 11 energies of energy_sumup with the cached bp energies + Ekin
 * (src/write_out.F90:83-110) -- into one of two staging slots and starts the device-to-host copies
 * on a second stream: sisgpu_md_steps called afterwards overlaps with them.  _end waits for that
 * slot only; records come back 1-based in the reference's bp_mp order as (imp, jmp, real(ene, 4)).
 * Any output pointer may be NULL. */
int sisgpu_snapshot_begin(sisgpu_t *h, int irep, double tempK, int flags, int *slot);
int sisgpu_snapshot_end(sisgpu_t *h, int slot, float *xyz32, int *nbp, int *bp_i, int *bp_j,
                        float *bp_e, int cap, double *energies, double *ekin);
/* device-side wall clock of a region of calls: sisgpu_mark(h, 0) / sisgpu_mark(h, 1) record a CUDA
 * event on the library's stream, sisgpu_elapsed_ms returns the time between the two marks (it
 * waits for mark 1).  Used by bench.py: torch.cuda.Event only sees torch's own stream */
/* name of the step kernel the last sisgpu_md_steps call launched (k_md_fast<1,false|true>, k_md_fused,
 * or the per-phase kernels of launch modes 0 / 1) */
const char *sisgpu_last_step_kernel(sisgpu_t *h);
int sisgpu_mark(sisgpu_t *h, int which);
int sisgpu_elapsed_ms(sisgpu_t *h, double *ms);
/* diagnostic: accumulated SM clocks of CTA c of k_md_fast in out[8c..8c+7] = pair units, wait for
 * tables / base pairs, integrate, position barrier, base-pair warps total, wait for the CTA's own
 * warps, active base pairs, spare (option "profile") */
int sisgpu_get_cta_profile(sisgpu_t *h, int ncta, double *out);

/* ---- REMD ------------------------------------------------------------------- */
/* init_replica (src/init_replica.F90) + make_replica_exchange_pair_tb
 * (src/var_replica.F90:162-259).  nrep[4] / values[4][256] in REPT order
 * TEMP, TWZDCF, ION, RG (src/const_idx.F90:87-94).  rank / nranks give the block
 * distribution grep = rank*nrep_proc + irep (src/init_replica.F90:117-119). */
int sisgpu_remd_setup(sisgpu_t *h, const int nrep[4], const double *values,
                      int nstep_exchange, int rank, int nranks);
/* NCCL plumbing for the one hot-loop collective, MPI_ALLREDUCE at src/job_md.F90:531-533:
 * rank 0 calls sisgpu_nccl_unique_id and broadcasts the bytes (MPI_BCAST / torch.distributed) */
int sisgpu_nccl_unique_id(unsigned char *id /*[SISGPU_NCCL_ID_BYTES]*/);
int sisgpu_comm_init(sisgpu_t *h, int nranks, int rank, const unsigned char *id);
/* Peer mapping for the exchange step inside the persistent kernel (ranks of one node): every rank exports one
 * CUDA IPC handle of its exchange buffer, the driver all-gathers the SISGPU_IPC_HANDLE_BYTES-byte handles
 * (MPI_ALLGATHER) and every rank imports all of them.  The MPI_ALLREDUCE of src/job_md.F90:531-533 then becomes
 * peer stores + arrival counters over NVLink inside sisgpu_md_steps; without the import a multi-rank run uses
 * the ncclAllGather launch chain.  Call after sisgpu_remd_setup. */
int sisgpu_peer_export(sisgpu_t *h, unsigned char *handle /*[SISGPU_IPC_HANDLE_BYTES]*/);
int sisgpu_peer_import(sisgpu_t *h, int nranks, const unsigned char *handles /*[nranks][SISGPU_IPC_HANDLE_BYTES]*/);
int sisgpu_remd_pair_table(sisgpu_t *h, int *tb);
int sisgpu_remd_state(sisgpu_t *h, int *rep2lab, int *lab2rep, double *lab2val);
int sisgpu_remd_stats(sisgpu_t *h, int64_t *n_attempt, int64_t *n_accept);
/* REPLICA block of a restart (src/write_rst.F90:41-60, src/read_rst.F90:405-460): the label permutation
 * rep2lab(1:nrep_all) and the position in the exchange schedule (var_replica's idx_pair) */
int sisgpu_remd_set_labels(sisgpu_t *h, const int *rep2lab, int idx_pair);
int sisgpu_remd_get_idx_pair(sisgpu_t *h, int *idx_pair);
/* call energy_replica(...) on an exchange step: src/energy_replica.F90:45-163;
 * energies [nrep_proc][11], replica_energies [nrep_all][2] after the allgather */
int sisgpu_energy_replica(sisgpu_t *h, double *energies, double *replica_energies);
/* call replica_exchange(velos, replica_energies, tempk): src/replica_exchange.F90:44-154,
 * with caller-supplied uniforms (NULL = the library's counter-based stream) */
int sisgpu_replica_exchange(sisgpu_t *h, const double *replica_energies,
                            const double *uniforms, int *n_attempt, int *n_accept);

#ifdef __cplusplus
}
#endif
#endif
