/* mmd_b200.h -- C ABI of the B200-native constrained-HMC hot path.
 *
 * Drop-in boundary for the path that the reference drives through
 * sde/mici_extensions.py (ConditionedDiffusionConstrainedSystem :208-1259, the projection solvers
 * :1323-1476, SwitchPartitionTransition :1262-1282) and Mici's ConstrainedLeapfrogIntegrator
 * (call site scripts/utils.py:284-290).  The reference has no FFI of its own (it is pure Python on
 * JAX); each entry point below names the reference function it replaces.  All chains of a handle
 * share one model configuration and are processed as one batch.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in signatures; `double*` arguments are HOST pointers unless
 *     the name ends in `_dev`;
 *   - host arrays use the reference's per-chain layout: q, p, vct: [n_chains][dim_q] row-major with
 *     q = [u | v_0 | v_seq(T*S, dim_v) | n(T, dim_y)] (mici_extensions.py:476-484);
 *     x_obs_seq: [n_chains][T][dim_x]; constraint vectors: [n_chains][n_c(partition)];
 *   - return value 0 = ok, <0 = API / CUDA error (text via mmd_last_error_string); numerical
 *     failures are per-chain status bits, never errors across the ABI:
 *       1 not converged, 2 diverged / NaN   -> mici.errors.ConvergenceError (:1393-1402)
 *       4 non-reversible step               -> mici.errors.NonReversibleStepError
 *       8 non-finite Hamiltonian            -> mici.errors.HamiltonianDivergenceError
 *   - calls are asynchronous on the handle's CUDA stream; getters synchronise.  One handle per host
 *     thread / GPU (thread-compatible, not thread-safe).
 */
#ifndef MMD_B200_H
#define MMD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmd_handle_s* mmd_handle;

enum { MMD_MODEL_FHN = 0, MMD_MODEL_SIR = 1,
       MMD_MODEL_FHN_NOTEBOOK = 2 /* FHN with the prior parametrisation of FitzHugh-Nagumo_example.ipynb */ };
enum { MMD_NOISE_NONE = 0, MMD_NOISE_FIXED = 1, MMD_NOISE_PARAM = 2 };
enum { MMD_SOLVER_QUASI_NEWTON = 0, MMD_SOLVER_NEWTON = 1 };

typedef struct {
  int model;               /* MMD_MODEL_*  (sde/example_models/{fhn,sir}.py) */
  int num_obs;             /* T   y_seq.shape[0] */
  int num_steps_per_obs;   /* S */
  int num_obs_per_subseq;  /* R; <= 0 or == T: no blocking (mici_extensions.py:321-324) */
  int dim_u;               /* dim_z, +1 when the observation noise scale is inferred */
  int noise;               /* MMD_NOISE_*  (generate_sigma None / number / callable, :353-358) */
  double sigma_fixed;      /* used when noise == MMD_NOISE_FIXED */
  int gaussian_splitting;  /* use_gaussian_splitting (:303) */
  double obs_interval;
  const double* y_seq;     /* host [T * dim_y] */
  int n_chains;
  int device;              /* CUDA device ordinal */
} mmd_config;

/* ConditionedDiffusionConstrainedSystem.__init__ (:211-377) for a batch of chains. */
int mmd_create(const mmd_config* cfg, mmd_handle* out);
int mmd_destroy(mmd_handle h);
const char* mmd_last_error_string(void);

int mmd_dim_q(mmd_handle h);
int mmd_num_partition(mmd_handle h);                 /* system.num_partition (:362) */
int mmd_num_constraints(mmd_handle h, int partition); /* len(constr(...)) */
int mmd_num_blocks(mmd_handle h, int partition);
int mmd_n_chains(mmd_handle h);

/* ---- resident chain state (ConditionedDiffusionHamiltonianState :1285-1320) ------------------ */
/* Upload positions / momenta / conditioned states; p may be NULL (left unchanged).  Invalidates
 * the cached linearisation like a Mici ChainState variable assignment does. */
int mmd_set_state(mmd_handle h, const double* q, const double* p, const double* x_obs_seq, int partition);
/* The same without waiting for the copies: the host buffers must be page-locked and must stay untouched until the
 * next mmd_synchronize (or any blocking call) on this handle.  Lets a host thread keep several handles busy, e.g.
 * upload the next batch of chains on one handle while another handle computes (bench.py's end-to-end pipeline). */
int mmd_set_state_async(mmd_handle h, const double* q, const double* p, const double* x_obs_seq, int partition);
int mmd_get_state(mmd_handle h, double* q, double* p, double* x_obs_seq);
int mmd_set_momentum(mmd_handle h, const double* p);
/* Same with DEVICE pointers in the reference layout ([n_chains][dim_q], [n_chains][T][dim_x]): the
 * zero-copy path for DLPack producers (JAX / CuPy / torch arrays already on the GPU).  The library
 * re-tiles on device into its internal tile layout; asynchronous. */
int mmd_set_state_dev(mmd_handle h, const double* q_dev, const double* p_dev, const double* x_obs_seq_dev,
                      int partition);
int mmd_get_state_dev(mmd_handle h, double* q_dev, double* p_dev, double* x_obs_seq_dev);
/* Stream ordering for the device-pointer path.  All work of a handle runs on its own non-blocking CUDA stream.
 *   mmd_get_stream   the handle's cudaStream_t: a DLPack consumer passes it to the producer's
 *                    __dlpack__(stream=...) so that the producer orders its pending writes before our reads;
 *   mmd_wait_stream  the handle's stream waits for everything queued so far on `other_stream` (producer side,
 *                    for callers that hand over raw pointers instead of going through the DLPack protocol);
 *   mmd_stream_wait  `other_stream` waits for everything queued so far on the handle's stream (consumer side of
 *                    mmd_get_state_dev).
 * Without one of these (or mmd_synchronize) the device-pointer calls are NOT ordered against other streams. */
void* mmd_get_stream(mmd_handle h);
int mmd_get_device(mmd_handle h);
int mmd_wait_stream(mmd_handle h, void* other_stream);
int mmd_stream_wait(mmd_handle h, void* other_stream);
/* [u | v_0] of every chain's current position, [n_chains][dim_u + dim_v_0] (host array): the inputs of the reference
 * scripts' trace functions (generate_z(u), generate_x_0(z, v_0); fhn_model_noiseless_obs_chmc_experiment.py:102-117)
 * without reading back the whole position.  Blocking. */
int mmd_get_head(mmd_handle h, double* out);
/* mmd_get_state without the final synchronisation (separate staging buffers per array; the host arrays are valid
 * after mmd_synchronize or any blocking call; page-locked arrays overlap the copies with other handles' work) */
int mmd_get_state_async(mmd_handle h, double* q, double* p, double* x_obs_seq);
int mmd_get_partition(mmd_handle h);
/* chains per CTA tile chosen at create time (tuning knob MMD_CPB; informational) */
int mmd_chains_per_tile(mmd_handle h);
/* global index of this handle's first chain: offsets the Philox counters so that ranks that shard
 * one population of chains draw disjoint, rank-count-independent streams */
int mmd_set_chain_offset(mmd_handle h, int chain0);
/* Chain regrouping (throughput option, off by default).  The iteration loops of the projection solves run per CTA
 * tile for as long as the tile's slowest chain; a chain's iteration count is persistent (it follows the stiffness
 * of its parameters).  With regrouping on, every partition switch (mmd_switch_partition, mmd_transition_end,
 * mmd_hmc_transition) re-assigns the chains to slots sorted by the iteration count of their last step -- the
 * re-tiling pass of the switch moves every position anyway.  Every chain's results are bit-identical to a run
 * without regrouping (Philox streams are keyed by chain id, nothing depends on the tile neighbours); per-chain
 * OUTPUTS are then in slot order and mmd_get_slot_chains returns the chain id of every slot.  Setting states
 * (mmd_set_state*, mmd_init_linear_interpolation) returns to chain order.  Not combinable with per-chain step
 * sizes, adaptation or the mmd_vec_* work vectors. */
int mmd_set_chain_regrouping(mmd_handle h, int on);
int mmd_get_slot_chains(mmd_handle h, int* chain_of_slot /* [n_chains] */);

/* ---- system ops on the resident state ----------------------------------------------------------- */
/* jacob_constr_blocks + chol_gram_blocks + log_det_sqrt_gram (+ grad_log_det_sqrt_gram when
 * with_grad): one evaluation fills every cached quantity (:1151-1184). */
int mmd_linearize(mmd_handle h, int with_grad);
/* constr (:473-519, :1151-1155) -> c_out [n_chains][n_c] */
int mmd_constr(mmd_handle h, double* c_out);
/* log_det_sqrt_gram (:1169-1171) -> [n_chains] ; grad_log_det_sqrt_gram (:1173-1184) -> [n_chains][dim_q] */
int mmd_log_det_sqrt_gram(mmd_handle h, double* out);
int mmd_grad_log_det_sqrt_gram(mmd_handle h, double* out);
/* h = h1 + h2 (:1186-1202) -> [n_chains] */
int mmd_hamiltonian(mmd_handle h, double* out);
/* project_onto_cotangent_space (:1252-1254) applied to the resident momentum */
int mmd_project_momentum(mmd_handle h);
/* normal_space_component (:1243-1250) of host vectors vct [n_chains][dim_q] at the resident position */
int mmd_normal_space_component(mmd_handle h, const double* vct, double* out);
/* update_x_obs_seq / generate_x_obs_seq (:384-397, :1240-1241) from the resident position */
int mmd_update_x_obs_seq(mmd_handle h);
/* SwitchPartitionTransition.sample (:1279-1282) */
int mmd_switch_partition(mmd_handle h);
/* sample_momentum (:1256-1259): Philox-4x32-10 standard normals, then cotangent projection */
int mmd_sample_momentum(mmd_handle h, uint64_t seed, uint64_t offset);

/* Compressed factors behind jacob_constr_blocks / chol_gram_blocks, for tests and for rebuilding
 * the reference's dense blocks on the host.  name in {"K","Psib","xend","A","L","DinvA"}: out is
 * [n_chains][n_blocks][rows] with block-local rows (K: [obs in block][S][dim_x][dim_v], Psib:
 * [obs][dim_x][dim_x], A / DinvA: [row in block][dim_u], L: packed lower Cholesky factor of D_b with
 * the diagonal stored INVERTED); "LC": [n_chains][rows] packed factor of the capacitance matrix.
 * Returns rows via *rows_out when out==NULL. */
int mmd_get_factor(mmd_handle h, const char* name, double* out, int* rows_out);

/* ---- integrator ----------------------------------------------------------------------------- */
typedef struct {
  int solver;              /* MMD_SOLVER_* */
  double constraint_tol;   /* projection_solver_kwargs (scripts/utils.py:278-282) */
  double position_tol;
  double divergence_tol;
  int max_iters;
  double reverse_check_tol; /* ConstrainedLeapfrogIntegrator(reverse_check_tol=...) */
} mmd_integrator_opts;
void mmd_default_integrator_opts(mmd_integrator_opts* o);

/* Run-time parameters of the model's generators (the reference passes generate_z / generate_x_0 as Python callables,
 * mici_extensions.py:211-240; here a parametrised family per model, so priors change without recompiling).
 * FHN (22 values): z_i = a_i u_i + b_i, exponentiated where m_i != 0, x_0 = v_0 + c + E z, laid out
 * [a (4) | b (4) | m (4) | c (2) | E (2 x 4 row-major)]; default = sde/example_models/fhn.py:41-51; the model id
 * MMD_MODEL_FHN_NOTEBOOK is the same model created with the notebook's values.  SIR has none (0 values).
 * Set them right after mmd_create, before the first state is loaded. */
int mmd_num_generator_params(mmd_handle h);
int mmd_get_generator_params(mmd_handle h, double* params);
int mmd_set_generator_params(mmd_handle h, const double* params, int n);

/* One ConstrainedLeapfrogIntegrator.step (n_inner_step = 1) of size `dt` (signed = dir * step_size)
 * for every chain.  Chains whose step fails keep their state; their status bits say why. */
int mmd_leapfrog_step(mmd_handle h, double dt, const mmd_integrator_opts* opts);
/* The same step with the B part split into `n_inner_step` inner steps of size dt / n_inner_step, each with its own
 * projection and reverse check (Mici ConstrainedLeapfrogIntegrator(n_inner_step=...); the reference's
 * --num-inner-h2-step, scripts/utils.py:132, 286).  A chain that fails in any inner step is left exactly where it
 * was before the call.  n_inner_step = 1 is mmd_leapfrog_step. */
int mmd_leapfrog_step_inner(mmd_handle h, double dt, int n_inner_step, const mmd_integrator_opts* opts);
/* One full Markov transition for every chain, entirely on device: IndependentMomentumTransition
 * (Philox draw keyed by (seed, iter) + cotangent projection), `n_leapfrog` constrained leapfrog
 * steps of size dt, Metropolis accept on the Hamiltonian error (integrator errors reject), then
 * optionally SwitchPartitionTransition (scripts/utils.py:292-301; static instead of dynamic
 * trajectory length). */
int mmd_hmc_transition(mmd_handle h, double dt, int n_leapfrog, uint64_t seed, uint64_t iter,
                       const mmd_integrator_opts* opts, int switch_partition);
/* The same transition in three calls so a driver can interleave its own work between steps. */
int mmd_transition_begin(mmd_handle h, uint64_t seed, uint64_t iter);
int mmd_transition_steps(mmd_handle h, double dt, int n_steps, const mmd_integrator_opts* opts);
int mmd_transition_end(mmd_handle h, uint64_t seed, uint64_t iter, int switch_partition);
/* Per-chain step sizes [n_chains] (host): used by the leapfrog / transition entry points instead of their
 * scalar `dt` argument, whose sign still gives the integration direction; NULL switches back to the scalar. */
int mmd_set_step_sizes(mmd_handle h, const double* dt);
int mmd_get_step_sizes(mmd_handle h, double* dt);
/* mici.adapters.DualAveragingStepSizeAdapter (scripts/utils.py:303-306: target 0.8, regularisation coefficient
 * 0.1; Mici defaults decay 0.75, offset 10) run per chain on device: after every transition the chain's step
 * size is updated from its accept_stat.  mmd_adapt_stop = finalize: step size <- exp(smoothed log step size)
 * per chain, or (pool != 0) the mean over chains as Mici does for several chains. */
int mmd_adapt_start(mmd_handle h, double init_step_size, double target, double reg_coefficient, double iter_decay,
                    double iter_offset);
/* the same with one initial step size per chain (Mici initialises each chain's adapter from its own coarse search,
 * DualAveragingStepSizeAdapter._find_and_set_init_step_size; batched in adaptation.find_init_step_sizes) */
int mmd_adapt_start_per_chain(mmd_handle h, const double* init_step_sizes, double target, double reg_coefficient,
                              double iter_decay, double iter_offset);
int mmd_adapt_stop(mmd_handle h, int pool);
/* same update from accept statistics supplied by the caller (dynamic transitions built by the host) */
int mmd_adapt_update(mmd_handle h, const double* accept_stat);

/* ---- vector primitives for host-driven tree building -------------------------------------------
 * (mici.transitions.MultinomialDynamicIntegrationTransition, scripts/utils.py:292-301, batched: the host
 * keeps the per-chain tree bookkeeping, the device holds every state vector and does every O(dim_q)
 * operation.)  Vectors are addressed by id: MMD_VEC_Q / MMD_VEC_P = the chains' live position / momentum,
 * 0 .. n-1 = auxiliary arrays reserved with mmd_aux_reserve.  mask: host int[n_chains] or NULL (all). */
enum { MMD_VEC_Q = -1, MMD_VEC_P = -2 };
int mmd_aux_reserve(mmd_handle h, int n_arrays);
/* dst = beta * dst + alpha * src for the masked chains */
int mmd_vec_axpby(mmd_handle h, int dst, int src, double alpha, double beta, const int* mask);
/* per chain, with s = c - d + a:  out1 = a . s,  out2 = e . s  (no-U-turn criterion on momentum sums;
 * e may be MMD_VEC_P) */
int mmd_vec_uturn(mmd_handle h, int a, int d, int c, int e, double* out1, double* out2);
/* park (mask != 0) / release chains: parked chains are skipped by every kernel like failed ones, without an
 * error bit; clear_errors also clears the integrator error bits of all chains */
int mmd_set_inactive(mmd_handle h, const int* mask, int clear_errors);
/* re-evaluate everything cached at the live position (after mmd_vec_axpby wrote it) without touching the
 * per-chain status words */
int mmd_relinearize(mmd_handle h);
/* accepted flag, accept_stat = min(1, exp(h0 - h1)), integrator status of the last transition */
int mmd_get_transition_stats(mmd_handle h, int* accepted, double* accept_prob, int* status);
/* per-chain status / diagnostics of the last step (any pointer may be NULL) */
int mmd_get_step_info(mmd_handle h, int* status, int* iters_fwd, int* iters_rev, double* rev_dist);
/* jitted_solve_projection_onto_manifold_quasi_newton (:1323-1402) or, with opts->solver ==
 * MMD_SOLVER_NEWTON, jitted_solve_projection_onto_manifold_newton (:1405-1476; per iteration the
 * constraint is re-linearised at the iterate and the non-symmetric block products J(q) J(q_prev)^T are
 * LU-factorised, :689-763, :944-981) on host inputs: projects q [n][dim_q] using the linearisation at
 * the resident position; returns the projected q, per-chain status bits and iteration counts. */
int mmd_project_quasi_newton(mmd_handle h, const double* q_in, double dt, const mmd_integrator_opts* opts,
                             double* q_out, int* status, int* iters);

/* find_initial_state_by_linear_interpolation (:1479-1547) for every chain: u [n][dim_u],
 * v_0 [n][dim_v_0], x_obs_seq [n][T][dim_x] (host); fills the resident position and x_obs_seq. */
int mmd_init_linear_interpolation(mmd_handle h, const double* u, const double* v_0, const double* x_obs_seq,
                                  int partition);

/* Standard-HMC target of the noisy-observation model: conditioned_diffusion_neg_log_dens_and_grad
 * (sde/mici_extensions.py:82-205; the EuclideanMetricSystem baseline of scripts/*_hmc_experiment.py).
 * q [n][mmd_hmc_dim] = [u | v_0 | v_seq] (no observation-noise variables), host.  val [n] receives
 *   1/2 sum_k ((y_k - h(x_k)) / sigma)^2 + T dim_y log sigma  (+ 1/2 |q|^2 when add_prior != 0, i.e. without
 * Gaussian splitting, :183-187); grad [n][mmd_hmc_dim] its gradient (the reference: jax.value_and_grad through
 * lax.scan, :189) and resid [n][T] the scaled residuals (y - h(x)) / sigma; grad and resid may be NULL.  Needs a
 * handle created with observation noise (fixed or inferred). */
int mmd_hmc_dim(mmd_handle h);
int mmd_hmc_target(mmd_handle h, const double* q, int add_prior, double* val, double* grad, double* resid);
/* Batched Adam descent on that objective: find_initial_state_by_gradient_descent_noisy_system (:1679-1801; Adam as
 * jax.experimental.optimizers.adam, :1733).  begin: (re)start the chains with mask != 0 (NULL: all) from
 * u_v [n][mmd_hmc_dim] with cleared moments and iteration counter; eval: objective gradient and residuals at the
 * current iterates, msr [n] = mean squared residual (:1759), val [n] (may be NULL) the objective; update: one Adam
 * step (:1737-1741) for the chains with upd != 0; get: the iterates and the residuals of the last eval. */
int mmd_adam_begin(mmd_handle h, const double* u_v, const int* mask);
int mmd_adam_eval(mmd_handle h, double* msr, double* val);
int mmd_adam_update(mmd_handle h, double step_size, const int* upd);
int mmd_adam_get(mmd_handle h, double* u_v, double* resid);

/* number of kernel launches issued since creation (bench.py's gpu_launches) and event timing on the
 * handle's stream */
long long mmd_launch_count(mmd_handle h);
/* total successful chain leapfrog steps since creation / last reset (device-side counter) */
long long mmd_successful_steps(mmd_handle h, int reset);
/* total projection-solver iterations executed over all chains (forward + reverse solves) */
long long mmd_total_qn_iterations(mmd_handle h, int reset);
/* Development aid: per-phase cycle counters of the fused leapfrog kernel (thread 0 of every CTA), available only
 * in builds with -DMMD_PHASE_CLOCK (tools/phase_times.py); returns an error in product builds. */
int mmd_debug_phase_cycles(mmd_handle h, unsigned long long* out64, int reset);
/* per-kernel CUDA-event timing on the handle's stream: kernel ids 0 linearise (k_point), 1 momentum
 * projection (k_project), 2 quasi-Newton projection (k_qn), 3 fused leapfrog step(s) (k_leapfrog) */
int mmd_profile_enable(mmd_handle h, int on, int max_launches);
int mmd_profile_summary(mmd_handle h, int kernel_id, int* count, double* total_ms);
int mmd_timer_start(mmd_handle h);
int mmd_timer_stop_ms(mmd_handle h, float* ms);
int mmd_synchronize(mmd_handle h);

#ifdef __cplusplus
}
#endif
#endif /* MMD_B200_H */
