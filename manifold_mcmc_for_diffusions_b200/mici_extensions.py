"""Drop-in for the constrained-HMC part of ``sde.mici_extensions`` (reference ``sde/mici_extensions.py``):
the same names, signatures and error behaviour, with every numerical operation executed by the CUDA
library through the C ABI (``include/mmd_b200.h``).  Mici samplers / integrators call it unchanged.

    ConditionedDiffusionConstrainedSystem   :208-1259
    SwitchPartitionTransition               :1262-1282
    ConditionedDiffusionHamiltonianState    :1285-1320
    jitted_solve_projection_onto_manifold_quasi_newton   :1323-1402
    jitted_solve_projection_onto_manifold_newton         :1405-1476
    find_initial_state_by_linear_interpolation           :1479-1547
    find_initial_state_by_gradient_descent_noisy_system  :1679-1801
    conditioned_diffusion_neg_log_dens_and_grad          :82-205   (the standard-HMC baseline's target)
    OnlineBlockDiagonalMetricAdapter                     :1804-1931
    split, split_and_reshape                :31-53

This module is the *compatibility* path: one Mici chain = one resident chain on the device, one C
call per system method (the way Mici drives the reference's jitted JAX functions).  The throughput
path is :class:`manifold_mcmc_for_diffusions_b200.BatchedChains`, which keeps thousands of chains
resident and runs whole leapfrog steps / transitions per launch.  There is no CPU fallback: model
callables must be the tagged functions of :mod:`.example_models` (arbitrary Python callables cannot
be compiled into device functors and raise ``NotImplementedError``).
"""

from numbers import Number

import numpy as np

try:  # prefer the real Mici when it is installed
    from mici.adapters import Adapter
    from mici.errors import AdaptationError, ConvergenceError, HamiltonianDivergenceError
    from mici.matrices import (DensePositiveDefiniteMatrix, IdentityMatrix,
                               PositiveDefiniteBlockDiagonalMatrix)
    from mici.states import ChainState, _cache_key_func
    from mici.systems import System, cache_in_state, cache_in_state_with_aux
    from mici.transitions import Transition
except ImportError:  # pragma: no cover - exercised in this environment
    from .mici_compat.adapters import Adapter
    from .mici_compat.errors import AdaptationError, ConvergenceError, HamiltonianDivergenceError
    from .mici_compat.matrices import (DensePositiveDefiniteMatrix, IdentityMatrix,
                                       PositiveDefiniteBlockDiagonalMatrix)
    from .mici_compat.states import ChainState, _cache_key_func, cache_in_state, cache_in_state_with_aux
    from .mici_compat.systems import System
    from .mici_compat.transitions import Transition

from .batched import BatchedChains

SOLVER_QUASI_NEWTON, SOLVER_NEWTON = 0, 1


def split(v, lengths):
    """Split an array along first dimension into slices of specified lengths (:31-40)."""
    i = 0
    parts = []
    for length in lengths:
        parts.append(v[i: i + length])
        i += length
    if i < len(v):
        parts.append(v[i:])
    return parts


def split_and_reshape(array, shapes):
    """Split an array along first dimension into subarrays of specified shapes (:43-53)."""
    array = np.asarray(array)
    lengths = [int(np.prod(shape)) for shape in shapes]
    return tuple(part.reshape(shape + array.shape[1:]) for part, shape in
                 zip(split(array, lengths), shapes))


class DeviceArray(np.ndarray):
    """Result type of the batched private handles (the role `jax.numpy.DeviceArray` plays for the reference's jitted
    closures): a NumPy array whose computation has already been synchronised, or a zero-size ticket for block
    factors that stay on the device."""

    def block_until_ready(self):
        return self


def standard_normal_neg_log_dens(q):
    """Unnormalised negative log density of standard normal vector (:56-58); a leading batch axis is mapped."""
    q = np.asarray(q)
    if q.ndim > 1:
        return (0.5 * np.sum(np.square(q), axis=-1)).view(DeviceArray)
    return 0.5 * float(np.sum(np.square(q)))


def standard_normal_grad_neg_log_dens(q):
    """Gradient and value of negative log density of standard normal vector (:61-63)."""
    q = np.asarray(q)
    if q.ndim > 1:
        return q.view(DeviceArray), (0.5 * np.sum(np.square(q), axis=-1)).view(DeviceArray)
    return q, 0.5 * float(np.sum(np.square(q)))


def _model_tag(*funcs):
    tags = {getattr(f, "_mmd_model", None) for f in funcs}
    if None in tags or len(tags) != 1:
        raise NotImplementedError(
            "ConditionedDiffusionConstrainedSystem runs on generated CUDA device functors: forward_func, "
            "generate_x_0, generate_z, obs_func (and a callable generate_σ) must be the functions of "
            "manifold_mcmc_for_diffusions_b200.example_models.<model>; arbitrary callables are not supported "
            "and there is no CPU fallback")
    return tags.pop()


class ConditionedDiffusionConstrainedSystem(System):
    """Specialised mici system class for conditioned diffusion problems (reference :208-1259),
    identity metric, evaluated on the GPU."""

    def __init__(self, obs_interval, num_steps_per_obs, num_obs_per_subseq, y_seq, dim_u, dim_x, dim_v,
                 forward_func, generate_x_0, generate_z, obs_func, generate_σ=None,
                 use_gaussian_splitting=False, metric=None, dim_v_0=None, device=0):
        if use_gaussian_splitting and metric is not None:
            raise ValueError("Only identity matrix metric can be used with Gaussian splitting.")  # :293-297
        if metric is not None and not isinstance(metric, IdentityMatrix):
            raise NotImplementedError("Only the identity metric is implemented on the device path.")  # :305-315
        super().__init__(neg_log_dens=standard_normal_neg_log_dens,
                         grad_neg_log_dens=standard_normal_grad_neg_log_dens)
        self.use_gaussian_splitting = use_gaussian_splitting
        self.metric = IdentityMatrix()
        funcs = [forward_func, generate_x_0, generate_z, obs_func]
        if generate_σ is None:
            noise, sigma = 0, 0.0
        elif isinstance(generate_σ, Number):
            noise, sigma = 1, float(generate_σ)
        else:
            noise, sigma = 2, 0.0
            funcs.append(generate_σ)
        model = _model_tag(*funcs)
        self._model = model
        # generators built by example_models.fhn.make_generators carry run-time parameters (priors): both must carry
        # the same ones
        gps = [getattr(f, "_mmd_generator_params", None) for f in (generate_x_0, generate_z)]
        if (gps[0] is None) != (gps[1] is None) or (gps[0] is not None and not np.array_equal(gps[0], gps[1])):
            raise ValueError("generate_x_0 and generate_z must come from the same make_generators(...) call")
        self._gen_params = gps[0]
        y_seq = np.asarray(y_seq, dtype=np.float64)
        num_obs, dim_y = y_seq.shape
        dim_v_0 = dim_x if dim_v_0 is None else dim_v_0
        self._bc = BatchedChains(model, obs_interval, num_steps_per_obs, num_obs_per_subseq, y_seq, dim_u, 1,
                                 noise=noise, sigma_fixed=sigma, use_gaussian_splitting=use_gaussian_splitting,
                                 device=device, generator_params=self._gen_params)
        self.num_partition = self._bc.num_partition
        self.dim_q = self._bc.dim_q
        self.model_dict = {
            "dim_u": dim_u, "dim_v": dim_v, "dim_v_0": dim_v_0, "dim_x": dim_x, "dim_y": dim_y,
            "num_obs": num_obs, "num_steps_per_obs": num_steps_per_obs, "δ": obs_interval / num_steps_per_obs,
            "generate_z": generate_z, "generate_x_0": generate_x_0, "generate_σ": generate_σ,
            "forward_func": forward_func, "obs_func": obs_func, "y_seq": y_seq,
        }
        # y_subseqs (:321-356): the observations split per partition into (first, stacked middle, last) blocks
        if num_obs_per_subseq is None or num_obs_per_subseq == num_obs:
            y_subseq_shapes = [((num_obs,),)]
        else:
            y_subseq_shapes = []
            for init in (num_obs_per_subseq, num_obs_per_subseq // 2):
                num_full, num_rem = divmod(num_obs - init, num_obs_per_subseq)
                num_middle = num_full - 1 if num_rem == 0 else num_full
                fin = num_obs_per_subseq if num_rem == 0 else num_rem
                y_subseq_shapes.append(((init,),) + (((num_middle, num_obs_per_subseq),) if num_middle > 0 else ())
                                       + ((fin,),))
        self.y_subseqs = [split_and_reshape(y_seq, shapes) for shapes in y_subseq_shapes]
        self._ctor = dict(model=model, obs_interval=obs_interval, S=num_steps_per_obs, R=num_obs_per_subseq, y=y_seq,
                          dim_u=dim_u, noise=noise, sigma=sigma, gaussian=use_gaussian_splitting, device=device)
        self._batch = {}          # n_states -> BatchedChains used by the batched private handles below
        self._dims = (num_obs, num_steps_per_obs, num_obs_per_subseq, dim_u, dim_x, dim_v, dim_v_0, noise)
        self._resident = None     # (pos bytes, x_obs bytes, partition) of the chain resident on the device
        self._linearised = False
        self._grad = self._ld = None

    # ---- device residency ----------------------------------------------------------------------
    def _upload(self, state):
        key = (np.asarray(state.pos).tobytes(), np.asarray(state.x_obs_seq).tobytes(), int(state.partition))
        if key != self._resident:
            self._bc.set_state(np.asarray(state.pos, dtype=np.float64)[None],
                               np.asarray(state.x_obs_seq, dtype=np.float64)[None], int(state.partition))
            self._resident = key
            self._linearised = False

    def _linearise(self, state):
        """jacob_constr_blocks + chol_gram_blocks + log_det_sqrt_gram + grad_log_det_sqrt_gram in one
        device pass (the reference also computes them together, :1173-1184)."""
        self._upload(state)
        if not self._linearised:
            self._bc.linearize(True)
            self._ld = float(self._bc.log_det_sqrt_gram()[0])
            self._grad = self._bc.grad_log_det_sqrt_gram()[0]
            self._linearised = True

    # ---- cached system functions (:1151-1184) ----------------------------------------------------
    @cache_in_state("pos", "x_obs_seq", "partition")
    def constr(self, state):
        self._upload(state)
        return self._bc.constr()[0]

    @cache_in_state("pos", "x_obs_seq", "partition")
    def jacob_constr_blocks(self, state):
        """Dense ``(dc_du_blocks, dc_dv_blocks, dc_dn_blocks)`` rebuilt on the host from the compressed
        device factors (only tests / diagnostics need them: the device solvers use the factors directly)."""
        self._linearise(state)
        return _dense_jacobian_blocks(self, int(state.partition))

    @cache_in_state("pos", "x_obs_seq", "partition")
    def chol_gram_blocks(self, state):
        """``(chol_C, chol_D_blocks)`` (lower Cholesky factors, reference layout) from the device factors."""
        self._linearise(state)
        return _dense_cholesky_blocks(self, int(state.partition))

    @cache_in_state("pos", "x_obs_seq", "partition")
    def log_det_sqrt_gram(self, state):
        self._linearise(state)
        return self._ld

    @cache_in_state_with_aux(("pos", "x_obs_seq", "partition"), ("log_det_sqrt_gram",))
    def grad_log_det_sqrt_gram(self, state):
        self._linearise(state)
        return np.array(self._grad, copy=True), self._ld

    # ---- private handles with a leading batch axis (:1137-1149) --------------------------------------
    # The reference exposes its jitted closures as `system._constr(q, x_obs_seq, partition)` etc.; the
    # operation-time script maps them over 1000 states with `lax.map`
    # (scripts/fhn_model_noiseless_obs_chmc_operation_times.py:30-65).  Here they take either one state (1-D q) or a
    # batch (2-D q, leading axis = states) and run ONE device launch for the whole batch.  Block factors stay on the
    # device: the Jacobian / Cholesky "blocks" they return are tickets (DeviceArray, zero host bytes) that the
    # follow-up handles accept in place of the dense blocks.  Jacobian, Gram matrices and Cholesky factors are
    # produced by the same fused kernel (k_point), so `_chol_gram_blocks` / `_lu_jacob_product_blocks` of a ticket
    # that is still current cost nothing; the LU products of the Newton solver exist only inside its iteration.
    def _batch_chains(self, n):
        bc = self._batch.get(n)
        if bc is None:
            c = self._ctor
            bc = BatchedChains(c["model"], c["obs_interval"], c["S"], c["R"], c["y"], c["dim_u"], n, noise=c["noise"],
                               sigma_fixed=c["sigma"], use_gaussian_splitting=c["gaussian"], device=c["device"],
                               generator_params=self._gen_params)
            bc._ticket = 0
            self._batch[n] = bc
        return bc

    def _batch_load(self, q, x_obs_seq, partition):
        q = np.asarray(q, dtype=np.float64)
        single = q.ndim == 1
        q2 = q[None] if single else q
        x2 = np.asarray(x_obs_seq, dtype=np.float64)
        x2 = x2[None] if single else x2
        bc = self._batch_chains(q2.shape[0])
        bc.set_state(q2, x2, int(partition))
        bc._ticket += 1
        return bc, single

    @staticmethod
    def _out(a, single):
        a = np.asarray(a)
        return (a[0] if single else a).view(DeviceArray)

    def _tickets(self, bc, single, k):
        out = []
        for _ in range(k):
            t = np.empty((0,) if single else (bc.n_chains, 0)).view(DeviceArray)
            t._mmd_ticket = (bc, bc._ticket, single)
            out.append(t)
        return tuple(out)

    @staticmethod
    def _redeem(*tickets):
        for t in tickets:
            tk = getattr(t, "_mmd_ticket", None)
            if tk is not None:
                bc, version, single = tk
                if version != bc._ticket:
                    raise ValueError("stale block ticket: the states it was computed for are no longer resident")
                return bc, single
        raise TypeError("expected the block tickets returned by _jacob_constr_blocks / _chol_gram_blocks "
                        "(dense blocks are not accepted: the factors live on the device)")

    def _constr(self, q, x_obs_seq, partition):
        bc, single = self._batch_load(q, x_obs_seq, partition)
        return self._out(bc.constr(), single)

    def _generate_x_obs_seq(self, q):
        q = np.asarray(q, dtype=np.float64)
        single = q.ndim == 1
        q2 = q[None] if single else q
        T, X = self._dims[0], self._dims[4]
        bc, _ = self._batch_load(q2, np.zeros((q2.shape[0], T, X)), 0)
        bc.update_x_obs_seq()
        return self._out(bc.get_state()[2], single)

    def _jacob_constr_blocks(self, q, x_obs_seq, partition):
        bc, single = self._batch_load(q, x_obs_seq, partition)
        bc.linearize(False)
        bc.synchronize()
        return self._tickets(bc, single, 3)

    def _chol_gram_blocks(self, dc_du_blocks, dc_dv_blocks, dc_dn_blocks=None):
        bc, single = self._redeem(dc_du_blocks, dc_dv_blocks, dc_dn_blocks)
        return self._tickets(bc, single, 2)

    def _lu_jacob_product_blocks(self, *blocks):
        bc, single = self._redeem(*blocks)
        return self._tickets(bc, single, 2)

    def _log_det_sqrt_gram_from_chol(self, chol_C, chol_D_blocks):
        bc, single = self._redeem(chol_C, chol_D_blocks)
        return self._out(bc.log_det_sqrt_gram(), single)

    def _grad_log_det_sqrt_gram(self, q, x_obs_seq, partition):
        """((log_det_sqrt_gram, (jacob_constr_blocks, chol_gram_blocks)), grad), the layout of
        jax.value_and_grad(..., has_aux=True) at :1143-1146."""
        bc, single = self._batch_load(q, x_obs_seq, partition)
        bc.linearize(True)
        val = self._out(bc.log_det_sqrt_gram(), single)
        grad = self._out(bc.grad_log_det_sqrt_gram(), single)
        return (val, (self._tickets(bc, single, 3), self._tickets(bc, single, 2))), grad

    def _normal_space_component(self, vct, jacob_constr_blocks, chol_gram_blocks):
        bc, single = self._redeem(*jacob_constr_blocks, *chol_gram_blocks)
        v = np.asarray(vct, dtype=np.float64)
        return self._out(bc.normal_space_component(v[None] if single else v), single)

    # ---- Hamiltonian components (:1186-1238) -----------------------------------------------------
    def h1(self, state):
        if self.use_gaussian_splitting:
            return self.log_det_sqrt_gram(state)
        return self.neg_log_dens(state) + self.log_det_sqrt_gram(state)

    def dh1_dpos(self, state):
        if self.use_gaussian_splitting:
            return self.grad_log_det_sqrt_gram(state)
        return self.grad_neg_log_dens(state) + self.grad_log_det_sqrt_gram(state)

    def h2(self, state):
        if self.use_gaussian_splitting:
            return 0.5 * state.pos @ state.pos + 0.5 * state.mom @ state.mom
        return 0.5 * state.mom @ state.mom

    def dh2_dmom(self, state):
        return state.mom

    def dh2_dpos(self, state):
        return state.pos if self.use_gaussian_splitting else 0 * state.pos

    def dh_dpos(self, state):
        if self.use_gaussian_splitting:
            return self.dh1_dpos(state) + self.dh2_dpos(state)
        return self.dh1_dpos(state)

    def h2_flow(self, state, dt):
        if self.use_gaussian_splitting:
            sin_dt, cos_dt = np.sin(dt), np.cos(dt)
            pos = state.pos.copy()
            state.pos = cos_dt * pos + sin_dt * state.mom
            state.mom = cos_dt * state.mom - sin_dt * pos
        else:
            state.pos = state.pos + dt * self.dh2_dmom(state)

    def dh2_flow_dmom(self, dt):
        if self.use_gaussian_splitting:
            return np.sin(dt) * IdentityMatrix(), np.cos(dt) * IdentityMatrix()
        return dt * IdentityMatrix(), IdentityMatrix()

    # ---- :1240-1259 ----------------------------------------------------------------------------------
    def update_x_obs_seq(self, state):
        self._upload(state)
        self._bc.update_x_obs_seq()
        _, _, x = self._bc.get_state()
        state.x_obs_seq = x[0]
        self._resident = None   # x_obs_seq changed on the device: refresh the key on next use

    def normal_space_component(self, state, vct):
        self._linearise(state)
        return self._bc.normal_space_component(np.asarray(vct, dtype=np.float64)[None])[0]

    def project_onto_cotangent_space(self, mom, state):
        mom = mom - self.normal_space_component(state, mom)
        return mom

    def sample_momentum(self, state, rng):
        mom = rng.standard_normal(state.pos.shape)
        return self.project_onto_cotangent_space(mom, state)

    # ---- projection solves on the device ----------------------------------------------------------
    def _project(self, state_prev, q_in, solver, constraint_tol, position_tol, divergence_tol, max_iters):
        self._linearise(state_prev)
        o = self._bc.opts
        o.solver, o.constraint_tol, o.position_tol = solver, constraint_tol, position_tol
        o.divergence_tol, o.max_iters = divergence_tol, int(max_iters)
        q_out, status, iters = self._bc.project_quasi_newton(np.asarray(q_in, dtype=np.float64)[None])
        return q_out[0], int(status[0]), int(iters[0])


def _solve_projection(state, state_prev, dt, system, solver, name, constraint_tol, position_tol, divergence_tol,
                      max_iters):
    q_in = np.array(state.pos, dtype=np.float64, copy=True)
    q_out, status, iters = system._project(state_prev, q_in, solver, constraint_tol, position_tol, divergence_tol,
                                           max_iters)
    if state._call_counts is not None:
        # quasi-Newton counts one constr call per iteration (:1382-1387); Newton also re-evaluates the Jacobian and
        # the LU-factorised block products every iteration (:1451-1461)
        methods = [system.constr] if solver == SOLVER_QUASI_NEWTON else [
            system.constr, system.jacob_constr_blocks, "lu_jacob_product_blocks"]
        for method in methods:
            key = _cache_key_func(system, method)
            state._call_counts[key] = state._call_counts.get(key, 0) + iters
    if status == 0:
        state.pos = q_out
        if state.mom is not None:
            # mu = dq for the identity metric; returned as mu / dt (or mu / sin dt), then
            # state.mom -= dh2_flow_mom_dmom @ mu   (:1060-1063, :1388-1392)
            scale = (np.cos(dt) / np.sin(dt)) if system.use_gaussian_splitting else 1.0 / dt
            state.mom = state.mom - scale * (q_in - q_out)
        return state
    if status & 2:
        raise ConvergenceError(f"{name} iteration diverged on iteration {iters}.")
    raise ConvergenceError(f"{name} iteration did not converge (status {status}) after {iters} iterations.")


def jitted_solve_projection_onto_manifold_quasi_newton(
        state, state_prev, dt, system, constraint_tol=1e-8, position_tol=1e-8, divergence_tol=1e10, max_iters=50):
    """Symmetric quasi-Newton solver for projecting points onto manifold (:1323-1402), on device."""
    return _solve_projection(state, state_prev, dt, system, SOLVER_QUASI_NEWTON, "Quasi-Newton",
                             constraint_tol, position_tol, divergence_tol, max_iters)


def jitted_solve_projection_onto_manifold_newton(
        state, state_prev, dt, system, constraint_tol=1e-8, position_tol=1e-8, divergence_tol=1e10, max_iters=50):
    """Newton solver for projecting points onto manifold (:1405-1476), on device."""
    return _solve_projection(state, state_prev, dt, system, SOLVER_NEWTON, "Newton",
                             constraint_tol, position_tol, divergence_tol, max_iters)


class SwitchPartitionTransition(Transition):
    """Markov transition that deterministically switches conditioned partition (:1262-1282)."""

    def __init__(self, system):
        self.system = system
        self.num_partition = system.num_partition

    state_variables = {"partition", "x_obs_seq"}
    statistic_types = None

    def sample(self, state, rng):
        state.partition = (state.partition + 1) % self.num_partition
        self.system.update_x_obs_seq(state)
        return state, None


class ConditionedDiffusionHamiltonianState(ChainState):
    """Markov chain state for conditioned diffusion Hamiltonian system (:1285-1320)."""

    def __init__(self, pos, x_obs_seq, partition=0, mom=None, dir=1, _call_counts=None, _dependencies=None,
                 _cache=None, _read_only=False):
        if _call_counts is None:
            _call_counts = {}
        super().__init__(pos=pos, x_obs_seq=x_obs_seq, partition=partition, mom=mom, dir=dir,
                         _call_counts=_call_counts, _dependencies=_dependencies, _cache=_cache,
                         _read_only=_read_only)


def find_initial_state_by_linear_interpolation(system, rng, generate_x_obs_seq_init, u=None, v_0=None,
                                               **model_dict):
    """Find an initial constraint satisfying state linearly interpolating noise sequence (:1479-1547):
    the per-step linear solves run on the device (k_init_interp)."""
    md = system.model_dict if not model_dict else model_dict
    u = rng.standard_normal(md["dim_u"]) if u is None else np.asarray(u, dtype=np.float64)
    v_0 = rng.standard_normal(md["dim_v_0"]) if v_0 is None else np.asarray(v_0, dtype=np.float64)
    x_obs_seq = np.asarray(generate_x_obs_seq_init(rng), dtype=np.float64)
    bc = system._bc
    bc.init_linear_interpolation(u[None], v_0[None], x_obs_seq[None], 0)
    q, _, _ = bc.get_state()
    system._resident = None
    state = ConditionedDiffusionHamiltonianState(pos=q[0], x_obs_seq=x_obs_seq)
    state.mom = system.sample_momentum(state, rng)
    return state


def conditioned_diffusion_neg_log_dens_and_grad(obs_interval, num_steps_per_obs, y_seq, dim_u, dim_v_0, dim_v,
                                                forward_func, generate_x_0, generate_z, generate_σ, obs_func,
                                                use_gaussian_splitting=False, return_jax_funcs=False, device=0):
    """Negative log target density + gradient functions for the diffusion model (:82-205), the target of the
    reference's standard-HMC baseline (``mici.systems.EuclideanMetricSystem``): forward simulation and the reverse
    (adjoint) sweep run in one CUDA kernel (``k_hmc_target``).  Returns ``(neg_log_dens, grad_neg_log_dens)`` with
    the reference's conventions: ``grad_neg_log_dens(q) -> (grad, value)``, non-finite values raise
    ``HamiltonianDivergenceError`` (:193-204).  ``return_jax_funcs=True`` returns the unwrapped pair (arrays in,
    arrays out, no exception), the role the raw JAX functions play in the reference."""
    funcs = [forward_func, generate_x_0, generate_z, obs_func]
    if isinstance(generate_σ, Number):
        noise, sigma = 1, float(generate_σ)
    else:
        noise, sigma = 2, 0.0
        funcs.append(generate_σ)
    model = _model_tag(*funcs)
    y_seq = np.asarray(y_seq, dtype=np.float64)
    # (the handle's block structure is irrelevant to this target; any admissible blocking will do)
    bc = BatchedChains(model, obs_interval, num_steps_per_obs, 5 if y_seq.shape[0] > 5 else None, y_seq, dim_u, 1,
                       noise=noise, sigma_fixed=sigma,
                       use_gaussian_splitting=use_gaussian_splitting, device=device)
    if bc.hmc_dim != dim_u + dim_v_0 + dim_v * num_steps_per_obs * y_seq.shape[0]:
        raise ValueError("dim_u / dim_v_0 / dim_v do not match the model")

    def _neg_log_dens(q):
        val, _ = bc.neg_log_dens_and_grad(np.asarray(q, dtype=np.float64)[None], use_gaussian_splitting, with_grad=False)
        return val[0]

    def _grad_neg_log_dens(q):
        val, grad = bc.neg_log_dens_and_grad(np.asarray(q, dtype=np.float64)[None], use_gaussian_splitting)
        return grad[0], val[0]

    if return_jax_funcs:
        return _neg_log_dens, _grad_neg_log_dens

    def neg_log_dens(q):
        val = float(_neg_log_dens(q))
        if not np.isfinite(val):
            raise HamiltonianDivergenceError("Hamiltonian non-finite")
        return val

    def grad_neg_log_dens(q):
        grad, val = _grad_neg_log_dens(q)
        if not np.isfinite(val):
            raise HamiltonianDivergenceError("Hamiltonian non-finite")
        return np.asarray(grad), float(val)

    neg_log_dens._bc = grad_neg_log_dens._bc = bc      # keeps the device handle alive
    return neg_log_dens, grad_neg_log_dens


def find_initial_state_by_gradient_descent_noisy_system(system, rng, adam_step_size=2e-2, max_iters=1000,
                                                        max_init_tries=100, max_num_tries=10, threshold=1.0,
                                                        slow_progress_ratio=0.8, check_iter=100, **model_dict):
    """Find an initial constraint satisfying state by a gradient descent based scheme (:1679-1801): Adam on the
    negative log posterior density of the noisy-observation system until the mean squared residual is below
    ``threshold``; the state on the manifold then takes the residuals as its observation-noise variables.  The
    objective gradients and the Adam updates run on the device (``BatchedChains.init_gradient_descent``)."""
    if not isinstance(system, ConditionedDiffusionConstrainedSystem):
        raise NotImplementedError("device path: pass the ConditionedDiffusionConstrainedSystem of the noisy model")
    bc = system._bc
    q, _ = bc.init_gradient_descent([rng], adam_step_size, max_iters, max_init_tries, max_num_tries, threshold,
                                    slow_progress_ratio, check_iter)
    state = ConditionedDiffusionHamiltonianState(pos=q[0], x_obs_seq=None, _call_counts={})
    system._resident = None
    system.update_x_obs_seq(state)
    state.mom = system.sample_momentum(state, rng)
    return state


class OnlineBlockDiagonalMetricAdapter(Adapter):
    """Block diagonal metric adapter using online covariance estimates (:1804-1931): Welford accumulation of the
    covariance of the first ``dim_param`` position components per chain, Schubert-Gertz / Chan combination across
    chains, Stan-style regularisation towards a scaled identity; sets ``transition.system.metric`` to
    blockdiag(inverse covariance estimate, identity).  Host-side bookkeeping of a few ``dim_param``-sized arrays per
    chain, as in the reference (it belongs to the standard-HMC baseline, not to the constrained path)."""

    is_fast = False

    def __init__(self, dim_param, reg_iter_offset=5, reg_scale=1e-3):
        self.dim_param = dim_param
        self.reg_iter_offset = reg_iter_offset
        self.reg_scale = reg_scale

    def initialize(self, chain_state, transition):
        dtype = chain_state.pos.dtype
        return {
            "iter": 0,
            "mean": np.zeros(shape=(self.dim_param,), dtype=dtype),
            "sum_diff_outer": np.zeros(shape=(self.dim_param, self.dim_param), dtype=dtype),
            "dim_pos": chain_state.pos.shape[0],
        }

    def update(self, adapt_state, chain_state, trans_stats, transition):
        adapt_state["iter"] += 1
        par = chain_state.pos[: self.dim_param]
        pos_minus_mean = par - adapt_state["mean"]
        adapt_state["mean"] += pos_minus_mean / adapt_state["iter"]
        adapt_state["sum_diff_outer"] += pos_minus_mean[None, :] * (par - adapt_state["mean"])[:, None]

    def _regularize_covar_est(self, covar_est, n_iter):
        covar_est *= n_iter / (self.reg_iter_offset + n_iter)
        diag = np.einsum("ii->i", covar_est)
        diag += self.reg_scale * (self.reg_iter_offset / (self.reg_iter_offset + n_iter))

    def finalize(self, adapt_state, transition):
        if isinstance(adapt_state, dict):
            n_iter = adapt_state["iter"]
            covar_est = adapt_state.pop("sum_diff_outer")
            dim_pos = adapt_state["dim_pos"]
        else:
            for i, a in enumerate(adapt_state):
                if i == 0:
                    n_iter = a["iter"]
                    mean_est = a.pop("mean")
                    covar_est = a.pop("sum_diff_outer")
                    dim_pos = a["dim_pos"]
                else:
                    n_iter_prev = n_iter
                    n_iter += a["iter"]
                    mean_diff = mean_est - a["mean"]
                    mean_est *= n_iter_prev
                    mean_est += a["iter"] * a["mean"]
                    mean_est /= n_iter
                    covar_est += a["sum_diff_outer"]
                    covar_est += np.outer(mean_diff, mean_diff) * (a["iter"] * n_iter_prev) / n_iter
        if n_iter < 2:
            raise AdaptationError("At least two chain samples required to compute a variance estimates.")
        covar_est /= n_iter - 1
        self._regularize_covar_est(covar_est, n_iter)
        transition.system.metric = PositiveDefiniteBlockDiagonalMatrix(
            (DensePositiveDefiniteMatrix(covar_est).inv, IdentityMatrix(dim_pos - self.dim_param)))


# ---------------------------------------------------------------------------------------------------
# dense blocks in the reference's layout, rebuilt from the compressed device factors (host, NumPy)
# ---------------------------------------------------------------------------------------------------
def _block_shapes(system, partition):
    T, S, R, U, X, V, V0, noise = system._dims
    if system.num_partition == 1:
        return [(0, T, True, True)]
    init = R if partition == 0 else R // 2
    out = [(0, init, True, False)]
    o = init
    while T - o > R:
        out.append((o, R, False, False))
        o += R
    out.append((o, T - o, False, True))
    return out


def _dense_jacobian_blocks(system, partition):
    """Per block (first, middle..., last): dc_du [rows, U], dc_dv [rows, (V0 +) n*S*V], dc_dn [rows, n] or None.
    Row r of d c / d v_t inside interval k is H_r Phi(t_kr, t_k) K_t."""
    T, S, R, U, X, V, V0, noise = system._dims
    bc = system._bc
    K = bc.get_factor("K")[0]
    Psib = bc.get_factor("Psib")[0]
    A = bc.get_factor("A")[0]
    xend = bc.get_factor("xend")[0]
    q = np.frombuffer(system._resident[0], dtype=np.float64)
    sigma = 0.0 if noise == 0 else (bc._sigma_fixed if noise == 1 else float(np.exp(q[U - 1])))
    dc_du, dc_dv, dc_dn = [], [], []
    for b, (o, n, ini, fin) in enumerate(_block_shapes(system, partition)):
        ny = n if (fin or noise) else n - 1
        nx = 0 if fin else X
        nrows = ny + nx
        Kb = K[b].reshape(-1, S, X, V)[:n]
        Pb = Psib[b].reshape(-1, X, X)[:n]
        Ab = A[b].reshape(-1, U)[:nrows]
        Jv = np.zeros((nrows, (V0 if ini else 0) + n * S * V))
        off = V0 if ini else 0
        for r in range(nrows):
            kr = r if r < ny else n - 1
            h = np.zeros(X)
            if r < ny:
                if system._model == "sir":     # obs_func = exp(x[1]): gradient at the state of observation r
                    h[1] = float(np.exp(xend[b].reshape(-1, X)[r, 1]))
                else:                          # FHN: obs_func = x[0]
                    h[0] = 1.0
            else:
                h[r - ny] = 1.0
            vec = h
            for k in range(kr, -1, -1):
                if k < kr:
                    vec = Pb[k + 1].T @ vec
                Jv[r, off + k * S * V: off + (k + 1) * S * V] = np.einsum("i,tij->tj", vec, Kb[k]).reshape(-1)
            if ini:
                d0 = Pb[0].T @ vec            # d c_r / d x_0; d x_0 / d v_0 = I (FHN) or e_2 (SIR)
                Jv[r, :V0] = d0[:V0] if system._model != "sir" else d0[2:3]
        Jn = None
        if noise:
            Jn = np.zeros((nrows, n))
            Jn[np.arange(ny), np.arange(ny)] = sigma
        dc_du.append(Ab.copy())
        dc_dv.append(Jv)
        dc_dn.append(Jn)
    return tuple(dc_du), tuple(dc_dv), tuple(dc_dn)


def _dense_cholesky_blocks(system, partition):
    T, S, R, U, X, V, V0, noise = system._dims
    bc = system._bc
    L = bc.get_factor("L")[0]
    LC = bc.get_factor("LC")[0]

    def unpack(packed, n):
        M = np.zeros((n, n))
        M[np.tril_indices(n)] = packed[: n * (n + 1) // 2]
        M[np.diag_indices(n)] = 1.0 / np.diag(M)   # the device stores the diagonal inverted
        return M

    chol_D = []
    for b, (o, n, ini, fin) in enumerate(_block_shapes(system, partition)):
        nrows = (n if (fin or noise) else n - 1) + (0 if fin else X)
        chol_D.append(unpack(L[b], nrows))
    return unpack(LC, U), tuple(chol_D)
