"""Batched (many-chain) host interface to the CUDA constrained-HMC path.

``BatchedChains`` mirrors, for a batch of chains held resident in HBM, the per-state methods of the
reference's ``ConditionedDiffusionConstrainedSystem`` (``sde/mici_extensions.py:1151-1259``) and one
``ConstrainedLeapfrogIntegrator.step``.  Array arguments are NumPy float64 in the reference's
per-chain layout (``q`` is ``[n_chains, dim_q]``); they are passed to C zero-copy.
"""

import ctypes as C

import numpy as np

from ._lib import MmdConfig, MmdIntegratorOpts, check, lib

MODEL_IDS = {"fhn": 0, "sir": 1, "fhn_notebook": 2}
STATUS_NOT_CONVERGED, STATUS_DIVERGED, STATUS_NON_REVERSIBLE, STATUS_NON_FINITE = 1, 2, 4, 8


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class BatchedChains:
    def __init__(
        self,
        model,
        obs_interval,
        num_steps_per_obs,
        num_obs_per_subseq,
        y_seq,
        dim_u,
        n_chains,
        noise=0,
        sigma_fixed=0.0,
        use_gaussian_splitting=False,
        device=0,
        generator_params=None,
    ):
        """generator_params: run-time parameters of the model's generate_z / generate_x_0 (FHN: the 22 values of
        `example_models.fhn.generator_params(...)`; None keeps the reference's generators)."""
        self._L = lib()
        self._y = _c(np.asarray(y_seq).reshape(-1))
        cfg = MmdConfig(
            MODEL_IDS[model],
            int(np.asarray(y_seq).shape[0]),
            int(num_steps_per_obs),
            int(num_obs_per_subseq) if num_obs_per_subseq else 0,
            int(dim_u),
            int(noise),
            float(sigma_fixed),
            int(bool(use_gaussian_splitting)),
            float(obs_interval),
            _dp(self._y),
            int(n_chains),
            int(device),
        )
        self._h = C.c_void_p()
        check(self._L.mmd_create(C.byref(cfg), C.byref(self._h)))
        self.n_chains = int(n_chains)
        self._model = model
        self._dim_u = int(dim_u)
        self._dim_v_0 = {"fhn": 2, "fhn_notebook": 2, "sir": 1}[model]
        self._sigma_fixed = float(sigma_fixed)
        self.dim_q = self._L.mmd_dim_q(self._h)
        self.num_partition = self._L.mmd_num_partition(self._h)
        self.num_obs = cfg.num_obs
        self.opts = MmdIntegratorOpts()
        self._L.mmd_default_integrator_opts(C.byref(self.opts))
        if generator_params is not None:
            self.set_generator_params(generator_params)

    # ---- run-time generator parameters (priors) ---------------------------------------------
    def set_generator_params(self, params):
        params = _c(np.asarray(params).reshape(-1))
        check(self._L.mmd_set_generator_params(self._h, _dp(params), int(params.size)))

    def get_generator_params(self):
        out = np.empty(max(self._L.mmd_num_generator_params(self._h), 0))
        if out.size:
            check(self._L.mmd_get_generator_params(self._h, _dp(out)))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.mmd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state -------------------------------------------------------------------------
    @property
    def partition(self):
        return self._L.mmd_get_partition(self._h)

    def set_chain_offset(self, chain0):
        """Global index of this handle's first chain (keeps Philox streams disjoint across ranks)."""
        check(self._L.mmd_set_chain_offset(self._h, int(chain0)))

    def chains_per_tile(self):
        return int(self._L.mmd_chains_per_tile(self._h))

    def num_constraints(self, partition=None):
        return self._L.mmd_num_constraints(self._h, self.partition if partition is None else partition)

    def set_state(self, q, x_obs_seq, partition=0, p=None, blocking=True):
        """Upload positions, conditioned states and (optionally) momenta.  blocking=False returns before the copies
        have finished: the arrays must then be page-locked, C-contiguous float64 and stay untouched until
        `synchronize()` (or any call that reads results back); it lets one host thread overlap the upload for one
        BatchedChains object with the computation of another."""
        if blocking:
            q, x = _c(q), _c(x_obs_seq)
            pp = None if p is None else _c(p)
        else:
            q, x, pp = q, x_obs_seq, p
            for a in (q, x) + (() if pp is None else (pp,)):
                if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
                    raise ValueError("set_state(blocking=False) needs C-contiguous float64 arrays (no hidden copies)")
        assert q.shape == (self.n_chains, self.dim_q), q.shape
        assert x.shape[0] == self.n_chains and x.size == self.n_chains * x.shape[1] * x.shape[2]
        fn = self._L.mmd_set_state if blocking else self._L.mmd_set_state_async
        check(fn(self._h, _dp(q), None if pp is None else _dp(pp), _dp(x), int(partition)))
        self._dim_x = x.shape[2]

    # ---- zero-copy device path (DLPack) ---------------------------------------------------
    @property
    def stream(self):
        """The handle's cudaStream_t as an int (what a DLPack producer expects in ``__dlpack__(stream=...)``)."""
        return int(self._L.mmd_get_stream(self._h) or 0)

    def set_state_dlpack(self, q, x_obs_seq, partition=0, p=None):
        """`set_state` from CUDA float64 arrays of any DLPack producer (torch, CuPy, JAX ...) in the reference layout
        (`q`, `p`: [n_chains, dim_q]; `x_obs_seq`: [n_chains, T, dim_x]) without a host round trip: the device
        pointers go to ``mmd_set_state_dev``, which re-tiles on the GPU.  Ordering follows the DLPack protocol: the
        producers are handed this object's stream and make their pending writes visible to it; the borrowed tensors
        are released after the packing kernels have been queued and the stream has been synchronised."""
        from ._dlpack import DeviceArray

        dev = int(self._L.mmd_get_device(self._h))
        st = self.stream or 1
        arrs = []
        try:
            qa = DeviceArray(q, st, dev, (self.n_chains, self.dim_q), "q")
            arrs.append(qa)
            xa = DeviceArray(x_obs_seq, st, dev, None, "x_obs_seq")
            arrs.append(xa)
            if xa.shape[0] != self.n_chains or len(xa.shape) != 3 or xa.shape[1] != self.num_obs:
                raise ValueError(f"x_obs_seq: shape {xa.shape} is not [n_chains, num_obs, dim_x]")
            pa = None
            if p is not None:
                pa = DeviceArray(p, st, dev, (self.n_chains, self.dim_q), "p")
                arrs.append(pa)
            check(self._L.mmd_set_state_dev(self._h, qa.ptr, None if pa is None else pa.ptr, xa.ptr, int(partition)))
            self._dim_x = int(xa.shape[2])
            self.synchronize()      # the producers may reuse their buffers as soon as we return
        finally:
            for a in arrs:
                a.release()

    def get_state_dlpack(self, q_out=None, p_out=None, x_out=None):
        """`get_state` into caller-provided CUDA float64 arrays (DLPack producers, reference layout): device-to-device
        un-tiling via ``mmd_get_state_dev``, no host copy.  Returns after the handle's stream has finished."""
        from ._dlpack import DeviceArray

        dev = int(self._L.mmd_get_device(self._h))
        st = self.stream or 1
        arrs = []
        try:
            ptrs = []
            for obj, shape, nm in ((q_out, (self.n_chains, self.dim_q), "q_out"), (p_out, (self.n_chains, self.dim_q), "p_out"),
                                   (x_out, (self.n_chains, self.num_obs, self._dim_x), "x_out")):
                if obj is None:
                    ptrs.append(None)
                else:
                    a = DeviceArray(obj, st, dev, shape, nm)
                    arrs.append(a)
                    ptrs.append(a.ptr)
            check(self._L.mmd_get_state_dev(self._h, *ptrs))
            self.synchronize()
        finally:
            for a in arrs:
                a.release()

    def get_state_into(self, q, p, x_obs_seq, blocking=True):
        """`get_state` into caller-owned host arrays (C-contiguous float64; page-locked for real overlap).
        blocking=False queues the copies on the handle's stream and returns: valid after `synchronize()`."""
        for a in (q, p, x_obs_seq):
            if a is not None and not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
                raise ValueError("get_state_into needs C-contiguous float64 arrays")
        fn = self._L.mmd_get_state if blocking else self._L.mmd_get_state_async
        check(fn(self._h, None if q is None else _dp(q), None if p is None else _dp(p),
                 None if x_obs_seq is None else _dp(x_obs_seq)))

    def set_chain_regrouping(self, on=True):
        """Throughput option: at every partition switch re-assign the chains to CTA tiles sorted by the projection
        iteration count of their last step (slow chains then delay only each other).  Per-chain results are
        bit-identical; per-chain outputs (get_state, transition_stats, ...) come in SLOT order afterwards and
        `slot_chains()` gives the chain id of every slot: canonical[slot_chains()] = slot_ordered."""
        check(self._L.mmd_set_chain_regrouping(self._h, int(bool(on))))

    def slot_chains(self):
        out = np.empty(self.n_chains, dtype=np.int32)
        check(self._L.mmd_get_slot_chains(self._h, out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def init_linear_interpolation(self, u, v_0, x_obs_seq, partition=0):
        """Batched `find_initial_state_by_linear_interpolation` (mici_extensions.py:1479-1547)."""
        u, v_0, x = _c(u), _c(v_0), _c(x_obs_seq)
        self._dim_x = x.shape[2]
        check(self._L.mmd_init_linear_interpolation(self._h, _dp(u), _dp(v_0), _dp(x), int(partition)))

    # ---- standard-HMC target and the Adam initialiser (noisy-observation systems) --------------------
    @property
    def hmc_dim(self):
        """dim_u + dim_v_0 + T S dim_v: the position of the standard-HMC target has no noise variables."""
        return int(self._L.mmd_hmc_dim(self._h))

    def neg_log_dens_and_grad(self, q, use_gaussian_splitting=False, with_grad=True, with_residuals=False):
        """Batched ``conditioned_diffusion_neg_log_dens_and_grad`` (mici_extensions.py:82-205): q [n_chains, hmc_dim]
        -> (value [n], gradient [n, hmc_dim] or None[, residuals [n, T]])."""
        q = _c(q)
        assert q.shape == (self.n_chains, self.hmc_dim), q.shape
        val = np.empty(self.n_chains)
        grad = np.empty_like(q) if with_grad else None
        res = np.empty((self.n_chains, self.num_obs)) if with_residuals else None
        check(self._L.mmd_hmc_target(self._h, _dp(q), int(not use_gaussian_splitting), _dp(val),
                                     None if grad is None else _dp(grad), None if res is None else _dp(res)))
        return (val, grad, res) if with_residuals else (val, grad)

    def init_gradient_descent(self, rngs, adam_step_size=2e-2, max_iters=1000, max_init_tries=100, max_num_tries=10,
                              threshold=1.0, slow_progress_ratio=0.8, check_iter=100, partition=0):
        """Batched ``find_initial_state_by_gradient_descent_noisy_system`` (mici_extensions.py:1679-1801): every
        chain runs the reference's procedure with its own generator ``rngs[c]`` -- draw u_v ~ N(0, I) until the
        residuals are finite (:1745-1752), Adam on 1/2 |r|^2 + T log sigma + 1/2 |u_v|^2 (:1701-1731) until the mean
        squared residual of the CURRENT iterate is below ``threshold`` (:1757-1763), restart on divergence (:1760-1762)
        or slow progress (:1780-1788) -- with all gradient evaluations and Adam updates on the device.  On success
        the resident position is [u_v | residuals] (:1767-1771) with x_obs_seq regenerated; returns (q, tries)."""
        n, dim, T = self.n_chains, self.hmc_dim, self.num_obs
        u_v = np.empty((n, dim))
        tries = np.zeros(n, dtype=np.int32)         # Adam runs started per chain
        done = np.zeros(n, dtype=bool)
        it = np.zeros(n, dtype=np.int64)
        prev = np.zeros(n)
        q_out = np.empty((n, self.dim_q))

        def fresh(idx):
            # a finite starting point per chain (:1745-1755)
            need = np.array(idx, dtype=np.int64)
            init_tries = np.zeros(n, dtype=np.int64)
            while need.size:
                for c in need:
                    u_v[c] = rngs[c].standard_normal(dim)
                    init_tries[c] += 1
                _, _, res = self.neg_log_dens_and_grad(u_v_safe(), with_grad=False, with_residuals=True)
                ok = np.isfinite(res).all(axis=1)
                need = np.array([c for c in need if not ok[c]], dtype=np.int64)
                if need.size and init_tries[need].max() >= max_init_tries:
                    raise RuntimeError(f"Did not find valid initial state in {max_init_tries} tries.")
            mask = np.zeros(n, dtype=np.int32)
            mask[idx] = 1
            check(self._L.mmd_adam_begin(self._h, _dp(u_v), _ip(mask)))
            it[idx] = 0
            tries[idx] += 1
            _, _, res = self.neg_log_dens_and_grad(u_v_safe(), with_grad=False, with_residuals=True)
            prev[idx] = np.mean(res[idx] ** 2, axis=1)

        def u_v_safe():
            # chains that are finished keep their last (finite) row; rows never drawn yet are zero
            return np.where(np.isfinite(u_v), u_v, 0.0)

        u_v[:] = 0.0
        fresh(list(range(n)))
        msr = np.empty(n)
        while not done.all():
            check(self._L.mmd_adam_eval(self._h, _dp(msr), None))
            active = ~done
            diverged = active & ~np.isfinite(msr)
            found = active & np.isfinite(msr) & (msr < threshold)
            if found.any():
                cur = np.empty((n, dim))
                res = np.empty((n, T))
                check(self._L.mmd_adam_get(self._h, _dp(cur), _dp(res)))
                for c in np.nonzero(found)[0]:
                    q_out[c] = np.concatenate([cur[c], res[c]])
                    u_v[c] = cur[c]
                done |= found
            cont = active & ~diverged & ~found
            # slow-progress test on the iterations the reference checks (:1777-1792)
            chk = cont & (it % check_iter == 0)
            slow = chk & (it > 0) & (it < max_iters // 2) & (msr / np.where(prev > 0, prev, 1.0) > slow_progress_ratio)
            prev = np.where(chk & ~slow, msr, prev)
            upd = cont & ~slow
            exhausted = upd & (it + 1 >= max_iters)
            if upd.any():
                check(self._L.mmd_adam_update(self._h, float(adam_step_size), _ip(upd.astype(np.int32))))
                it[upd] += 1
            restart = diverged | slow | exhausted
            if restart.any():
                idx = [int(c) for c in np.nonzero(restart)[0]]
                if (tries[idx] >= max_num_tries).any():
                    raise RuntimeError(f"Did not find valid state in {max_num_tries} tries.")
                fresh(idx)
        xo = np.zeros((n, T, self._dim_x_model()))
        self.set_state(q_out, xo, partition)
        self.update_x_obs_seq()
        return q_out, tries

    def _dim_x_model(self):
        return {"fhn": 2, "fhn_notebook": 2, "sir": 3}[self._model]

    def set_momentum(self, p):
        check(self._L.mmd_set_momentum(self._h, _dp(_c(p))))

    def get_state(self):
        q = np.empty((self.n_chains, self.dim_q))
        p = np.empty((self.n_chains, self.dim_q))
        x = np.empty((self.n_chains, self.num_obs, self._dim_x))
        check(self._L.mmd_get_state(self._h, _dp(q), _dp(p), _dp(x)))
        return q, p, x

    def get_head(self, dim_v_0=None):
        """(u, v_0) of every chain's current position -- what the reference scripts' trace functions need
        (`generate_z(u)`, `generate_x_0(z, v_0)`) -- without reading back the whole position."""
        rows = self.dim_q - self._body_and_noise_rows()
        out = np.empty((self.n_chains, rows))
        check(self._L.mmd_get_head(self._h, _dp(out)))
        du = self._dim_u
        return out[:, :du], out[:, du:]

    def _body_and_noise_rows(self):
        return self.dim_q - (self._dim_u + self._dim_v_0)

    # ---- system ops --------------------------------------------------------------------
    def linearize(self, with_grad=True):
        check(self._L.mmd_linearize(self._h, int(with_grad)))

    def constr(self):
        out = np.empty((self.n_chains, self.num_constraints()))
        check(self._L.mmd_constr(self._h, _dp(out)))
        return out

    def log_det_sqrt_gram(self):
        out = np.empty(self.n_chains)
        check(self._L.mmd_log_det_sqrt_gram(self._h, _dp(out)))
        return out

    def grad_log_det_sqrt_gram(self):
        out = np.empty((self.n_chains, self.dim_q))
        check(self._L.mmd_grad_log_det_sqrt_gram(self._h, _dp(out)))
        return out

    def hamiltonian(self):
        out = np.empty(self.n_chains)
        check(self._L.mmd_hamiltonian(self._h, _dp(out)))
        return out

    def project_momentum(self):
        check(self._L.mmd_project_momentum(self._h))

    def normal_space_component(self, vct):
        vct = _c(vct)
        out = np.empty_like(vct)
        check(self._L.mmd_normal_space_component(self._h, _dp(vct), _dp(out)))
        return out

    def update_x_obs_seq(self):
        check(self._L.mmd_update_x_obs_seq(self._h))

    def switch_partition(self):
        check(self._L.mmd_switch_partition(self._h))

    def sample_momentum(self, seed, offset=0):
        check(self._L.mmd_sample_momentum(self._h, int(seed), int(offset)))

    def get_factor(self, name):
        rows = C.c_int()
        check(self._L.mmd_get_factor(self._h, name.encode(), None, C.byref(rows)))
        nb = self._L.mmd_num_blocks(self._h, self.partition)
        shape = (self.n_chains, rows.value) if name == "LC" else (self.n_chains, nb, rows.value)
        out = np.empty(shape)
        check(self._L.mmd_get_factor(self._h, name.encode(), _dp(out), C.byref(rows)))
        return out

    # ---- integrator --------------------------------------------------------------------
    def leapfrog_step(self, dt, n_inner_step=1):
        """One ConstrainedLeapfrogIntegrator.step of every chain; n_inner_step > 1 splits the h2 flow + projection
        into inner steps of size dt / n_inner_step (Mici's n_inner_step, scripts/utils.py:132)."""
        if n_inner_step == 1:
            check(self._L.mmd_leapfrog_step(self._h, float(dt), C.byref(self.opts)))
        else:
            check(self._L.mmd_leapfrog_step_inner(self._h, float(dt), int(n_inner_step), C.byref(self.opts)))

    def hmc_transition(self, dt, n_leapfrog, seed, it, switch_partition=True):
        """Momentum refresh + static constrained trajectory + Metropolis accept (+ partition switch)."""
        check(
            self._L.mmd_hmc_transition(
                self._h, float(dt), int(n_leapfrog), int(seed), int(it), C.byref(self.opts), int(switch_partition)
            )
        )

    def transition_begin(self, seed, it):
        check(self._L.mmd_transition_begin(self._h, int(seed), int(it)))

    def transition_steps(self, dt, n_steps=1):
        check(self._L.mmd_transition_steps(self._h, float(dt), int(n_steps), C.byref(self.opts)))

    def transition_end(self, seed, it, switch_partition=True):
        check(self._L.mmd_transition_end(self._h, int(seed), int(it), int(switch_partition)))

    def set_step_sizes(self, dt):
        """Per-chain step sizes (array [n_chains]) or None to return to the scalar step size."""
        check(self._L.mmd_set_step_sizes(self._h, None if dt is None else _dp(_c(np.broadcast_to(dt, (self.n_chains,))))))

    def get_step_sizes(self):
        out = np.empty(self.n_chains)
        check(self._L.mmd_get_step_sizes(self._h, _dp(out)))
        return out

    def adapt_start(self, init_step_size, target=0.8, reg_coefficient=0.05, iter_decay=0.75, iter_offset=10):
        """On-device per-chain DualAveragingStepSizeAdapter (Mici defaults; scripts use target 0.8, reg 0.1)."""
        if np.ndim(init_step_size) == 0:
            check(self._L.mmd_adapt_start(self._h, float(init_step_size), float(target), float(reg_coefficient),
                                          float(iter_decay), float(iter_offset)))
        else:   # one initial step size per chain (adaptation.find_init_step_sizes)
            init = _c(np.broadcast_to(init_step_size, (self.n_chains,)))
            check(self._L.mmd_adapt_start_per_chain(self._h, _dp(init), float(target), float(reg_coefficient),
                                                    float(iter_decay), float(iter_offset)))

    def adapt_stop(self, pool=False):
        check(self._L.mmd_adapt_stop(self._h, int(pool)))

    def adapt_update(self, accept_stat):
        check(self._L.mmd_adapt_update(self._h, _dp(_c(accept_stat))))

    # ---- vector primitives for host-driven tree building (see nuts.py) -------------------
    VEC_Q, VEC_P = -1, -2

    def aux_reserve(self, n_arrays):
        check(self._L.mmd_aux_reserve(self._h, int(n_arrays)))

    @staticmethod
    def _mask(mask):
        return None if mask is None else np.ascontiguousarray(mask, dtype=np.int32)

    def vec_axpby(self, dst, src, alpha=1.0, beta=0.0, mask=None):
        m = self._mask(mask)
        check(self._L.mmd_vec_axpby(self._h, int(dst), int(src), float(alpha), float(beta), None if m is None else _ip(m)))

    def vec_uturn(self, a, d, c, e):
        o1, o2 = np.empty(self.n_chains), np.empty(self.n_chains)
        check(self._L.mmd_vec_uturn(self._h, int(a), int(d), int(c), int(e), _dp(o1), _dp(o2)))
        return o1, o2

    def set_inactive(self, mask=None, clear_errors=False):
        m = self._mask(mask)
        check(self._L.mmd_set_inactive(self._h, None if m is None else _ip(m), int(clear_errors)))

    def relinearize(self):
        check(self._L.mmd_relinearize(self._h))

    def successful_steps(self, reset=False):
        return int(self._L.mmd_successful_steps(self._h, int(reset)))

    def total_qn_iterations(self, reset=False):
        return int(self._L.mmd_total_qn_iterations(self._h, int(reset)))

    def profile_enable(self, on, max_launches=4096):
        check(self._L.mmd_profile_enable(self._h, int(on), int(max_launches)))

    def profile_summary(self, kernel_id):
        n = C.c_int()
        ms = C.c_double()
        check(self._L.mmd_profile_summary(self._h, int(kernel_id), C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def transition_stats(self):
        acc = np.empty(self.n_chains, dtype=np.int32)
        ap = np.empty(self.n_chains)
        st = np.empty(self.n_chains, dtype=np.int32)
        check(self._L.mmd_get_transition_stats(self._h, _ip(acc), _dp(ap), _ip(st)))
        return {"accepted": acc, "accept_stat": ap, "status": st}

    def step_info(self):
        st = np.empty(self.n_chains, dtype=np.int32)
        i0 = np.empty(self.n_chains, dtype=np.int32)
        i1 = np.empty(self.n_chains, dtype=np.int32)
        rv = np.empty(self.n_chains)
        check(self._L.mmd_get_step_info(self._h, _ip(st), _ip(i0), _ip(i1), _dp(rv)))
        return {"status": st, "iters_fwd": i0, "iters_rev": i1, "rev_dist": rv}

    def project_quasi_newton(self, q_in, dt=1.0):
        q_in = _c(q_in)
        out = np.empty_like(q_in)
        st = np.empty(self.n_chains, dtype=np.int32)
        it = np.empty(self.n_chains, dtype=np.int32)
        check(
            self._L.mmd_project_quasi_newton(
                self._h, _dp(q_in), float(dt), C.byref(self.opts), _dp(out), _ip(st), _ip(it)
            )
        )
        return out, st, it

    # ---- instrumentation ----------------------------------------------------------------
    def launch_count(self):
        return int(self._L.mmd_launch_count(self._h))

    def timer_start(self):
        check(self._L.mmd_timer_start(self._h))

    def timer_stop_ms(self):
        ms = C.c_float()
        check(self._L.mmd_timer_stop_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def synchronize(self):
        check(self._L.mmd_synchronize(self._h))
