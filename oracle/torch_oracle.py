"""CPU float64 oracle for the constrained-HMC hot path (TEST INFRASTRUCTURE ONLY).

Line-by-line restatement of ``sde/mici_extensions.py`` of the reference
(``ConditionedDiffusionConstrainedSystem`` :208-1259, the projection solvers :1323-1476, the
linear-interpolation initialiser :1479-1547) with ``torch`` / ``torch.func`` (float64, CPU) in place
of JAX, plus the step order of Mici 0.1.10's ``ConstrainedLeapfrogIntegrator`` (un-vendored
third-party dependency, ``requirements.txt:1``; restated from its published algorithm, see
SURVEY.md section 3.3).  All derivatives come from automatic differentiation exactly where the
reference uses ``jax.jacrev`` / ``jax.value_and_grad``, so this file is independent of the
hand-derived adjoint recursions in the CUDA kernels.

PARITY STATUS: the reference ships no golden vectors or tests and none of its pinned dependencies
(mici 0.1.10, symnum 0.1.2, jax 0.2.21) can be installed in this environment.  Since round 2 this
restatement is pinned, at trace level, to the REFERENCE'S OWN SOURCE FILE: ``oracle/reference_runner.py``
executes ``/root/reference/sde/mici_extensions.py`` unmodified with a torch-backed ``jax`` stand-in
(``oracle/jax_torch_shim.py``) and ``tests/test_reference_pin_cpu.py`` compares it with this file quantity by
quantity (constraint, log-det, its gradient, cotangent projection, Hamiltonian, whole leapfrog steps with both
solvers; 1e-10 .. 1e-13), and the GPU tests compare the CUDA path with vectors generated from that run
(``tests/golden/reference_pin_golden.npz``).  The model callables of that run are the reference's own
``sde/example_models/*.py`` executed through a SymPy-backed ``symnum`` stand-in (``oracle/symnum_sympy_shim.py``).
What is substituted, and therefore NOT pinned: the array / code-generation libraries (torch and SymPy instead of XLA
and SymNum) and Mici's integrator / state classes (``mici_compat`` stand-ins of Mici 0.1.10).  Further
pins: the invariants of ``tests/test_oracle_invariants.py`` and the notebook's recorded posterior.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this.
"""

from numbers import Number

import numpy as onp
import torch
from torch.func import jacrev, vmap, grad

torch.set_default_dtype(torch.float64)


def _t(a):
    return a if isinstance(a, torch.Tensor) else torch.as_tensor(onp.asarray(a), dtype=torch.float64)


def split(v, lengths):
    """mici_extensions.py:31-40."""
    i = 0
    parts = []
    for j in lengths:
        parts.append(v[i : i + j])
        i += j
    if i < len(v):
        parts.append(v[i:])
    return parts


def split_and_reshape(array, shapes):
    """mici_extensions.py:43-53 (`onp.product` -> `onp.prod`)."""
    i = 0
    parts = []
    for s in shapes:
        j = int(onp.prod(s))
        parts.append(array[i : i + j].reshape(tuple(s) + tuple(array.shape[1:])))
        i += j
    if i < array.shape[0]:
        parts.append(array[i:])
    return parts


def _scan(step, x_0, v_seq):
    """`lax.scan(lambda x, v: (f(x, v),) * 2, x_0, v_seq)` -> stacked states."""
    xs = []
    x = x_0
    for t in range(v_seq.shape[0]):
        x = step(x, v_seq[t])
        xs.append(x)
    return torch.stack(xs)


class ConvergenceError(Exception):
    pass


class NonReversibleStepError(Exception):
    pass


class OracleSystem:
    """Restatement of ConditionedDiffusionConstrainedSystem (mici_extensions.py:208-1259).

    Identity metric only (the only metric the CHMC scripts construct, scripts/utils.py:255-270)."""

    def __init__(
        self,
        obs_interval,
        num_steps_per_obs,
        num_obs_per_subseq,
        y_seq,
        dim_u,
        dim_x,
        dim_v,
        forward_func,
        generate_x_0,
        generate_z,
        obs_func,
        generate_σ=None,
        use_gaussian_splitting=False,
        dim_v_0=None,
    ):
        y_seq = _t(y_seq)
        self.use_gaussian_splitting = use_gaussian_splitting
        log_det_sqrt_metric_0 = 0
        # mici_extensions.py:317-352
        num_obs, dim_y = y_seq.shape
        δ = obs_interval / num_steps_per_obs
        num_step = num_obs * num_steps_per_obs
        obs_indices = slice(num_steps_per_obs - 1, None, num_steps_per_obs)
        if num_obs_per_subseq is None or num_obs_per_subseq == num_obs:
            y_subseq_shapes = [((num_obs,),)]
            v_subseq_shapes = [((num_obs * num_steps_per_obs,),)]
            subseqs_are_batched = [(False,)]
        else:
            y_subseq_shapes, v_subseq_shapes, subseqs_are_batched = [], [], []
            for init_subseq_size in [num_obs_per_subseq, num_obs_per_subseq // 2]:
                num_full, num_remaining = divmod(num_obs - init_subseq_size, num_obs_per_subseq)
                num_middle = num_full - 1 if num_remaining == 0 else num_full
                final_subseq_size = num_obs_per_subseq if num_remaining == 0 else num_remaining
                y_subseq_shapes.append(
                    ((init_subseq_size,),)
                    + (((num_middle, num_obs_per_subseq),) if num_middle > 0 else ())
                    + ((final_subseq_size,),)
                )
                v_subseq_shapes.append(
                    ((init_subseq_size * num_steps_per_obs,),)
                    + (
                        ((num_middle, num_obs_per_subseq * num_steps_per_obs),)
                        if num_middle > 0
                        else ()
                    )
                    + ((final_subseq_size * num_steps_per_obs,),)
                )
                subseqs_are_batched.append(
                    (False, True, False) if num_middle > 0 else (False, False)
                )
        y_subseqs = [split_and_reshape(y_seq, shapes) for shapes in y_subseq_shapes]
        noisy_observations = generate_σ is not None
        if generate_σ is not None and isinstance(generate_σ, Number):
            σ_const = generate_σ

            def generate_σ(u):
                return σ_const

        dim_v_0 = dim_x if dim_v_0 is None else dim_v_0
        self.y_subseqs = y_subseqs
        self.y_subseq_shapes = y_subseq_shapes
        self.v_subseq_shapes = v_subseq_shapes
        self.subseqs_are_batched = subseqs_are_batched
        self.num_partition = len(y_subseqs)
        self.noisy_observations = noisy_observations
        self.dim_q = dim_u + dim_v_0 + num_step * dim_v + (num_obs * dim_y if noisy_observations else 0)
        self.model_dict = {
            "dim_u": dim_u,
            "dim_v": dim_v,
            "dim_v_0": dim_v_0,
            "dim_x": dim_x,
            "dim_y": dim_y,
            "num_obs": num_obs,
            "num_steps_per_obs": num_steps_per_obs,
            "δ": δ,
            "generate_z": generate_z,
            "generate_x_0": generate_x_0,
            "generate_σ": generate_σ,
            "forward_func": forward_func,
            "obs_func": obs_func,
            "y_seq": y_seq,
        }

        def step_func(z, x, v):  # :379-382
            return forward_func(z, x, v, δ)

        def generate_x_obs_seq(q):  # :384-397
            if noisy_observations:
                u, v_0, v_seq_flat, _ = split(
                    q, (dim_u, dim_v_0, num_obs * num_steps_per_obs * dim_v)
                )
            else:
                u, v_0, v_seq_flat = split(q, (dim_u, dim_v_0))
            z = generate_z(u)
            x_0 = generate_x_0(z, v_0)
            v_seq = v_seq_flat.reshape((-1, dim_v))
            x_seq = _scan(lambda x, v: step_func(z, x, v), x_0, v_seq)
            return x_seq[obs_indices]

        def generate_y_bar(z, w_0, v_seq, σ_n_seq, initial_subseq, final_subseq):  # :399-411
            x_0 = generate_x_0(z, w_0) if initial_subseq else w_0
            x_seq = _scan(lambda x, v: step_func(z, x, v), x_0, v_seq)
            y_seq_ = obs_func(x_seq[obs_indices])
            if noisy_observations:
                y_seq_ = y_seq_ + σ_n_seq
            if final_subseq:
                return y_seq_.flatten()
            elif noisy_observations:
                return torch.cat((y_seq_.flatten(), x_seq[-1]))
            else:
                return torch.cat((y_seq_[:-1].flatten(), x_seq[-1]))

        def partition_into_subseqs(v_seq, v_0, n_seq, x_obs_seq, partition=0):  # :413-471
            end_y = None if noisy_observations else -1
            partition_size = len(y_subseq_shapes[partition])
            v_subseqs = split_and_reshape(v_seq, v_subseq_shapes[partition])
            if noisy_observations:
                n_subseqs = split_and_reshape(n_seq, y_subseq_shapes[partition])
            else:
                n_subseqs = (None,) * partition_size
            x_obs_subseqs = split_and_reshape(x_obs_seq, y_subseq_shapes[partition])
            w_inits = [v_0]
            prev_batched = False
            for b in range(1, partition_size):
                if subseqs_are_batched[partition][b]:
                    w_inits.append(
                        torch.vstack(
                            [
                                x_obs_subseqs[b - 1][(-1, -1) if prev_batched else -1],
                                x_obs_subseqs[b][:-1, -1],
                            ]
                        )
                    )
                    prev_batched = True
                else:
                    w_inits.append(x_obs_subseqs[b - 1][(-1, -1) if prev_batched else (-1,)])
                    prev_batched = False
            y_bars = []
            for b in range(0, partition_size - 1):
                if subseqs_are_batched[partition][b]:
                    y_bars.append(
                        torch.cat(
                            (
                                y_subseqs[partition][b][:, :end_y].reshape(
                                    (y_subseqs[partition][b].shape[0], -1)
                                ),
                                x_obs_subseqs[b][:, -1],
                            ),
                            -1,
                        )
                    )
                else:
                    y_bars.append(
                        torch.cat(
                            (y_subseqs[partition][b][:end_y].flatten(), x_obs_subseqs[b][-1])
                        )
                    )
            y_bars.append(y_subseqs[partition][-1].flatten())
            return v_subseqs, n_subseqs, w_inits, y_bars

        def _split_q(q):
            if noisy_observations:
                u, v_0, v_seq_flat, n_flat = split(
                    q, (dim_u, dim_v_0, num_step * dim_v, num_obs * dim_y)
                )
                n_seq = n_flat.reshape((-1, dim_y))
            else:
                u, v_0, v_seq_flat = split(q, (dim_u, dim_v_0))
                n_seq = None
            return u, v_0, v_seq_flat.reshape((-1, dim_v)), n_seq

        def constr(q, x_obs_seq, partition=0):  # :473-519
            u, v_0, v_seq, n_seq = _split_q(q)
            z = generate_z(u)
            v_subseqs, n_subseqs, w_inits, y_bars = partition_into_subseqs(
                v_seq, v_0, n_seq, x_obs_seq, partition
            )
            partition_size = len(v_subseqs)
            gen_funcs = [
                (
                    lambda z, w, v, sn, i, f: vmap(
                        lambda w_, v_, sn_: generate_y_bar(z, w_, v_, sn_, i, f),
                        in_dims=(0, 0, 0 if noisy_observations else None),
                    )(w, v, sn)
                )
                if is_batched
                else generate_y_bar
                for is_batched in subseqs_are_batched[partition]
            ]
            if noisy_observations:
                σ = generate_σ(u)
                σ_n_subseqs = [σ * n_subseq for n_subseq in n_subseqs]
            else:
                σ_n_subseqs = (None,) * partition_size
            return torch.cat(
                [
                    (
                        gen_funcs[b](
                            z, w_inits[b], v_subseqs[b], σ_n_subseqs[b], b == 0, b == partition_size - 1
                        )
                        - y_bars[b]
                    ).flatten()
                    for b in range(partition_size)
                ]
            )

        def jacob_constr_blocks(q, x_obs_seq, partition=0):  # :521-624
            def g_y_bar(u, v, n, w_0, initial_subseq, final_subseq):
                z = generate_z(u)
                if noisy_observations:
                    σ = generate_σ(u)
                    σ_n = σ * n
                else:
                    σ_n = None
                if initial_subseq:
                    w_0, v = split(v, (dim_v_0,))
                v_seq = v.reshape((-1, dim_v))
                return generate_y_bar(z, w_0, v_seq, σ_n, initial_subseq, final_subseq)

            u, v_0, v_seq, n_seq = _split_q(q)
            v_subseqs, n_subseqs, w_inits, _ = partition_into_subseqs(
                v_seq, v_0, n_seq, x_obs_seq, partition
            )
            partition_size = len(v_subseqs)
            v_bars = [torch.cat([v_0, v_subseqs[0].flatten()])]
            for b in range(1, partition_size):
                v_bars.append(
                    v_subseqs[b].reshape((v_subseqs[b].shape[0], -1))
                    if subseqs_are_batched[partition][b]
                    else v_subseqs[b].flatten()
                )

            def jacob_func(b, is_batched):
                i, f = b == 0, b == partition_size - 1

                def single(u_, v_, n_, w_):
                    return jacrev(lambda uu, vv: g_y_bar(uu, vv, n_, w_, i, f), argnums=(0, 1))(
                        u_, v_
                    )

                if is_batched:
                    return lambda u_, v_, n_, w_: vmap(
                        lambda vv, nn, ww: single(u_, vv, nn, ww),
                        in_dims=(0, 0 if noisy_observations else None, 0),
                    )(v_, n_, w_)
                return single

            if noisy_observations:
                σ = generate_σ(u)
                dc_dn_blocks = tuple(
                    (σ * torch.ones_like(n_subseqs[b])).reshape(
                        (n_subseqs[b].shape[0], -1) if is_batched else (-1,)
                    )
                    for b, is_batched in enumerate(subseqs_are_batched[partition])
                )
            else:
                dc_dn_blocks = (None,) * partition_size
            dc_du_blocks, dc_dv_blocks = zip(
                *(
                    jacob_func(b, subseqs_are_batched[partition][b])(
                        u, v_bars[b], n_subseqs[b], w_inits[b]
                    )
                    for b in range(partition_size)
                )
            )
            return tuple(dc_du_blocks), tuple(dc_dv_blocks), dc_dn_blocks

        def get_M_0_matrix():  # :794-798
            return torch.eye(dim_u)

        def compute_D_blocks(dc_dv_l_blocks, dc_dn_l_blocks, dc_dv_r_blocks, dc_dn_r_blocks):
            # :765-792
            D_blocks = [
                torch.einsum("...ij,...kj->...ik", l, r)
                for l, r in zip(dc_dv_l_blocks, dc_dv_r_blocks)
            ]
            if noisy_observations:
                for b, (D_block, l, r) in enumerate(
                    zip(D_blocks[:-1], dc_dn_l_blocks[:-1], dc_dn_r_blocks[:-1])
                ):
                    add = torch.cat(
                        [l * r, torch.zeros(tuple(D_block.shape[-3:-2]) + (dim_x,))], dim=-1
                    )
                    D_blocks[b] = D_block + torch.diag_embed(add)
                D_blocks[-1] = D_blocks[-1] + torch.diag_embed(
                    dc_dn_l_blocks[-1] * dc_dn_r_blocks[-1]
                )
            return D_blocks

        def _cho_solve(chol, rhs):
            if rhs.ndim == chol.ndim - 1:
                return torch.cholesky_solve(rhs.unsqueeze(-1), chol).squeeze(-1)
            return torch.cholesky_solve(rhs, chol)

        def chol_gram_blocks(dc_du_blocks, dc_dv_blocks, dc_dn_blocks):  # :626-687
            M_0 = get_M_0_matrix()
            D_blocks = compute_D_blocks(dc_dv_blocks, dc_dn_blocks, dc_dv_blocks, dc_dn_blocks)
            chol_D_blocks = tuple(torch.linalg.cholesky(D_block) for D_block in D_blocks)
            D_inv_dc_du_blocks = tuple(
                _cho_solve(chol_D_block, dc_du_block)
                for chol_D_block, dc_du_block in zip(chol_D_blocks, dc_du_blocks)
            )
            chol_C = torch.linalg.cholesky(
                M_0
                + sum(
                    dc_du_block.T @ D_inv_dc_du_block
                    if dc_du_block.ndim == 2
                    else torch.einsum("ijk,ijl->kl", dc_du_block, D_inv_dc_du_block)
                    for dc_du_block, D_inv_dc_du_block in zip(dc_du_blocks, D_inv_dc_du_blocks)
                )
            )
            return chol_C, chol_D_blocks

        def _lu_solve(lu_piv, rhs):
            lu, piv = lu_piv
            if rhs.ndim == lu.ndim - 1:
                return torch.linalg.lu_solve(lu, piv, rhs.unsqueeze(-1)).squeeze(-1)
            return torch.linalg.lu_solve(lu, piv, rhs)

        def lu_jacob_product_blocks(
            dc_du_l_blocks, dc_dv_l_blocks, dc_dn_l_blocks, dc_du_r_blocks, dc_dv_r_blocks, dc_dn_r_blocks
        ):  # :689-763
            M_0 = get_M_0_matrix()
            D_blocks = compute_D_blocks(
                dc_dv_l_blocks, dc_dn_l_blocks, dc_dv_r_blocks, dc_dn_r_blocks
            )
            lu_and_piv_D_blocks = tuple(tuple(torch.linalg.lu_factor(D)) for D in D_blocks)
            D_inv_dc_du_l_blocks = tuple(
                _lu_solve(lu_piv, dc_du_l_block)
                for lu_piv, dc_du_l_block in zip(lu_and_piv_D_blocks, dc_du_l_blocks)
            )
            lu_and_piv_C = tuple(
                torch.linalg.lu_factor(
                    M_0
                    + sum(
                        dc_du_r_block.T @ D_inv_dc_du_l_block
                        if dc_du_r_block.ndim == 2
                        else torch.einsum("ijk,ijl->kl", dc_du_r_block, D_inv_dc_du_l_block)
                        for dc_du_r_block, D_inv_dc_du_l_block in zip(
                            dc_du_r_blocks, D_inv_dc_du_l_blocks
                        )
                    )
                )
            )
            return lu_and_piv_C, lu_and_piv_D_blocks

        def log_det_sqrt_gram_from_chol(chol_C, chol_D_blocks):  # :800-810
            return (
                sum(
                    torch.log(torch.abs(torch.diagonal(chol_D_block, 0, -2, -1))).sum()
                    for chol_D_block in chol_D_blocks
                )
                + torch.log(torch.abs(torch.diagonal(chol_C))).sum()
                - log_det_sqrt_metric_0
            )

        def log_det_sqrt_gram(q, x_obs_seq, partition=0):  # :812-820
            jac_blocks = jacob_constr_blocks(q, x_obs_seq, partition)
            chol_blocks = chol_gram_blocks(*jac_blocks)
            return log_det_sqrt_gram_from_chol(*chol_blocks), (jac_blocks, chol_blocks)

        def lmult_by_jacob_constr(dc_du_blocks, dc_dv_blocks, dc_dn_blocks, vct):  # :822-877
            if noisy_observations:
                vct_u, vct_v, vct_n = split(
                    vct, (dim_u, dim_v_0 + num_obs * num_steps_per_obs * dim_v)
                )
            else:
                vct_u, vct_v = split(vct, (dim_u,))
            vct_v_parts = split_and_reshape(
                vct_v,
                [
                    tuple(dc_dv_block.shape[0:3:2]) if dc_dv_block.ndim == 3 else tuple(dc_dv_block.shape[1:2])
                    for dc_dv_block in dc_dv_blocks
                ],
            )
            dc_du = torch.vstack(
                [
                    dc_du_block.reshape((-1, dim_u)) if dc_du_block.ndim == 3 else dc_du_block
                    for dc_du_block in dc_du_blocks
                ]
            )
            jacob_vct = dc_du @ vct_u + torch.cat(
                [
                    torch.einsum("ijk,ik->ij", dc_dv_block, vct_v_part).flatten()
                    if dc_dv_block.ndim == 3
                    else dc_dv_block @ vct_v_part
                    for dc_dv_block, vct_v_part in zip(dc_dv_blocks, vct_v_parts)
                ]
            )
            if noisy_observations:
                vct_n_parts = split_and_reshape(
                    vct_n, [tuple(dc_dn_block.shape) for dc_dn_block in dc_dn_blocks]
                )
                jacob_vct = jacob_vct + torch.cat(
                    [
                        torch.cat(
                            [dc_dn_block * vct_n_part, torch.zeros((dc_dn_block.shape[0], dim_x))],
                            dim=1,
                        ).flatten()
                        if dc_dn_block.ndim == 2
                        else torch.cat([dc_dn_block * vct_n_part, torch.zeros(dim_x)])
                        for dc_dn_block, vct_n_part in zip(dc_dn_blocks[:-1], vct_n_parts[:-1])
                    ]
                    + [dc_dn_blocks[-1] * vct_n_parts[-1]]
                )
            return jacob_vct

        def rmult_by_jacob_constr(dc_du_blocks, dc_dv_blocks, dc_dn_blocks, vct):  # :879-913
            vct_parts = split_and_reshape(
                vct, [tuple(dc_du_block.shape[:-1]) for dc_du_block in dc_du_blocks]
            )
            return torch.cat(
                [
                    sum(
                        torch.einsum("ij,ijk->k", vct_part, dc_du_block)
                        if vct_part.ndim == 2
                        else vct_part @ dc_du_block
                        for vct_part, dc_du_block in zip(vct_parts, dc_du_blocks)
                    )
                ]
                + [
                    torch.einsum("ij,ijk->ik", vct_part, dc_dv_block).flatten()
                    if vct_part.ndim == 2
                    else vct_part @ dc_dv_block
                    for vct_part, dc_dv_block in zip(vct_parts, dc_dv_blocks)
                ]
                + (
                    [
                        (vct_part[:, :-dim_x] * dc_dn_block).flatten()
                        if vct_part.ndim == 2
                        else vct_part[:-dim_x] * dc_dn_block
                        for vct_part, dc_dn_block in zip(vct_parts[:-1], dc_dn_blocks[:-1])
                    ]
                    + [vct_parts[-1] * dc_dn_blocks[-1]]
                    if noisy_observations
                    else []
                )
            )

        def lmult_by_inv_gram(dc_du_blocks, dc_dv_blocks, dc_dn_blocks, chol_C, chol_D_blocks, vct):
            # :915-942
            vct_parts = split_and_reshape(
                vct, [tuple(dc_du_block.shape[:-1]) for dc_du_block in dc_du_blocks]
            )
            D_inv_vct_blocks = [
                _cho_solve(chol_D_block, vct_part)
                for chol_D_block, vct_part in zip(chol_D_blocks, vct_parts)
            ]
            dc_du_T_D_inv_vct = sum(
                torch.einsum("...jk,...j->k", dc_du_block, D_inv_vct_block)
                for dc_du_block, D_inv_vct_block in zip(dc_du_blocks, D_inv_vct_blocks)
            )
            C_inv_dc_du_T_D_inv_vct = _cho_solve(chol_C, dc_du_T_D_inv_vct)
            return torch.cat(
                [
                    _cho_solve(
                        chol_D_block, vct_part - dc_du_block @ C_inv_dc_du_T_D_inv_vct
                    ).flatten()
                    for chol_D_block, vct_part, dc_du_block in zip(
                        chol_D_blocks, vct_parts, dc_du_blocks
                    )
                ]
            )

        def lmult_by_inv_jacob_product(
            dc_du_l_blocks,
            dc_dv_l_blocks,
            dc_dn_l_blocks,
            dc_du_r_blocks,
            dc_dv_r_blocks,
            dc_dn_r_blocks,
            lu_and_piv_C,
            lu_and_piv_D_blocks,
            vct,
        ):  # :944-981
            vct_parts = split_and_reshape(
                vct, [tuple(dc_du_l_block.shape[:-1]) for dc_du_l_block in dc_du_l_blocks]
            )
            D_inv_vct_blocks = [
                _lu_solve(lu_piv, vct_part)
                for lu_piv, vct_part in zip(lu_and_piv_D_blocks, vct_parts)
            ]
            dc_du_r_T_D_inv_vct = sum(
                torch.einsum("...jk,...j->k", dc_du_r_block, D_inv_vct_block)
                for dc_du_r_block, D_inv_vct_block in zip(dc_du_r_blocks, D_inv_vct_blocks)
            )
            C_inv_dc_du_r_T_D_inv_vct = _lu_solve(lu_and_piv_C, dc_du_r_T_D_inv_vct)
            return torch.cat(
                [
                    _lu_solve(
                        lu_piv, vct_part - dc_du_l_block @ C_inv_dc_du_r_T_D_inv_vct
                    ).flatten()
                    for lu_piv, vct_part, dc_du_l_block in zip(
                        lu_and_piv_D_blocks, vct_parts, dc_du_l_blocks
                    )
                ]
            )

        def normal_space_component(vct, jacob_constr_blocks, chol_gram_blocks):  # :983-993
            return rmult_by_jacob_constr(
                *jacob_constr_blocks,
                lmult_by_inv_gram(
                    *jacob_constr_blocks,
                    *chol_gram_blocks,
                    lmult_by_jacob_constr(*jacob_constr_blocks, vct),
                ),
            )

        def norm(x):  # :995-997
            return torch.max(torch.abs(x))

        def _loop(body_func, q, constraint_tol, position_tol, divergence_tol, max_iters, dt):
            # lax.while_loop with cond_func :1047-1055 / :1119-1127
            val = (q, torch.zeros_like(q), 0, float("inf"), -1.0)
            while True:
                _, _, i, norm_delta_q, error = val
                diverged = error > divergence_tol or error != error
                converged = error < constraint_tol and norm_delta_q < position_tol
                if i >= max_iters or diverged or converged:
                    break
                val = body_func(val)
            q, mu, i, norm_delta_q, error = val
            if use_gaussian_splitting:
                return q, mu / onp.sin(dt), i, norm_delta_q, error
            else:
                return q, mu / dt, i, norm_delta_q, error

        def quasi_newton_projection(
            q,
            x_obs_seq,
            partition,
            jacob_constr_blocks_prev,
            chol_gram_blocks_prev,
            dt,
            constraint_tol,
            position_tol,
            divergence_tol,
            max_iters,
        ):  # :1009-1063
            def body_func(val):
                q, mu, i, _, _ = val
                c = constr(q, x_obs_seq, partition)
                error = float(norm(c))
                delta_mu = rmult_by_jacob_constr(
                    *jacob_constr_blocks_prev,
                    lmult_by_inv_gram(*jacob_constr_blocks_prev, *chol_gram_blocks_prev, c),
                )
                delta_q = delta_mu
                mu = mu + delta_mu
                q = q - delta_q
                i += 1
                return q, mu, i, float(norm(delta_q)), error

            return _loop(body_func, q, constraint_tol, position_tol, divergence_tol, max_iters, dt)

        def newton_projection(
            q,
            x_obs_seq,
            partition,
            jacob_constr_blocks_prev,
            dt,
            constraint_tol,
            position_tol,
            divergence_tol,
            max_iters,
        ):  # :1075-1135
            def body_func(val):
                q, mu, i, _, _ = val
                c = constr(q, x_obs_seq, partition)
                jacob_constr_blocks_curr = jacob_constr_blocks(q, x_obs_seq, partition)
                lu_and_piv_jacob_product_blocks = lu_jacob_product_blocks(
                    *jacob_constr_blocks_curr, *jacob_constr_blocks_prev
                )
                error = float(norm(c))
                delta_mu = rmult_by_jacob_constr(
                    *jacob_constr_blocks_prev,
                    lmult_by_inv_jacob_product(
                        *jacob_constr_blocks_curr,
                        *jacob_constr_blocks_prev,
                        *lu_and_piv_jacob_product_blocks,
                        c,
                    ),
                )
                delta_q = delta_mu
                mu = mu + delta_mu
                q = q - delta_q
                i += 1
                return q, mu, i, float(norm(delta_q)), error

            return _loop(body_func, q, constraint_tol, position_tol, divergence_tol, max_iters, dt)

        self._generate_x_obs_seq = generate_x_obs_seq
        self._constr = constr
        self._jacob_constr_blocks = jacob_constr_blocks
        self._chol_gram_blocks = chol_gram_blocks
        self._lu_jacob_product_blocks = lu_jacob_product_blocks
        self._log_det_sqrt_gram_from_chol = log_det_sqrt_gram_from_chol
        self._log_det_sqrt_gram = log_det_sqrt_gram

        def _grad_log_det_sqrt_gram(q, x_obs_seq, partition=0):  # :1143-1146
            def val_and_aux(q_):
                val, ((du, dv, dn), chol) = log_det_sqrt_gram(q_, x_obs_seq, partition)
                # torch.func aux outputs must be tensors: swap `None` (noiseless dc_dn) for ()
                dn_ = () if dn[0] is None else dn
                return val, (val, du, dv, dn_, chol)

            g, (val, du, dv, dn, chol) = grad(val_and_aux, has_aux=True)(q)
            if len(dn) == 0:
                dn = (None,) * len(du)
            return (val, ((du, dv, dn), chol)), g

        self._grad_log_det_sqrt_gram = _grad_log_det_sqrt_gram
        self._normal_space_component = normal_space_component
        self._lmult_by_jacob_constr = lmult_by_jacob_constr
        self._rmult_by_jacob_constr = rmult_by_jacob_constr
        self._lmult_by_inv_gram = lmult_by_inv_gram
        self._quasi_newton_projection = quasi_newton_projection
        self._newton_projection = newton_projection

    # ------------------------------------------------------------------------------------
    # System-level methods on a plain `State` (no Mici cache): mici_extensions.py:1151-1259
    # ------------------------------------------------------------------------------------

    def point(self, q, x_obs_seq, partition):
        """One `grad_log_det_sqrt_gram` evaluation = everything Mici caches at a position
        (:1173-1184): gradient, log-det, Jacobian blocks, Cholesky blocks."""
        (val, (jac, chol)), g = self._grad_log_det_sqrt_gram(_t(q), _t(x_obs_seq), partition)
        jac = tuple(tuple(None if a is None else a.detach() for a in part) for part in jac)
        chol = (chol[0].detach(), tuple(a.detach() for a in chol[1]))
        return {"grad_ld": g.detach(), "ld": float(val), "jac": jac, "chol": chol}

    def dh1_dpos(self, q, pt):  # :1192-1196
        if self.use_gaussian_splitting:
            return pt["grad_ld"]
        return q + pt["grad_ld"]

    def h(self, q, p, pt):  # :1186-1202
        return 0.5 * float(q @ q) + pt["ld"] + 0.5 * float(p @ p)

    def h2_flow(self, q, p, dt):  # :1222-1231
        if self.use_gaussian_splitting:
            s, c = onp.sin(dt), onp.cos(dt)
            return c * q + s * p, c * p - s * q
        return q + dt * p, p

    def project_onto_cotangent_space(self, p, pt):  # :1243-1254
        return p - self._normal_space_component(p, pt["jac"], pt["chol"])


def leapfrog_step(
    system,
    q,
    p,
    x_obs_seq,
    partition,
    dt,
    pt=None,
    solver="quasi_newton",
    constraint_tol=1e-9,
    position_tol=1e-8,
    divergence_tol=1e10,
    max_iters=50,
    reverse_check_tol=2e-8,
    n_inner_step=1,
):
    """One ConstrainedLeapfrogIntegrator.step (Mici 0.1.10, SURVEY.md 3.3: _step_a, n_inner_step x _step_b, _step_a)
    using the reference's projection solvers (mici_extensions.py:1323-1476).

    Returns (q, p, pt_new, info).  Raises ConvergenceError / NonReversibleStepError like Mici."""
    q, p, x_obs_seq = _t(q), _t(p), _t(x_obs_seq)
    if pt is None:
        pt = system.point(q, x_obs_seq, partition)
    info = {}

    def project(q_, q_prev, pt_prev, dt_):
        if solver == "quasi_newton":
            q_new, mu, i, ndq, err = system._quasi_newton_projection(
                q_, x_obs_seq, partition, pt_prev["jac"], pt_prev["chol"], dt_,
                constraint_tol, position_tol, divergence_tol, max_iters,
            )
        else:
            q_new, mu, i, ndq, err = system._newton_projection(
                q_, x_obs_seq, partition, pt_prev["jac"], dt_,
                constraint_tol, position_tol, divergence_tol, max_iters,
            )
        if err < constraint_tol and ndq < position_tol:
            return q_new, mu, i
        elif err > divergence_tol or err != err:
            raise ConvergenceError(f"diverged on iteration {i}: |c|={err:.1e} |dq|={ndq}")
        else:
            raise ConvergenceError(f"did not converge: |c|={err:.1e} |dq|={ndq}")

    # A(dt/2)
    p = p - 0.5 * dt * system.dh1_dpos(q, pt)
    p = system.project_onto_cotangent_space(p, pt)
    # B(dt): n_inner_step inner steps, each with its own projection and reverse check
    dt_i = dt / n_inner_step
    n_fwd_all, n_back_all, rev = [], [], 0.0
    for _ in range(n_inner_step):
        q_prev, pt_prev = q, pt
        q_, p_ = system.h2_flow(q, p, dt_i)
        q_new, mu, n_fwd = project(q_, q_prev, pt_prev, dt_i)
        cos_or_one = onp.cos(dt_i) if system.use_gaussian_splitting else 1.0
        p = p_ - cos_or_one * mu  # state.mom -= dh2_flow_mom_dmom @ mu  (:1391, :1233-1238)
        pt_new = system.point(q_new, x_obs_seq, partition)
        p = system.project_onto_cotangent_space(p, pt_new)
        q_b, _ = system.h2_flow(q_new, p, -dt_i)
        q_back, _, n_back = project(q_b, q_new, pt_new, -dt_i)
        rev = float(torch.max(torch.abs(q_back - q_prev)))
        n_fwd_all.append(n_fwd)
        n_back_all.append(n_back)
        info.update(n_fwd=n_fwd, n_back=n_back, rev_diff=rev, n_fwd_inner=n_fwd_all, n_back_inner=n_back_all)
        if rev > reverse_check_tol:
            raise NonReversibleStepError(f"reverse error {rev:.2e}")
        q, pt = q_new, pt_new
    # A(dt/2)
    p = p - 0.5 * dt * system.dh1_dpos(q_new, pt_new)
    p = system.project_onto_cotangent_space(p, pt_new)
    return q_new, p, pt_new, info


def find_initial_state_by_linear_interpolation(system, rng, generate_x_obs_seq_init, u=None, v_0=None):
    """mici_extensions.py:1479-1547 (without the momentum draw, returned separately)."""
    md = system.model_dict
    S, dim_v = md["num_steps_per_obs"], md["dim_v"]

    def mean_and_sqrt_covar_step_diff(z, x, δ):  # :1495-1501
        v = torch.zeros(dim_v)

        def step_diff_func(v):
            return md["forward_func"](z, x, v, δ) - x

        return step_diff_func(v), torch.func.jacfwd(step_diff_func)(v)

    def solve_for_v_seq(x_obs_seq, x_0, z):  # :1503-1526
        def solve_inner(x, Δx):
            mean_diff, sqrt_covar_diff = mean_and_sqrt_covar_step_diff(z, x, md["δ"])
            # np.linalg.lstsq on a square full-rank system == solve
            return torch.linalg.solve(sqrt_covar_diff, Δx - mean_diff)

        def solve_outer(x_0_, x_1_):
            Δx = (x_1_ - x_0_) / S
            x_seq = x_0_[None] + torch.arange(S, dtype=x_0_.dtype)[:, None] * Δx[None]
            return vmap(solve_inner, in_dims=(0, None))(x_seq, Δx)

        x_0_seq = torch.cat((x_0[None], x_obs_seq[:-1]))
        x_1_seq = x_obs_seq
        return vmap(solve_outer)(x_0_seq, x_1_seq).reshape((-1, dim_v))

    u = rng.standard_normal(md["dim_u"]) if u is None else u
    u = _t(u)
    z = md["generate_z"](u)
    v_0 = rng.standard_normal(md["dim_v_0"]) if v_0 is None else v_0
    v_0 = _t(v_0)
    x_0 = md["generate_x_0"](z, v_0)
    x_obs_seq = _t(generate_x_obs_seq_init(rng))
    v_seq = solve_for_v_seq(x_obs_seq, x_0, z)
    if md["generate_σ"] is not None:
        n = torch.zeros(md["dim_y"] * md["num_obs"])
        q = torch.cat([u, v_0, v_seq.flatten(), n])
    else:
        q = torch.cat([u, v_0, v_seq.flatten()])
    return q, x_obs_seq


# ---------------------------------------------------------------------------------------------------
# standard-HMC target and the Adam initialiser for noisy systems
# ---------------------------------------------------------------------------------------------------
def conditioned_diffusion_neg_log_dens_and_grad(
    obs_interval, num_steps_per_obs, y_seq, dim_u, dim_v_0, dim_v, forward_func, generate_x_0, generate_z,
    generate_σ, obs_func, use_gaussian_splitting=False,
):
    """mici_extensions.py:82-205 -- (neg_log_dens, value_and_grad) on torch tensors; the gradient is reverse-mode
    automatic differentiation through the time loop, as jax.value_and_grad through lax.scan in the reference."""
    y_seq = _t(y_seq)
    num_obs, dim_y = y_seq.shape
    δ = obs_interval / num_steps_per_obs
    if isinstance(generate_σ, Number):
        σ_fixed = generate_σ

        def generate_σ(u):  # noqa: F811  (:146-150)
            return torch.as_tensor(σ_fixed, dtype=torch.float64)

    def _neg_log_dens(q):  # :152-171
        num_step = num_steps_per_obs * num_obs
        u, v_0, v_seq_flat = split(q, (dim_u, dim_v_0, dim_v * num_step))
        z = generate_z(u)
        σ = generate_σ(u)
        x = generate_x_0(z, v_0)
        v_seq = v_seq_flat.reshape((num_step, dim_v))
        xs = []
        for t in range(num_step):
            x = forward_func(z, x, v_seq[t], δ)
            xs.append(x)
        x_seq = torch.stack(xs)
        y_seq_mean = obs_func(x_seq[num_steps_per_obs - 1 :: num_steps_per_obs])
        return (
            0.5 * torch.sum(((y_seq - y_seq_mean) / σ) ** 2)
            + num_obs * dim_y * torch.log(σ)
            + (0 if use_gaussian_splitting else 0.5 * torch.sum(q ** 2))
        )

    def _value_and_grad(q):  # :173
        q = _t(q).clone().requires_grad_(True)
        val = _neg_log_dens(q)
        (g,) = torch.autograd.grad(val, q)
        return val.detach(), g

    return (lambda q: _neg_log_dens(_t(q))), _value_and_grad


def find_initial_state_by_gradient_descent_noisy_system(
    model_dict, rng, adam_step_size=2e-2, max_iters=1000, max_init_tries=100, max_num_tries=10, threshold=1.0,
    slow_progress_ratio=0.8, check_iter=100,
):
    """mici_extensions.py:1679-1801 restated: returns (u_v, residuals, tries, iterations) of the accepted point.
    Adam follows jax.experimental.optimizers.adam (b1 = 0.9, b2 = 0.999, eps = 1e-8; :1733)."""
    md = model_dict
    num_step = md["num_steps_per_obs"] * md["num_obs"]
    dim_u_v = md["dim_u"] + md["dim_v_0"] + num_step * md["dim_v"]
    y_seq = _t(md["y_seq"])
    gen_σ = md["generate_σ"]
    if isinstance(gen_σ, Number):
        σ_fixed = gen_σ
        gen_σ = lambda u: torch.as_tensor(σ_fixed, dtype=torch.float64)  # noqa: E731

    def init_objective(u_v):  # :1701-1731
        u, v_0, v_flat = split(u_v, (md["dim_u"], md["dim_v_0"], num_step * md["dim_v"]))
        v_seq = v_flat.reshape((num_step, md["dim_v"]))
        z = md["generate_z"](u)
        x = md["generate_x_0"](z, v_0)
        σ = gen_σ(u)
        xs = []
        for t in range(num_step):
            x = md["forward_func"](z, x, v_seq[t], md["δ"])
            xs.append(x)
        x_seq = torch.stack(xs)
        obs_slc = slice(md["num_steps_per_obs"] - 1, None, md["num_steps_per_obs"])
        residuals = (y_seq - md["obs_func"](x_seq[obs_slc])) / σ
        return 0.5 * torch.sum(residuals ** 2) + md["num_obs"] * torch.log(σ) + 0.5 * torch.sum(u_v ** 2), residuals

    def grad_init_objective(u_v):
        u_v = u_v.clone().requires_grad_(True)
        val, res = init_objective(u_v)
        (g,) = torch.autograd.grad(val, u_v)
        return g, res.detach()

    b1, b2, eps = 0.9, 0.999, 1e-8
    for t in range(max_num_tries):
        found_valid_init_state, init_tries = False, 0
        while not found_valid_init_state and init_tries < max_init_tries:
            u_v = _t(rng.standard_normal(dim_u_v))
            with torch.no_grad():
                _, residuals = init_objective(u_v)
            if bool(torch.isfinite(residuals).all()):
                found_valid_init_state = True
            init_tries += 1
        if init_tries == max_init_tries:
            raise RuntimeError(f"Did not find valid initial state in {init_tries} tries.")
        m, v = torch.zeros_like(u_v), torch.zeros_like(u_v)
        prev_mean_residual_sq = float(torch.mean(residuals ** 2))
        for i in range(max_iters):
            g, residuals = grad_init_objective(u_v)
            m_n = (1 - b1) * g + b1 * m
            v_n = (1 - b2) * g ** 2 + b2 * v
            u_v_next = u_v - adam_step_size * (m_n / (1 - b1 ** (i + 1))) / (torch.sqrt(v_n / (1 - b2 ** (i + 1))) + eps)
            mean_residuals_sq = float(torch.mean(residuals ** 2))
            if not onp.isfinite(mean_residuals_sq):
                break
            if mean_residuals_sq < threshold:
                return u_v.numpy(), residuals.numpy().flatten(), t + 1, i
            u_v, m, v = u_v_next, m_n, v_n
            if i % check_iter == 0:
                if i > 0 and i < max_iters // 2 and (mean_residuals_sq / prev_mean_residual_sq) > slow_progress_ratio:
                    break
                prev_mean_residual_sq = mean_residuals_sq
    raise RuntimeError(f"Did not find valid state in {max_num_tries} tries.")
