"""Cross-chain convergence diagnostics: rank-normalised split-R-hat and bulk effective sample size.

Replaces the ``arviz.summary`` call at the end of the reference's experiment scripts
(``scripts/utils.py:368-381``; ArviZ 0.11.2 is an un-vendored dependency).  The algorithms are the
published ones (Vehtari, Gelman, Simpson, Carpenter, Buerkner 2021): split each chain in two,
rank-normalise across all draws, R-hat from between/within variances, ESS from the FFT
autocorrelation with Geyer's initial monotone sequence truncation.

Runs on the host on the gathered traces (a few doubles per chain per iteration); with several GPUs
the traces are gathered with one all-gather (``parallel.allgather_chains``), the only collective
of a run.
"""

import numpy as np
from scipy import special, stats


def _split_chains(x):
    n = x.shape[1] // 2
    return np.concatenate([x[:, :n], x[:, -n:]], axis=0)


def _rank_normalise(x):
    r = stats.rankdata(x.reshape(-1), method="average").reshape(x.shape)
    return special.ndtri((r - 0.375) / (x.size + 0.25))


def _rhat_plain(x):
    m, n = x.shape
    chain_mean = x.mean(axis=1)
    chain_var = x.var(axis=1, ddof=1)
    between = n * chain_mean.var(ddof=1)
    within = chain_var.mean()
    return float(np.sqrt(((n - 1) / n * within + between / n) / within))


def rhat(x):
    """Rank-normalised split-R-hat of draws x[chain, iteration] (max of bulk and folded)."""
    x = np.asarray(x, dtype=np.float64)
    s = _split_chains(x)
    bulk = _rhat_plain(_rank_normalise(s))
    folded = _rhat_plain(_rank_normalise(np.abs(s - np.median(s))))
    return max(bulk, folded)


def _autocov(x):
    n = x.shape[-1]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    xc = x - x.mean(axis=-1, keepdims=True)
    f = np.fft.rfft(xc, n=m, axis=-1)
    ac = np.fft.irfft(f * np.conj(f), n=m, axis=-1)[..., :n]
    return ac / n


def _ess_plain(x):
    m, n = x.shape
    acov = _autocov(x)
    chain_mean = x.mean(axis=1)
    mean_var = acov[:, 0].mean() * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if m > 1:
        var_plus += chain_mean.var(ddof=1)
    rho = np.zeros(n)
    t = 0
    rho_even = 1.0
    rho[0] = rho_even
    rho_odd = 1.0 - (mean_var - acov[:, 1].mean()) / var_plus
    rho[1] = rho_odd
    t = 1
    while t < n - 3 and (rho_even + rho_odd) > 0.0:
        rho_even = 1.0 - (mean_var - acov[:, t + 1].mean()) / var_plus
        rho_odd = 1.0 - (mean_var - acov[:, t + 2].mean()) / var_plus
        if rho_even + rho_odd >= 0:
            rho[t + 1] = rho_even
            rho[t + 2] = rho_odd
        t += 2
    max_t = t - 2
    if rho_even > 0:
        rho[max_t + 1] = rho_even
    # Geyer's initial monotone sequence
    t = 1
    while t <= max_t - 2:
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    ess = m * n
    tau = -1.0 + 2.0 * rho[: max_t + 1].sum() + rho[max_t + 1]
    tau = max(tau, 1.0 / np.log10(ess))
    return float(ess / tau)


def ess_bulk(x):
    """Bulk effective sample size of draws x[chain, iteration]."""
    x = np.asarray(x, dtype=np.float64)
    return _ess_plain(_rank_normalise(_split_chains(x)))


def summary(traces):
    """`arviz.summary`-like table: {var: {mean, sd, ess_bulk, r_hat}} for traces[var][chain, iter]."""
    out = {}
    for k, v in traces.items():
        v = np.asarray(v, dtype=np.float64)
        out[k] = {"mean": float(v.mean()), "sd": float(v.std(ddof=1)), "ess_bulk": ess_bulk(v), "r_hat": rhat(v)}
    return out


def save_and_print_summary(output_dir, traces, summary_vars, sampling_time, step_size, call_counts=None,
                           verbose=True):
    """Output format of the reference's experiment scripts (``scripts/utils.py:368-381``): ``summary.json`` in
    the column-major layout of ``arviz.summary(...).to_dict()`` ({statistic: {variable: value}}) plus the totals the
    plotting / cost-per-ESS code reads (``load_summary_data``, ``utils.py:484-523``), and one
    ``trace_{chain}_{var}.npy`` per chain and variable (the names ``load_traces`` globs for, ``utils.py:555-556``).

    traces[var]: array [n_chains, n_iter] (scalar variables) or [n_chains, n_iter, k] (vectors, written as
    ``var[i]`` like ArviZ does)."""
    import json
    import os

    os.makedirs(output_dir, exist_ok=True)
    flat = {}
    for var in summary_vars:
        v = np.asarray(traces[var], dtype=np.float64)
        if v.ndim == 2:
            flat[var] = v
        else:
            for i in range(v.shape[2]):
                flat[f"{var}[{i}]"] = v[:, :, i]
    table = summary(flat)
    cols = ("mean", "sd", "ess_bulk", "r_hat")
    summary_dict = {c: {k: float(row[c]) for k, row in table.items()} for c in cols}
    summary_dict["mcse_mean"] = {k: float(row["sd"] / np.sqrt(max(row["ess_bulk"], 1.0))) for k, row in table.items()}
    summary_dict["total_sampling_time"] = float(sampling_time)
    summary_dict["final_integrator_step_size"] = float(step_size)
    for name, total in (call_counts or {}).items():
        summary_dict[f"total_{name}_calls"] = int(total)
    with open(os.path.join(output_dir, "summary.json"), mode="w") as f:
        json.dump(summary_dict, f, ensure_ascii=False, indent=2)
    for var, v in traces.items():
        v = np.asarray(v)
        for c in range(v.shape[0]):
            np.save(os.path.join(output_dir, f"trace_{c}_{var}.npy"), v[c])
    if verbose:
        print(f"Integrator step size = {step_size:.2g}")
        print(f"Total sampling time = {sampling_time:.0f} seconds")
        for k, row in table.items():
            print(f"{k:>10s}  mean {row['mean']:9.4f}  sd {row['sd']:8.4f}  ess_bulk {row['ess_bulk']:8.0f}  "
                  f"r_hat {row['r_hat']:.3f}")
    return summary_dict
