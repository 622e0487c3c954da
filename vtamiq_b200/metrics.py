"""Rank / linear correlation of predicted vs. subjective scores, computed where the scores live.

SURVEY.md §8f #4: the reference moves every batch of scores to the host and calls scipy
(``utils/misc/correlations.py:21-51`` from ``train.py:403-409,:548,:567,:637``).  Here SROCC, KROCC, PLCC and RMSE
are tensor expressions that run on the device holding the scores (no host sync until the caller reads the numbers);
only the optional logistic fit of ``PLCC`` / ``RMSE`` (``correlations.py:35-42``, ``FitFunction`` fit 1) is a host
least-squares problem and stays on scipy.

Definitions follow scipy's (and therefore the reference's): average ranks for ties, Kendall tau-b.
"""
from __future__ import annotations

import torch

SROCC, KROCC, PLCC, RMSE, PLCC_NOFIT, RMSE_NOFIT = "SROCC", "KROCC", "PLCC", "RMSE", "PLCC_NOFIT", "RMSE_NOFIT"
_FIT_EPS = 1e-6   # CORRELATIONS_EPS, correlations.py:11


def normalize_array(a: torch.Tensor) -> torch.Tensor:
    """min-max to [0, 1] unless the range is degenerate (utils/image_processing/image_tools.py:17-21)."""
    b = a - a.min()
    top = b.max()
    return torch.where(top.abs() > 1e-6, b / top.clamp_min(1e-30), b)


def average_ranks(x: torch.Tensor) -> torch.Tensor:
    """1-based ranks with ties sharing the mean of their positions (scipy.stats.rankdata, method='average')."""
    n = x.numel()
    order = torch.argsort(x, stable=True)
    xs = x[order]
    start = torch.ones(n, dtype=torch.bool, device=x.device)
    start[1:] = xs[1:] != xs[:-1]
    group = torch.cumsum(start, 0) - 1                               # tie-group id of each sorted element
    pos = torch.arange(1, n + 1, device=x.device, dtype=torch.float64)
    n_groups = int(group[-1].item()) + 1 if n else 0
    total = torch.zeros(n_groups, dtype=torch.float64, device=x.device).index_add_(0, group, pos)
    count = torch.zeros(n_groups, dtype=torch.float64, device=x.device).index_add_(0, group, torch.ones_like(pos))
    ranks = torch.empty(n, dtype=torch.float64, device=x.device)
    ranks[order] = (total / count)[group]
    return ranks


def pearson(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a, b = a.double() - a.double().mean(), b.double() - b.double().mean()
    return (a * b).sum() / torch.sqrt((a * a).sum() * (b * b).sum())


def spearman(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return pearson(average_ranks(a.reshape(-1)), average_ranks(b.reshape(-1)))


def kendall(a: torch.Tensor, b: torch.Tensor, block: int = 2048) -> torch.Tensor:
    """tau-b over all pairs, evaluated in (block x n) slabs so that n^2 never has to fit at once."""
    a, b = a.reshape(-1).double(), b.reshape(-1).double()
    n = a.numel()
    conc = torch.zeros((), dtype=torch.float64, device=a.device)
    ties_a = torch.zeros_like(conc)
    ties_b = torch.zeros_like(conc)
    for i in range(0, n, block):
        da = torch.sign(a[i:i + block, None] - a[None, :])
        db = torch.sign(b[i:i + block, None] - b[None, :])
        conc += (da * db).sum()                 # concordant - discordant, every unordered pair counted twice
        ties_a += (da == 0).sum()
        ties_b += (db == 0).sum()
    pairs = float(n) * (n - 1)
    ta, tb = ties_a - n, ties_b - n             # drop the diagonal; still "counted twice" like conc and pairs
    return conc / torch.sqrt((pairs - ta) * (pairs - tb))


def rmse(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return torch.sqrt(((a.double() - b.double()) ** 2).mean())


def _fit_logistic(source, target):
    """target ~ p0 (1/2 - 1/(1 + exp(p1 (x - p2)))) + |p3| x + p4, L1-style residuals (correlations.py:57-131)."""
    import numpy as np
    import scipy.optimize

    def model(p, x):
        with np.errstate(over="ignore"):   # the optimiser probes huge slopes; exp -> inf -> 1/(1+inf) = 0 is fine
            return p[0] * (0.5 - 1.0 / (1.0 + np.exp(p[1] * (x - p[2]) + _FIT_EPS))) + abs(p[3]) * x + p[4]

    guess = (1.0, 1.0, float(np.median(source)), 1.0, float(np.median(target)))
    p = scipy.optimize.leastsq(lambda q, x, y: y - model(q, x), guess, args=(source, target), full_output=True)[0]
    if np.isnan(np.asarray(p)).any():
        raise OverflowError("Fitting failed: result contains NaNs.")
    return model(p, source)


def compute_correlations(a: torch.Tensor, b: torch.Tensor, normalize: bool = True, fit: bool = True) -> dict:
    """Same keys and meaning as the reference's ``compute_correlations(a, b)``: a = subjective scores, b =
    predictions.  Everything but the logistic fit runs on ``a.device``; with ``fit=False`` no host transfer happens
    and the fitted entries equal the un-fitted ones."""
    a, b = a.reshape(-1).double(), b.reshape(-1).double().to(a.device)
    if normalize:
        a, b = normalize_array(a), normalize_array(b)
    out = {SROCC: spearman(a, b), KROCC: kendall(a, b), PLCC_NOFIT: pearson(a, b), RMSE_NOFIT: rmse(a, b)}
    bf = b
    if fit:
        try:
            bf = torch.from_numpy(_fit_logistic(b.cpu().numpy(), a.cpu().numpy())).to(a.device)
        except OverflowError:
            bf = b
    out[PLCC], out[RMSE] = pearson(a, bf), rmse(a, bf)
    return {k: float(v) for k, v in out.items()}
