"""Acceptance metrics of the separation path (host-side, plain torch; used by tests, smoke() and bench.py).

BASELINE.json's north star states two tolerances: max-abs <= 1e-3 on fp32 waveforms for the fp32-tolerance modes,
and an SI-SNR delta <= 0.05 dB for the bf16 mode.  The second needs a definition; it lives here so that every
consumer uses the same one.
"""
from __future__ import annotations

import torch

MIN_REF_SI_SNR_DB = -20.0   # masked variant: pairs whose oracle SI-SNR is below this are ill-conditioned (see below)
MIN_EST_VS_EST_DB = 45.0    # weight-independent bf16 gate: error at least 45 dB below the oracle's estimate
MAX_SI_SNR_DELTA_DB = 0.05  # north star


def si_snr_db(est: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """Scale-invariant SNR in dB of est against ref along the last dim (zero-mean)."""
    est = est.double() - est.double().mean(-1, keepdim=True)
    ref = ref.double() - ref.double().mean(-1, keepdim=True)
    proj = (est * ref).sum(-1, keepdim=True) * ref / (ref.pow(2).sum(-1, keepdim=True) + 1e-20)
    noise = est - proj
    return 10 * torch.log10(proj.pow(2).sum(-1) / (noise.pow(2).sum(-1) + 1e-20))


def si_snr_delta(est: torch.Tensor, oracle_est: torch.Tensor, mix: torch.Tensor,
                 min_ref_db: float | None = MIN_REF_SI_SNR_DB) -> tuple[float, int]:
    """The bf16 acceptance metric: (delta_db, n_pairs).

    est, oracle_est: [B,T,n_spk]; mix: [B,T].  delta = max over (item, speaker) of
    |SI-SNR(est, mix) - SI-SNR(oracle_est, mix)| in dB, taken over the ``n_pairs`` pairs whose oracle SI-SNR is
    >= ``min_ref_db`` (``None``: every pair, the unmasked delta).  The mixture is the reference because there are no
    trained weights here: with random-init weights SI-SNR against the true sources sits near -38 dB.

    Why a mask exists at all: a perturbation S dB below the estimate can move an SI-SNR of R dB by up to
    20 log10(1 + 10^((|R| - S) / 20)); for R -> -inf (estimate orthogonal to the reference) any perturbation moves it
    arbitrarily, and some random weight seeds put a speaker at R = -45 ... -73 dB.  A caller that relies on the masked
    delta MUST check ``n_pairs >= 1`` -- an empty mask proves nothing (``n_pairs == 0`` returns delta = nan so that a
    bare ``<=`` comparison fails instead of passing).  On the filterbank weight set
    (``weights.filterbank_init_state_dicts``) every pair sits at R = +5 ... +10 dB and the UNMASKED delta is the gate."""
    a = si_snr_db(est.permute(0, 2, 1), mix[:, None, :])
    b = si_snr_db(oracle_est.permute(0, 2, 1), mix[:, None, :])
    ok = torch.ones_like(b, dtype=torch.bool) if min_ref_db is None else b >= min_ref_db
    n = int(ok.sum().item())
    if n == 0:
        return float("nan"), 0
    return (a - b).abs()[ok].max().item(), n


def est_vs_est_db(est: torch.Tensor, oracle_est: torch.Tensor) -> float:
    """min over (item, speaker) of SI-SNR(est, oracle_est): how far the error sits below the oracle's estimate."""
    return si_snr_db(est.permute(0, 2, 1), oracle_est.permute(0, 2, 1)).min().item()
