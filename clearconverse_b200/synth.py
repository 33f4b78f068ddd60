"""Synthetic 8 kHz two-speaker mixtures and the 1-hour-meeting overlap-segment timeline.

There is no network for datasets, so the benchmark and the parity tests feed seeded synthetic
"speech-like" mixtures: each source is a 7-harmonic tone with a random
pitch under a syllabic on/off envelope plus a little noise; the mixture is peak-normalised to
1.0 the way the reference normalises a loaded file (/root/reference/back/api.py:834).
"""
from __future__ import annotations

import math

import torch

SAMPLE_RATE = 8000


def synth_sources(n_samples: int, seed: int, sample_rate: int = SAMPLE_RATE) -> torch.Tensor:
    """Two synthetic sources [2, n_samples] fp32 (not yet normalised)."""
    g = torch.Generator().manual_seed(int(seed))
    t = torch.arange(n_samples, dtype=torch.float64) / sample_rate
    out = []
    for _ in range(2):
        f0 = 90.0 + 160.0 * torch.rand((), generator=g, dtype=torch.float64)
        phases = 2 * math.pi * torch.rand(7, generator=g, dtype=torch.float64)
        s = torch.zeros(n_samples, dtype=torch.float64)
        for k in range(1, 8):
            s += torch.sin(2 * math.pi * f0 * k * t + phases[k - 1]) / k
        rate = 2.0 + 3.0 * torch.rand((), generator=g, dtype=torch.float64)
        ph = torch.rand((), generator=g, dtype=torch.float64)
        env = ((torch.floor(t * rate * 2 + ph * 2) % 2) == 0).to(torch.float64) * 0.8 + 0.2
        s = s * env + 0.005 * torch.randn(n_samples, generator=g, dtype=torch.float64)
        out.append(s)
    return torch.stack(out).to(torch.float32)


def synth_mixture(n_samples: int, seed: int, sample_rate: int = SAMPLE_RATE):
    """(mix [n_samples], sources [2, n_samples]) with max|mix| == 1."""
    src = synth_sources(n_samples, seed, sample_rate)
    mix = src.sum(0)
    peak = mix.abs().max().clamp_min(1e-8)
    return mix / peak, src / peak


def synth_batch(batch: int, n_samples: int, seed: int = 0) -> torch.Tensor:
    """[batch, n_samples] fp32 mixtures; item i uses seed 1000*seed + i."""
    return torch.stack([synth_mixture(n_samples, 1000 * seed + i)[0] for i in range(batch)])


def meeting_overlap_segments(total_overlap_s: float = 720.0, seed: int = 4, lo: float = 0.5, hi: float = 30.0,
                             sample_rate: int = SAMPLE_RATE) -> list[int]:
    """Sample counts of the overlap sub-segments of a synthetic 1-hour meeting with ~20 % overlap
    (BASELINE.json config 4): durations log-uniform on [lo, hi] s drawn until they sum to
    ``total_overlap_s``.  The reference's regions are >= 0.3-0.5 s (api.py:116-117, 1022)."""
    g = torch.Generator().manual_seed(seed)
    out, acc = [], 0.0
    while acc < total_overlap_s:
        u = torch.rand((), generator=g).item()
        d = math.exp(math.log(lo) + u * (math.log(hi) - math.log(lo)))
        d = min(d, total_overlap_s - acc) if total_overlap_s - acc >= lo else d
        out.append(max(int(round(d * sample_rate)), 16))
        acc += d
    return out
