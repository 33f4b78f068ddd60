"""Host-side feeder of the separation path: which audio spans of a file reach ``separate_batch``.

This is the step immediately before the hot path in the reference: for every diarization
segment that touches an overlap region, ``_process_overlap_segment`` (/root/reference/back/api.py:1066-1118) first
re-segments it with a sliding window of speaker embeddings (``_resegment_overlap``, api.py:961-1050; window 0.80 s,
step 0.40 s, api.py:133-134) and then calls the separator once per refined region (api.py:1073-1077).  The region
boundaries decide how many separator inputs a file produces and how long they are, i.e. the batch-shape mix the
batched driver (``SepformerSeparation.separate_regions`` / ``sharding.separate_sharded``) sees.

Everything here is list / float logic on the host; the speaker-embedding network (pyannote, out of scope) is a
callable the caller passes in.  The functions restate the reference's decisions -- same thresholds, same tie
handling, same clipping -- so that feeding their output to the batched driver separates exactly the spans the
reference's loop would have separated one by one.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Iterable, Optional, Sequence

Segment = tuple[float, float, str]            # (start s, end s, speaker label)


@dataclass
class FeederConfig:
    """The knobs of ``Config`` (api.py:112-135) this step reads."""
    sample_rate: int = 16000                  # target_sample_rate, api.py:115
    min_segment_duration: float = 0.45        # api.py:116: shorter diarization segments are skipped (api.py:1380)
    overlap_threshold: float = 0.50           # api.py:117: shorter overlap regions are ignored (api.py:888)
    window: float = 0.80                      # sliding_window_size, api.py:133
    step: float = 0.40                        # sliding_window_step, api.py:134


# ----------------------------------------------------------------------------------------------- overlap detection
def find_overlaps(segments: Iterable[Segment]) -> list[tuple[float, float, list[str]]]:
    """Sweep over segment boundaries (``find_segment_overlaps``, api.py:323-343): a region is emitted at every
    segment END that happens while more than one speaker is active, running from the instant the second speaker
    joined.  With three or more concurrent speakers the start is not reset until at most one remains, so regions
    may nest; regions with identical (start, end) collapse to the last one seen (the reference keys a dict by it)."""
    events = []
    for start, end, speaker in segments:
        events.append((start, +1, speaker))
        events.append((end, -1, speaker))
    events.sort(key=lambda e: (e[0], e[1]))    # at equal times ends (-1) come before starts (+1)
    active: set[str] = set()
    opened: Optional[float] = None
    found: dict[tuple[float, float], list[str]] = {}
    for when, kind, speaker in events:
        if kind > 0:
            active.add(speaker)
            if len(active) > 1 and opened is None:
                opened = when
        else:
            if len(active) > 1 and opened is not None:
                found[(opened, when)] = list(active)
            active.discard(speaker)
            if len(active) <= 1:
                opened = None
    return [(a, b, spk) for (a, b), spk in found.items()]


def detect_overlap_regions(segments: Iterable[Segment], cfg: FeederConfig = FeederConfig()):
    """``_detect_overlap_regions`` (api.py:881-891): overlaps of at least ``overlap_threshold`` seconds."""
    return [(a, b, spk) for a, b, spk in find_overlaps(segments) if (b - a) >= cfg.overlap_threshold and len(spk) > 1]


# ----------------------------------------------------------------------------------------------- sample slicing
def slice_indices(start: float, end: float, n_samples: int, sample_rate: int) -> Optional[tuple[int, int]]:
    """``_extract_segment`` (api.py:840-860): clamp to [0, duration], truncate times to sample indices.  ``None``
    stands for the reference's ``zeros(1, 100)`` placeholder (empty or inverted range)."""
    duration = n_samples / sample_rate
    start = max(start, 0.0)
    end = min(end, duration)
    i0, i1 = int(start * sample_rate), int(end * sample_rate)
    return None if i0 >= i1 else (i0, i1)


def cosine(a: Sequence[float], b: Sequence[float]) -> float:
    """``_calculate_embedding_similarity`` (api.py:877-878) on plain sequences / tensors."""
    try:
        import torch
        if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
            return torch.nn.functional.cosine_similarity(a, b, dim=0).item()
    except ImportError:                         # pragma: no cover
        pass
    num = sum(x * y for x, y in zip(a, b))
    na, nb = sum(x * x for x in a) ** 0.5, sum(y * y for y in b) ** 0.5
    return num / max(na * nb, 1e-8)


# ----------------------------------------------------------------------------------------------- re-segmentation
def _pick_speaker(scores: list[tuple[str, float]], previous: Optional[str]) -> tuple[str, float]:
    """The window's speaker (api.py:985-1001): best cosine, except that a narrow win (< 0.15) over the PREVIOUS
    window's speaker whose score is still > 65 % of the winner's keeps the previous speaker."""
    ranked = sorted(scores, key=lambda kv: kv[1], reverse=True)
    best, best_score = ranked[0]
    if len(ranked) > 1:
        runner, runner_score = ranked[1]
        if (best_score - runner_score) < 0.15 and previous and previous != best:
            if runner == previous and runner_score > 0.65 * best_score:
                return previous, runner_score
    return best, best_score


def resegment_overlap(n_samples: int, seg_start: float, seg_end: float, profiles: dict,
                      embed_fn: Callable[[int, int], Optional[object]], cfg: FeederConfig = FeederConfig()) -> list[Segment]:
    """``_resegment_overlap`` (api.py:961-1050) for one diarization segment of ``n_samples`` samples spanning
    [seg_start, seg_end] seconds.  ``embed_fn(i0, i1)`` returns the speaker embedding of samples [i0, i1) of the
    segment, or ``None`` when it cannot (the reference refuses windows shorter than half a second, api.py:864);
    ``profiles`` maps speaker label -> reference embedding.  Returns the refined (start, end, speaker) regions."""
    window, step = cfg.window, cfg.step
    duration = seg_end - seg_start
    if duration < 2.0:                                     # api.py:966-968: finer steps for short segments
        step = min(step, duration / 4)
    floor_len = min(0.3, duration / 10)                    # api.py:1022: shortest region kept

    votes: list[tuple[float, float, str, float]] = []
    previous: Optional[str] = None
    at = seg_start
    while at + window <= seg_end:                          # api.py:974
        idx = slice_indices(at - seg_start, at - seg_start + window, n_samples, cfg.sample_rate)
        emb = embed_fn(*idx) if idx is not None else embed_fn(0, 0)
        if emb is not None:
            who, conf = _pick_speaker([(s, cosine(emb, p)) for s, p in profiles.items()], previous)
            previous = who
        else:                                              # api.py:1005-1008: keep continuity
            who, conf = (previous if previous else "UNKNOWN"), 0.0
        votes.append((at, at + window, who, conf))
        at += step
    if not votes:
        return [(seg_start, seg_end, "UNKNOWN")]           # api.py:1013-1014: segment shorter than one window

    # consecutive windows of one speaker fuse when the gap between them is small (api.py:1017-1031)
    runs: list[Segment] = []
    run_a, run_b, run_who, _ = votes[0]
    for a, b, who, _conf in votes[1:]:
        if who == run_who and a - run_b <= max(step * 1.5, 0.2):
            run_b = b
        else:
            if (run_b - run_a) >= floor_len:
                runs.append((run_a, run_b, run_who))
            run_a, run_b, run_who = a, b, who
    if (run_b - run_a) >= floor_len:
        runs.append((run_a, run_b, run_who))

    # clip to the segment; a region that came out too short may borrow from a long predecessor (api.py:1033-1050)
    out: list[Segment] = []
    for i, (a, b, who) in enumerate(runs):
        a, b = max(seg_start, a), min(seg_end, b)
        if b - a < floor_len and i > 0 and out:
            pa, pb, pwho = out[-1]
            if pb - pa > floor_len * 1.5:
                need = floor_len - (b - a)
                pb -= min(need, pb - pa - floor_len)
                a = pb
                out[-1] = (pa, pb, pwho)
        if b - a >= floor_len:
            out.append((a, b, who))
    return [(max(seg_start, a), min(seg_end, b), who) for a, b, who in out]


# ----------------------------------------------------------------------------------------------- whole file
@dataclass
class SeparatorInput:
    """One ``separate_batch`` call of the reference: samples [i0, i1) of the file, for ``speaker``'s region."""
    i0: int
    i1: int
    speaker: str
    segment: Segment                                        # the diarization segment it came from


def separator_inputs(segments: Sequence[Segment], n_samples: int, profiles: dict,
                     embed_fn: Callable[[int, int], Optional[object]], cfg: FeederConfig = FeederConfig()) -> list[SeparatorInput]:
    """Every span of a file the reference would hand to ``separate_batch``, in its order: the segment loop of
    api.py:1378-1394 (skip segments shorter than ``min_segment_duration``; a segment is an overlap segment when it
    intersects any detected overlap region) followed by the region loop of api.py:1070-1077.  ``embed_fn(i0, i1)``
    takes FILE sample indices.  Two speakers' segments over the same overlap both qualify, so identical spans recur;
    ``sharding.dedupe_spans`` / ``separate_regions`` separate those once."""
    regions = detect_overlap_regions(segments, cfg)
    sr = cfg.sample_rate
    out: list[SeparatorInput] = []
    for seg in segments:
        s0, s1, _spk = seg
        if (s1 - s0) < cfg.min_segment_duration:
            continue
        if not any(max(s0, a) < min(s1, b) for a, b, _ in regions):
            continue
        cut = slice_indices(s0, s1, n_samples, sr)          # api.py:1394: the segment's audio
        if cut is None:
            continue
        base, seg_len = cut[0], cut[1] - cut[0]
        for a, b, who in resegment_overlap(seg_len, s0, s1, profiles, lambda i0, i1: embed_fn(base + i0, base + i1) if i1 > i0 else None, cfg):
            idx = slice_indices(a - s0, b - s0, seg_len, sr)  # api.py:1074
            if idx is None:
                continue                                     # the reference separates a zeros(1, 100) placeholder here
            out.append(SeparatorInput(base + idx[0], base + idx[1], who, seg))
    return out


# ----------------------------------------------------------------------------------------------- synthetic meeting
def synthetic_meeting(duration_s: float = 3600.0, overlap_fraction: float = 0.20, seed: int = 4):
    """A two-speaker diarization timeline with about ``overlap_fraction`` of its duration overlapped, and a toy
    embedding oracle for it: (segments, profiles, embed_fn_factory).  ``embed_fn_factory(sample_rate)`` returns an
    ``embed_fn(i0, i1)`` whose embedding is the time-weighted mix of the two speakers' unit vectors over the window
    (the later-starting speaker weighted 1.3x inside overlaps, so windows flip speakers part-way through a region),
    ``None`` for windows under half a second.  Deterministic; used by bench.py's meeting block and the CPU tests."""
    import random
    rng = random.Random(seed)
    segments: list[Segment] = []
    t = 0.0
    who = 0
    while t < duration_s:
        turn = rng.uniform(2.0, 14.0)
        end = min(duration_s, t + turn)
        segments.append((t, end, f"SPEAKER_{who:02d}"))
        # the next speaker barges in before this turn ends: overlap length log-uniform on [0.5, 8] s, sized so that
        # overlapped time / total time comes out near overlap_fraction
        mean_turn, mean_ov = 8.0, 2.7
        p_overlap = min(1.0, overlap_fraction * mean_turn / mean_ov)
        if rng.random() < p_overlap:
            import math
            ov = math.exp(rng.uniform(math.log(0.5), math.log(8.0)))
            nxt = max(t + 0.5, end - min(ov, turn * 0.8))
        else:
            nxt = end + rng.uniform(0.05, 0.6)
        t = nxt
        who ^= 1
    profiles = {"SPEAKER_00": (1.0, 0.0, 0.2), "SPEAKER_01": (0.0, 1.0, 0.2)}

    def factory(sample_rate: int):
        def embed(i0: int, i1: int):
            if i1 - i0 < sample_rate / 2:
                return None
            a, b = i0 / sample_rate, i1 / sample_rate
            w = [0.0, 0.0]
            for s0, s1, spk in segments:
                if s1 <= a or s0 >= b:
                    continue
                lo, hi = max(a, s0), min(b, s1)
                k = int(spk[-2:])
                # inside an overlap the speaker who started later dominates slightly
                later = any(o0 < s0 < o1 for o0, o1, ospk in segments if ospk != spk)
                w[k] += (hi - lo) * (1.3 if later else 1.0)
            if w[0] + w[1] == 0.0:
                return (0.0, 0.0, 1.0)
            return (w[0], w[1], 0.2 * (w[0] + w[1]))
        return embed
    return segments, profiles, factory
