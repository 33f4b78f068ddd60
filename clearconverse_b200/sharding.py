"""Host-side sharding of independent overlap segments across the GPUs of one box.

Every ``separate_batch`` call in the product is one segment with B=1
(/root/reference/back/api.py:1073-1077) and segments share nothing but the weights, so the
path shards by units with NO data-path collective: weights are replicated,
segments are length-bucketed (bucket = number of 150-frame chunks, i.e. 1,200-sample steps),
buckets are cut into batches under a token budget, batches are dealt to ranks greedily by
algorithmic FLOPs, and results are gathered on the host into segment order.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

KSZ, STRIDE, CHUNK = 16, 8, 150
FLOP_PER_INTRA_TOKEN = 11_714_560          # 16 intra layers x 732,160
FLOP_PER_FRAME_ENDS = 78_080               # encoder + output_fc + mask-mul + decoder
FLOP_MEM_CHUNK_BASE = 8 * 655_360          # memory transformer, per chunk, GEMM part
FLOP_MEM_CHUNK_PER_SEQ = 8 * 512           # ... attention part, per chunk per sequence element


def dedupe_spans(spans):
    """(start, end) sample ranges -> (unique ranges in first-seen order, index of each input range in that list)."""
    seen, unique, inverse = {}, [], []
    for a, b in spans:
        key = (int(a), int(b))
        if key not in seen:
            seen[key] = len(unique)
            unique.append(key)
        inverse.append(seen[key])
    return unique, inverse


def frames_of(T: int) -> int:
    return (T - KSZ) // STRIDE + 1


def chunks_of(T: int) -> int:
    return frames_of(T) // CHUNK + 1


def flops_of(T: int, mem_seq_len: int | None = None) -> int:
    """Algorithmic FLOPs of separating one T-sample segment (multiply-add = 2)."""
    L, S = frames_of(T), chunks_of(T)
    s_seq = S if mem_seq_len is None else mem_seq_len
    return S * CHUNK * FLOP_PER_INTRA_TOKEN + L * FLOP_PER_FRAME_ENDS + S * (FLOP_MEM_CHUNK_BASE + FLOP_MEM_CHUNK_PER_SEQ * s_seq)


@dataclass
class Batch:
    indices: list[int] = field(default_factory=list)   # positions in the caller's segment list
    lens: list[int] = field(default_factory=list)
    chunks: int = 0
    flops: int = 0


def bucket_segments(lens: list[int], max_chunks_per_batch: int = 2048) -> list[Batch]:
    """Group segments of similar length into batches of at most ``max_chunks_per_batch`` chunks
    (150 frames each; 2048 chunks = 307,200 token rows ~ 2.8 GB of fp32 workspace).  Segments
    are sorted by chunk count so a batch holds near-equal lengths; a segment longer than the
    budget gets a batch of its own."""
    order = sorted(range(len(lens)), key=lambda i: (chunks_of(lens[i]), lens[i], i), reverse=True)
    batches, cur = [], Batch()
    for i in order:
        s = chunks_of(lens[i])
        if cur.indices and cur.chunks + s > max_chunks_per_batch:
            batches.append(cur)
            cur = Batch()
        cur.indices.append(i)
        cur.lens.append(lens[i])
        cur.chunks += s
        cur.flops += flops_of(lens[i])
    if cur.indices:
        batches.append(cur)
    return batches


def assign_to_ranks(batches: list[Batch], world_size: int) -> list[list[int]]:
    """Longest-processing-time greedy: batches by descending FLOPs, each to the least-loaded
    rank.  Returns, per rank, the list of batch indices it runs."""
    load = [0] * world_size
    out: list[list[int]] = [[] for _ in range(world_size)]
    for b in sorted(range(len(batches)), key=lambda j: batches[j].flops, reverse=True):
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(b)
        load[r] += batches[b].flops
    return out


def plan_shards(lens: list[int], world_size: int, max_chunks_per_batch: int = 2048, batches_per_rank: int = 2):
    """(batches, per-rank batch indices).  One rank: length-bucketed batches under the chunk budget.  Several ranks: the
    SEGMENTS are dealt to the ranks first (longest-processing-time greedy on algorithmic FLOPs: the loads differ by at
    most one short segment), then each rank's share is bucketed into ``batches_per_rank`` batches (two: one per lane of
    the pipelined driver -- every batch pays the memory transformer's fixed latency, so fewer and larger is faster;
    measured at 8 GPUs on the 1-hour meeting: 4.2 ms with five batches per rank)."""
    if world_size <= 1:
        batches = bucket_segments(lens, max_chunks_per_batch)
        return batches, [list(range(len(batches)))]
    load = [0] * world_size
    mine: list[list[int]] = [[] for _ in range(world_size)]
    for i in sorted(range(len(lens)), key=lambda j: (flops_of(lens[j]), -j), reverse=True):
        r = min(range(world_size), key=lambda k: (load[k], k))
        mine[r].append(i)
        load[r] += flops_of(lens[i])
    batches: list[Batch] = []
    per_rank: list[list[int]] = [[] for _ in range(world_size)]
    for r in range(world_size):
        if not mine[r]:
            continue
        chunks = sum(chunks_of(lens[i]) for i in mine[r])
        k = max(batches_per_rank, -(-chunks // max_chunks_per_batch))     # more batches only if the chunk budget demands it
        order = sorted(mine[r], key=lambda i: (chunks_of(lens[i]), lens[i], i), reverse=True)   # similar lengths together
        while True:
            target = chunks / k
            cuts: list[Batch] = [Batch() for _ in range(k)]
            seen = 0
            for i in order:
                c = chunks_of(lens[i])
                b = cuts[min(k - 1, int((seen + c / 2) / target))]      # the batch the segment's midpoint falls into
                b.indices.append(i); b.lens.append(lens[i]); b.chunks += c; b.flops += flops_of(lens[i])
                seen += c
            if all(b.chunks <= max_chunks_per_batch or len(b.indices) <= 1 for b in cuts) or k >= len(order):
                break
            k += 1                                                      # a batch came out over the workspace budget
        for b in cuts:
            if b.indices:
                per_rank[r].append(len(batches))
                batches.append(b)
    return batches, per_rank


class SharedResults:
    """The host-side gather of the sharded path without pickling: ONE result buffer [sum_i T_i, n_spk] float32 in POSIX
    shared memory (``/dev/shm``), mapped by every rank of the box and page-locked (``cudaHostRegister``), so each rank's
    device->host copies land directly at their segments' final offsets and rank 0 reads all results in place after a
    barrier.  The barrier is a counter array in the same mapping (no collective, no process group needed).

    Rank 0 creates the buffer; the other ranks attach (the caller makes sure creation happens first, e.g. by
    constructing it before / after a ``dist.barrier()`` or by creating it in the parent).  ``view(i)`` is segment i's
    [T_i, n_spk] slice of the mapping: copy it if it has to outlive ``close()``."""

    HEADER = 4096                                     # bytes reserved for the barrier counters (int64 per rank)

    def __init__(self, lens: list[int], name: str, rank: int, world_size: int, n_spk: int = 2, pin: bool = True):
        import numpy as np
        import torch
        self.lens, self.rank, self.world, self.n_spk = [int(t) for t in lens], rank, world_size, n_spk
        self.offs = [0]
        for t in self.lens:
            self.offs.append(self.offs[-1] + t)
        self.path = os.path.join("/dev/shm", name)
        self.nbytes = self.HEADER + self.offs[-1] * n_spk * 4
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(self.nbytes)
        self._map = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(self.nbytes,))
        self._ctr = self._map[:self.HEADER].view(np.int64)
        self.data = torch.from_numpy(self._map[self.HEADER:].view(np.float32)).view(self.offs[-1], n_spk)
        self._epoch = 0
        self._pinned = False
        if pin and torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.data.data_ptr(), self.data.numel() * 4, 0)
            self._pinned = int(rc) == 0

    def view(self, i: int):
        return self.data[self.offs[i]:self.offs[i + 1]]

    def barrier(self, timeout_s: float = 120.0):
        """All ranks of the box have written their results (spins on the shared counters)."""
        import time
        self._epoch += 1
        self._ctr[self.rank] = self._epoch
        t0 = time.perf_counter()
        while True:
            if all(int(self._ctr[r]) >= self._epoch for r in range(self.world)):
                return
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError("SharedResults.barrier: a rank did not arrive")

    def close(self):
        import torch
        if self._pinned:
            torch.cuda.cudart().cudaHostUnregister(self.data.data_ptr())
            self._pinned = False
        self.data = None
        self._ctr = None
        self._map = None
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def separate_sharded(segments, separate_fn, rank: int = 0, world_size: int = 1, group=None,
                     max_chunks_per_batch: int = 2048, gather: bool = True, pipeline=None, shared: "SharedResults | None" = None):
    """Run this rank's share of ``segments`` (list of 1-D float32 tensors) through
    ``separate_fn(list_of_segments) -> list of [T_i, n_spk] tensors`` and, if ``gather``,
    collect all results on rank 0 in the original order (host-side gather, no NCCL).

    ``pipeline`` (optional): a callable taking an iterable of segment lists and yielding the result lists in order
    (``SepformerSeparation.separate_stream``): this rank's batches then run through it, two in flight, instead of
    one ``separate_fn`` call after the other.

    ``shared`` (optional, a ``SharedResults`` built over the same segment lengths on every rank): results are copied
    device->host straight into the shared, page-locked buffer and the gather is a barrier -- rank 0 gets views of the
    buffer, the other ranks None.  Without it the gather pickles the tensors through ``dist.gather_object``.

    Returns (results or None on non-zero ranks, audio samples this rank processed)."""
    lens = [int(s.numel()) for s in segments]
    batches, per_rank = plan_shards(lens, world_size, max_chunks_per_batch)
    mine: dict[int, object] = {}
    samples = 0

    def take(idx, outs):
        nonlocal samples
        for i, o in zip(idx, outs):
            if shared is not None:
                shared.view(i).copy_(o, non_blocking=True)      # D2H to the segment's final place
            else:
                mine[i] = o
            samples += lens[i]

    if pipeline is not None:
        mine_batches = [batches[b].indices for b in per_rank[rank]]
        for idx, outs in zip(mine_batches, pipeline([segments[i] for i in idx] for idx in mine_batches)):
            take(idx, outs)
    else:
        for b in per_rank[rank]:
            idx = batches[b].indices
            take(idx, separate_fn([segments[i] for i in idx]))
    if shared is not None:
        import torch
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()           # this rank's copies have landed
        if not gather:
            return None, samples
        shared.barrier()
        return ([shared.view(i) for i in range(len(segments))] if rank == 0 else None), samples
    if not gather:
        return mine, samples
    if world_size == 1:
        return [mine[i] for i in range(len(segments))], samples
    import torch.distributed as dist
    payload = {i: o.detach().cpu() for i, o in mine.items()}
    gathered = [None] * world_size if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0, group=group)
    if rank != 0:
        return None, samples
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(segments))], samples


# ---------------------------------------------------------------------------------------------
# One long recording split across ranks (optional row): the intra blocks are per chunk, so
# every rank takes a contiguous chunk range; the only data-path exchange is the chunk summaries [S, 128].
SAMPLES_PER_CHUNK = CHUNK * 8          # 150 frames x stride 8


def split_long(T: int, parts: int):
    """Cut one recording of T samples into at most ``parts`` spans on chunk boundaries.
    Returns a list of (sample_start, sample_end, first_chunk, n_chunks, inner): inner spans are [1200 c0, 1200 c1 + 8)
    (whole chunks + the decoder's 8-sample tail, overlapping the next span by 8 samples), the last span runs to T and
    is an ordinary item (it carries upstream's padding chunk).  Every span holds at least one real frame."""
    L = frames_of(T)
    real = -(-L // CHUNK)                       # chunks that hold at least one real frame
    parts = max(1, min(parts, real))
    base, extra = divmod(real, parts)
    spans, c0 = [], 0
    for r in range(parts):
        n = base + (1 if r < extra else 0)
        last = r == parts - 1
        if last:
            spans.append((c0 * SAMPLES_PER_CHUNK, T, c0, chunks_of(T) - c0, False))
        else:
            spans.append((c0 * SAMPLES_PER_CHUNK, (c0 + n) * SAMPLES_PER_CHUNK + 8, c0, n, True))
        c0 += n
    return spans


def assemble_spans(spans, ests, T: int):
    """Overlap-add the per-span outputs ([len_r, n_spk], any device) into [T, n_spk]: neighbouring spans overlap by the
    8 samples where one holds the upper decoder taps of its last frame and the other the lower taps of its first."""
    import torch
    out = torch.zeros(T, ests[0].shape[1], dtype=ests[0].dtype, device=ests[0].device)
    for (a, b, _, _, _), e in zip(spans, ests):
        out[a:b] += e.to(out.device)
    return out


def separate_long(sep, mix, rank: int = 0, world_size: int = 1, group=None, parts: int | None = None):
    """mix: 1-D float32 recording.  With ``group`` (torch.distributed, one rank per GPU) every rank runs its span and
    the summaries / results are exchanged with all_gather; without a group the ``parts`` (default: world_size) spans
    run one after the other on this separator's device -- the same arithmetic, and a way to bound the workspace of a
    recording that is too long for one launch sequence.  Returns est [T, n_spk] on ``sep.device`` (every rank)."""
    import torch
    eng = sep._engine
    mix = mix.reshape(-1)
    T = int(mix.numel())
    spans = split_long(T, parts if parts is not None else world_size)
    n_parts = len(spans)
    S = spans[-1][2] + spans[-1][3]
    dev = sep.device
    local = list(range(n_parts)) if group is None else ([rank] if rank < n_parts else [])
    means = {}
    for r in local:                                   # phase 1: encoder + seg_model[0] -> chunk summaries
        a, b, c0, n, inner = spans[r]
        means[r] = eng.span_phase1(mix[a:b].to(dev).contiguous(), inner, sep.precision, lane=r)
    all_means = torch.zeros(S, 128, dtype=torch.float32, device=dev)
    if group is None:
        for r in local:
            all_means[spans[r][2]:spans[r][2] + spans[r][3]] = means[r]
    else:
        import torch.distributed as dist
        nmax = max(sp[3] for sp in spans)
        mine = torch.zeros(nmax, 128, dtype=torch.float32, device=dev)
        if local:
            mine[:spans[rank][3]] = means[rank]
        gathered = [torch.empty_like(mine) for _ in range(world_size)]
        dist.all_gather(gathered, mine, group=group)  # THE exchange of this path: S x 128 fp32 in total
        for r in range(n_parts):
            all_means[spans[r][2]:spans[r][2] + spans[r][3]] = gathered[r][:spans[r][3]]
    hc = eng.memory_block(all_means, sep.precision)   # every rank, redundantly (its gLN needs the whole sequence anyway)
    ests = {}
    for r in local:                                   # phase 2: seg_model[1] + mask + decoder
        a, b, c0, n, inner = spans[r]
        ests[r] = eng.span_phase2(b - a, hc[c0:c0 + n].contiguous(), inner, sep.precision, lane=r)
    if group is None:
        return assemble_spans(spans, [ests[r] for r in range(n_parts)], T)
    import torch.distributed as dist
    lmax = max(sp[1] - sp[0] for sp in spans)
    mine = torch.zeros(lmax, 2, dtype=torch.float32, device=dev)
    if local:
        mine[:ests[rank].shape[0]] = ests[rank]
    gathered = [torch.empty_like(mine) for _ in range(world_size)]
    dist.all_gather(gathered, mine, group=group)
    return assemble_spans(spans, [gathered[r][:spans[r][1] - spans[r][0]] for r in range(n_parts)], T)
