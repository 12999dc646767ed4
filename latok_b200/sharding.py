"""Multi-GPU sharding of a packed batch.

Every string is an independent unit (context fills at latok.c:69-73,114-134; the block mask restarts
per call), so a batch is cut into G contiguous ranges of whole strings, balanced by bytes, and every
GPU runs the identical single-GPU pipeline on its range.  There is no data-path collective.  The only
cross-GPU datum is each shard's token (and character) count, needed to rebase the per-shard CSR
offsets into global ones: one all-gather of two int64 per rank (NCCL when the ranks are processes on
GPUs, gloo on CPU), or a host-side prefix sum when one process drives all devices.
"""
from __future__ import annotations

import threading
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .engine import SPANS, SPLITS, BatchResult


def shard_ranges(offsets: np.ndarray, n_shards: int) -> List[Tuple[int, int]]:
    """Cut [0, S) into n_shards contiguous string ranges whose boundaries are the string boundaries
    nearest to k * B / n_shards (B = total bytes).  Ranges may be empty when S < n_shards."""
    offsets = np.asarray(offsets, dtype=np.int64)
    S = len(offsets) - 1
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    total = int(offsets[-1])
    cuts = [0]
    for k in range(1, n_shards):
        target = total * k / n_shards
        j = int(np.searchsorted(offsets, target, side="left"))
        j = min(max(j, 0), S)
        if j > 0 and abs(int(offsets[j - 1]) - target) <= abs(int(offsets[j]) - target):
            j -= 1
        cuts.append(max(j, cuts[-1]))
    cuts.append(S)
    return [(cuts[i], cuts[i + 1]) for i in range(n_shards)]


def slice_shard(buf: np.ndarray, offsets: np.ndarray, s0: int, s1: int):
    """Byte range + rebased offsets of strings [s0, s1) (views, no copy of the bytes)."""
    b0, b1 = int(offsets[s0]), int(offsets[s1])
    return buf[b0:b1], (offsets[s0:s1 + 1] - b0)


def merge_results(parts: Sequence[BatchResult]) -> BatchResult:
    """Concatenate per-shard results of consecutive string ranges into one batch result, rebasing the
    CSR offsets by the running character / token counts (the host-side form of the count exchange)."""
    out = BatchResult(sum(p.n_strings for p in parts), sum(p.n_chars for p in parts), sum(p.n_tokens for p in parts))

    def cat(name):
        arrs = [getattr(p, name) for p in parts]
        return None if any(a is None for a in arrs) else np.concatenate(arrs)

    def cat_offsets(name, counts):
        arrs = [getattr(p, name) for p in parts]
        if any(a is None for a in arrs):
            return None
        base = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
        pieces = [a[:-1] + b for a, b in zip(arrs, base)]
        return np.concatenate(pieces + [np.array([int(np.sum(counts))], dtype=np.int64)])

    out.splits, out.spans, out.tok_feats, out.matrix = cat("splits"), cat("spans"), cat("tok_feats"), cat("matrix")
    out.char_offsets = cat_offsets("char_offsets", [p.n_chars for p in parts])
    out.tok_offsets = cat_offsets("tok_offsets", [p.n_tokens for p in parts])
    out.kernel_ms = max((p.kernel_ms for p in parts), default=0.0)
    out.lookahead_walks = sum(p.lookahead_walks for p in parts)
    return out


_engines = {}
_engines_lock = threading.Lock()


def device_engine(device: int):
    """The process-wide engine of a device (created on first use; an Engine serialises its callers itself)."""
    from .engine import Engine
    with _engines_lock:
        e = _engines.get(device)
        if e is None:
            e = _engines[device] = Engine(device)
        return e


def close_engines():
    with _engines_lock:
        for e in _engines.values():
            e.close()
        _engines.clear()


def tokenize_sharded(buf: np.ndarray, offsets: np.ndarray, devices: Sequence[int], what: int = SPLITS | SPANS,
                     run_fn: Optional[Callable[[int, np.ndarray, np.ndarray, int], BatchResult]] = None) -> BatchResult:
    """One process, several GPUs: one host thread per shard and one long-lived Engine per device, each shard on its
    byte-balanced string range; results are concatenated on the host.  `run_fn(device, buf, offsets, what)` can be
    injected (tests use it to exercise the host logic without a GPU)."""
    ranges = shard_ranges(offsets, len(devices))
    results: List[Optional[BatchResult]] = [None] * len(devices)
    errors: List[Optional[BaseException]] = [None] * len(devices)

    def default_run(dev, b, o, w):
        return device_engine(dev).run_packed(b, o, w)      # one engine per device, kept alive between calls

    fn = run_fn or default_run

    def work(i):
        try:
            s0, s1 = ranges[i]
            b, o = slice_shard(buf, offsets, s0, s1)
            results[i] = fn(devices[i], np.ascontiguousarray(b), np.ascontiguousarray(o), what)
        except BaseException as exc:  # re-raised on the caller's thread
            errors[i] = exc

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for exc in errors:
        if exc is not None:
            raise exc
    return merge_results(results)


def allgather_counts(n_chars: int, n_tokens: int, group=None):
    """One rank per GPU (torchrun): exchange (characters, tokens) of every rank's shard -- the path's only
    collective, 16 bytes per rank -- and return (all_counts [world, 2], char_base, token_base) for this rank."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor([n_chars, n_tokens], dtype=torch.int64, device=device)
    out = torch.zeros(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    counts = out.view(world, 2).cpu().numpy()
    base = counts[:rank].sum(axis=0) if rank else np.zeros(2, dtype=np.int64)
    return counts, int(base[0]), int(base[1])


def rebase_for_rank(result: BatchResult, char_base: int, token_base: int) -> BatchResult:
    """Turn a rank's shard-local CSR offsets into offsets into the global (all ranks) arrays."""
    if result.char_offsets is not None:
        result.char_offsets = result.char_offsets + char_base
    if result.tok_offsets is not None:
        result.tok_offsets = result.tok_offsets + token_base
    return result
