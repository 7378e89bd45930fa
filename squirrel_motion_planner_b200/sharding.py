"""Multi-GPU plumbing for the candidate-edge batch (BASELINE.json configs[3]).

States are independent, so rank r owns the contiguous, 32-aligned range [r*S/R, (r+1)*S/R); the voxel grid and
the robot description are replicated; every rank writes its packed verdict bits straight into its slot of
the gather buffer and ONE all-gather (NCCL over NVLink on the GPU box, gloo in the CPU tests) gives every rank
the full mask — or the kernel stores the words into every rank's buffer itself (peer_gather.py, `sqmp_check_states_
gather_dev`).  No other collective exists on this path; planning-query batches are replicas only.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_states: int, rank: int, world: int):
    """Contiguous 32-aligned shard of [0, n_states) — every shard but the last has the same length."""
    per = ((n_states + world - 1) // world + 31) // 32 * 32
    lo = min(rank * per, n_states)
    hi = min(lo + per, n_states)
    return lo, hi, per


def mask_words(n_states: int) -> int:
    return (n_states + 31) // 32


def gather_masks(local_words, n_states: int, rank: int, world: int, group=None):
    """All-gather of the per-rank mask words (torch tensors, int32).  `local_words` is this rank's padded slot
    (per/32 words); returns the assembled mask of ceil(n_states/32) words, identical on every rank."""
    import torch
    import torch.distributed as dist
    _, _, per = shard_range(n_states, rank, world)
    slot = per // 32
    assert local_words.numel() == slot
    full = torch.empty(slot * world, dtype=local_words.dtype, device=local_words.device)
    if world > 1:
        dist.all_gather_into_tensor(full, local_words.contiguous(), group=group)
    else:
        full.copy_(local_words)
    return full[: mask_words(n_states)]


def unpack_mask(words: np.ndarray, n_states: int) -> np.ndarray:
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:n_states].astype(bool)
