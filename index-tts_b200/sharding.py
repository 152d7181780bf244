"""Multi-GPU host logic: the path shards embarrassingly by utterance, one process
per GPU.  The only communication is off the critical path: speaker latents are broadcast once per
speaker and finished waveforms of ragged length are gathered on rank 0.  Works with any
torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
from typing import Dict, List, Sequence

import torch


def lpt_assign(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of utterances (by expected token count) to ranks
    (SURVEY §8d config 5).  Returns, per rank, the list of utterance indices in execution order."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += costs[i]
    return out


def broadcast_latents(dist, latents: Dict[str, torch.Tensor], src: int = 0) -> Dict[str, torch.Tensor]:
    """Broadcast the per-speaker conditioning (prompt_condition, ref_mel, style, emo_vec …) from the
    rank that ran the prompt encoders.  Shapes must already agree on all ranks (fixed by the speaker)."""
    for k in sorted(latents):
        dist.broadcast(latents[k], src=src)
    return latents


def gather_wavs(dist, wav: torch.Tensor, rank: int, world: int, dst: int = 0):
    """Gather ragged int16 waveforms on `dst`: lengths first, then zero-padded payloads.
    Returns the list of trimmed waveforms on dst, None elsewhere."""
    raw = wav.reshape(-1).contiguous().view(torch.uint8)   # int16 is not a collective dtype: ship bytes
    n = torch.tensor([raw.numel()], dtype=torch.int64, device=wav.device)
    lens = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(lens, n)
    m = int(max(int(x) for x in lens))
    pad = torch.zeros(m, dtype=torch.uint8, device=wav.device)
    pad[: raw.numel()] = raw
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return [b[: int(l)].view(wav.dtype) for b, l in zip(bufs, lens)]
