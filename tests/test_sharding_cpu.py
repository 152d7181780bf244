"""CPU: the N > 1 host logic (utterance sharding, latent broadcast, ragged waveform gather) with the
gloo backend, world size 2 — the same functions bench.py runs over NCCL on the GPU box."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from indextts_b200.sharding import broadcast_latents, gather_wavs, lpt_assign


def test_lpt_assignment_balances_and_covers():
    costs = [768, 128, 400, 399, 700, 130, 512, 256, 300]
    a = lpt_assign(costs, 4)
    flat = sorted(i for r in a for i in r)
    assert flat == list(range(len(costs)))
    loads = [sum(costs[i] for i in r) for r in a]
    assert max(loads) - min(loads) <= max(costs)
    assert lpt_assign([], 2) == [[], []]
    assert lpt_assign([5], 3) == [[0], [], []]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1)
    ref = {"prompt_condition": torch.randn(17, 8, generator=g), "style": torch.randn(192, generator=g)}
    lat = {k: (v.clone() if rank == 0 else torch.zeros_like(v)) for k, v in ref.items()}
    broadcast_latents(dist, lat, src=0)
    ok = all(torch.equal(lat[k], ref[k]) for k in ref)
    wav = (torch.arange(100 + 50 * rank, dtype=torch.int16) + 1000 * rank)
    got = gather_wavs(dist, wav, rank, world, dst=0)
    if rank == 0:
        ok = ok and len(got) == world and all(
            torch.equal(got[r], torch.arange(100 + 50 * r, dtype=torch.int16) + 1000 * r) for r in range(world))
    else:
        ok = ok and got is None
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)   # the max-over-ranks timing reduction of bench.py
    ok = ok and float(t) == float(world)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_broadcast_and_ragged_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=60) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
