"""CPU, world_size 2, gloo: the N > 1 host logic (sharding + logits gather) of whisper_at.parallel."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whisper_at.parallel import shard_bounds, tag_sharded


def fake_tag(a: torch.Tensor) -> torch.Tensor:
    """stand-in for model.tag_batch: per-clip, deterministic, [n, 3, 5]"""
    s = a.double().sum(dim=1, keepdim=True)
    return (s[:, :, None] * torch.arange(1, 16, dtype=torch.float64).reshape(1, 3, 5)).float()


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    audio = torch.arange(n * 7, dtype=torch.float32).reshape(n, 7)
    out = tag_sharded(fake_tag, audio)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    for n in (0, 1, 5, 8, 1024):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n", [8, 5, 1])
def test_two_rank_gather_matches_single_process(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = fake_tag(torch.arange(n * 7, dtype=torch.float32).reshape(n, 7))
    for r in range(2):
        assert torch.equal(got[r], ref)
