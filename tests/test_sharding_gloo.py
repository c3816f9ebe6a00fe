"""world_size-2 gloo test of the multi-GPU host logic: instances are
sharded in contiguous blocks with no collective on the hot path; the only collective is the
optional gather of frames to rank 0.  Runs on CPU: the per-rank "frames" are a deterministic
function of the global instance index standing in for the kernel output."""
import os
import socket

import numpy as np
import pytest

from synchropmu_b200.api import shard_instances

torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_inst, n_ch, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_instances(n_inst, rank, world)
    idx = torch.arange(lo, hi, dtype=torch.float64)
    frames = (idx[:, None, None] * 10 + torch.arange(n_ch, dtype=torch.float64)[None, :, None] + torch.arange(4, dtype=torch.float64)[None, None, :] / 10).contiguous()
    sizes = [shard_instances(n_inst, r, world) for r in range(world)]
    bufs = [torch.empty((h - l, n_ch, 4), dtype=torch.float64) for l, h in sizes] if rank == 0 else None
    dist.gather(frames, bufs, dst=0) if all(h - l == hi - lo for l, h in sizes) else _ragged_gather(dist, frames, bufs, sizes, rank)
    if rank == 0:
        q.put(torch.cat(bufs).numpy())
    dist.barrier()
    dist.destroy_process_group()


def _ragged_gather(dist, frames, bufs, sizes, rank):
    if rank == 0:
        bufs[0].copy_(frames)
        for r in range(1, len(sizes)):
            dist.recv(bufs[r], src=r)
    else:
        dist.send(frames, dst=0)


@pytest.mark.parametrize("n_inst", [4096, 4097])
def test_shards_cover_everything_and_gather_preserves_order(n_inst):
    import torch.multiprocessing as mp

    world, n_ch = 2, 6
    covered = []
    for r in range(world):
        lo, hi = shard_instances(n_inst, r, world)
        covered += list(range(lo, hi))
    assert covered == list(range(n_inst))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_inst, n_ch, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(n_inst, dtype=np.float64)
    want = idx[:, None, None] * 10 + np.arange(n_ch)[None, :, None] + np.arange(4)[None, None, :] / 10
    assert np.array_equal(got, want)
