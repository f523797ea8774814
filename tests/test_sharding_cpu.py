"""N>1 path on CPU: world_size-2 gloo processes shard scenes like the reference's DistributedSampler and gather results."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ssdnerf_b200.sharding import gather_scene_outputs, scene_indices


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, num_scenes, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    idx = scene_indices(num_scenes)
    local = torch.stack([torch.full((3, 4), float(i)) for i in idx]) if idx else torch.zeros(0, 3, 4)
    full = gather_scene_outputs(local, num_scenes)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, idx, full[:, 0, 0].tolist(), float(t)))
    dist.destroy_process_group()


def test_scene_striding_matches_reference_sampler():
    assert scene_indices(7, 0, 2) == [0, 2, 4, 6] and scene_indices(7, 1, 2) == [1, 3, 5]
    assert sorted(sum((scene_indices(704, r, 8) for r in range(8)), [])) == list(range(704))
    assert scene_indices(3, 5, 8) == []          # ragged: more ranks than scenes


def test_world_size_2_gloo_gather():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, idx, full, tmax in res:
        assert full == [0.0, 1.0, 2.0, 3.0, 4.0] and tmax == 11.0
        assert idx == list(range(rank, 5, 2))
