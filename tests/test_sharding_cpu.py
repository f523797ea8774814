"""N>1 path on CPU: world_size-2 gloo processes shard scenes like the reference's DistributedSampler and gather results."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ssdnerf_b200.sharding import broadcast_scene, gather_scene_outputs, gather_views, max_views_per_rank, scene_indices, view_range


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
    # view sharding (strong scaling): rank 0 owns the scene, every rank renders its view range, images are gathered in view order
    code = torch.arange(6, dtype=torch.float32).reshape(1, 6) * (1 if rank == 0 else -1)
    bits = torch.full((1, 4), 7 if rank == 0 else 0, dtype=torch.uint8)
    broadcast_scene(code, bits)
    V = 7
    lo, hi = view_range(V)
    local = torch.arange(lo, hi, dtype=torch.uint8)[:, None].repeat(1, 3) + bits[0, 0]
    views = gather_views(local, V)
    # stage-1 training: the shared decoder's gradient is averaged over ranks (the reference's DDP wrapper), scene-cache ownership
    # follows the reference's linspace split
    from ssdnerf_b200.nerf import _average_grads_across_ranks
    from ssdnerf_b200.scene_cache import SceneCache
    lin = torch.nn.Linear(3, 2)
    lin.weight.grad = torch.full((2, 3), float(rank + 1))
    lin.bias.grad = torch.tensor([10.0 * rank, 1.0])
    _average_grads_across_ranks(lin)
    owned = sorted(SceneCache(5, rank, world).entries)
    q.put((rank, idx, full[:, 0, 0].tolist(), float(t), code[0].tolist(), views[:, 0].tolist(),
           (lin.weight.grad.unique().tolist(), lin.bias.grad.tolist(), owned)))
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
    for rank, idx, full, tmax, code, views, train in res:
        assert train[0] == [1.5] and train[1] == [5.0, 1.0]
        assert train[2] == ([0, 1] if rank == 0 else [2, 3, 4])        # np.round(np.linspace(0, 5, 3)) = [0, 2, 5]
        assert full == [0.0, 1.0, 2.0, 3.0, 4.0] and tmax == 11.0
        assert idx == list(range(rank, 5, 2))
        assert code == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0] and views == [7 + v for v in range(7)]


def test_view_ranges_partition_the_views():
    for V, W in ((251, 8), (251, 1), (7, 2), (3, 8)):
        r = [view_range(V, k, W) for k in range(W)]
        assert r[0][0] == 0 and r[-1][1] == V and all(r[k][1] == r[k + 1][0] for k in range(W - 1))
        assert max_views_per_rank(V, W) - min(hi - lo for lo, hi in r) <= 1
