"""FID / KID statistics (ssdnerf_b200/metrics.py) vs the formulation the reference uses (NumPy mean / np.cov, scipy sqrtm Frechet distance
[mmgen FID._calc_fid, mmgen-memory], the StyleGAN2-ADA KID estimator of lib/core/evaluation/metrics.py:160-184) and the world_size-2
all-gather of `feed` on gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy import linalg

from ssdnerf_b200.metrics import FIDKID, feature_statistics, frechet_distance, kernel_inception_distance


def test_frechet_distance_and_statistics_match_scipy():
    g = torch.Generator().manual_seed(0)
    a = torch.randn(400, 24, generator=g) @ torch.randn(24, 24, generator=g) * 0.3 + 0.2
    b = torch.randn(300, 24, generator=g) @ torch.randn(24, 24, generator=g) * 0.25
    m1, c1 = feature_statistics(a)
    m2, c2 = feature_statistics(b)
    np.testing.assert_allclose(m1.numpy(), np.mean(a.numpy().astype(np.float64), 0), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(c1.numpy(), np.cov(a.numpy().astype(np.float64), rowvar=False), rtol=1e-10, atol=1e-12)
    covmean = linalg.sqrtm(c1.numpy() @ c2.numpy())
    ref_cov = np.trace(c1.numpy()) + np.trace(c2.numpy()) - 2 * np.trace(covmean.real)
    ref_mean = float(((m1 - m2) ** 2).sum())
    fid, mean, cov = frechet_distance(m1, c1, m2, c2)
    assert abs(mean - ref_mean) < 1e-10 and abs(cov - ref_cov) < 1e-6 * abs(ref_cov) and abs(fid - ref_mean - ref_cov) < 1e-6 * abs(fid)


def test_kid_estimator_matches_reference_loop():
    g = torch.Generator().manual_seed(1)
    real, fake = torch.randn(120, 16, generator=g), torch.randn(90, 16, generator=g) * 1.1 + 0.1
    gen = torch.Generator().manual_seed(5)
    kid = kernel_inception_distance(real, fake, num_subsets=7, max_subset_size=50, generator=gen)
    gen = torch.Generator().manual_seed(5)             # same subsets, the reference's per-subset loop in NumPy
    n, m = 16, 50
    ix = [torch.randperm(90, generator=gen)[:m] for _ in range(7)]
    iy = [torch.randperm(120, generator=gen)[:m] for _ in range(7)]
    t = 0
    for i in range(7):
        x, y = fake[ix[i]].double().numpy(), real[iy[i]].double().numpy()
        a = (x @ x.T / n + 1) ** 3 + (y @ y.T / n + 1) ** 3
        b = (x @ y.T / n + 1) ** 3
        t += (a.sum() - np.diag(a).sum()) / (m - 1) - b.sum() * 2 / m
    assert abs(kid - t / 7 / m) < 1e-12 * max(1.0, abs(kid))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    metric = FIDKID(num_images=10, num_subsets=3, max_subset_size=8, feature_fn=lambda x: x.flatten(1)[:, :6])
    g = torch.Generator().manual_seed(100 + rank)
    kept = []
    for _ in range(3):                                   # 3 feeds x 2 ranks x 2 images = 12 >= 10: the last feed is truncated
        kept.append(metric.feed(torch.randn(2, 3, 2, 2, generator=g), 'fakes'))
        metric.feed(torch.randn(2, 3, 2, 2, generator=g), 'reals')
    fid, mean, cov, kid = metric.summary(generator=torch.Generator().manual_seed(0))
    q.put((rank, kept, torch.cat(metric.fake_feats).tolist(), fid, kid))
    dist.destroy_process_group()


def test_feed_all_gathers_features_world_size_2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [4, 4, 2] and res[0][2] == res[1][2] and len(res[0][2]) == 10      # every rank holds the same gathered rows
    assert res[0][3] == res[1][3] and res[0][4] == res[1][4] and np.isfinite(res[0][3])
