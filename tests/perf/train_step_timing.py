"""Stage-1 training step at the reference config's own sizes (stage1_cars_recons16v: 4 scenes / GPU, 16 views 128x128, 15 extra
code-only steps + 1 joint step, 4096 rays per scene per step).  Prints ms per train_step and the cost of the fused backward with and
without decoder-weight gradients.   python tests/perf/train_step_timing.py [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.common import GOLDEN, spiral_poses  # noqa: E402


def main():
    import ssdnerf_b200 as S
    from ssdnerf_b200 import renderer as R
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    cuda = torch.device('cuda:0')
    c = json.load(open(os.path.join(GOLDEN, 'reference_configs.json')))['configs/paper_cfgs/stage1_cars_recons16v.py']
    train_cfg = {k: v for k, v in c['train_cfg'].items() if k != 'cache_load_from'}
    torch.manual_seed(0)
    model = S.build_model(dict(c['model'], cache_size=64), train_cfg=train_cfg, test_cfg=c['test_cfg']).to(cuda).train()
    B, V, res = c['samples_per_gpu'], 16, 128
    g = torch.Generator().manual_seed(0)
    poses = torch.from_numpy(spiral_poses(V))[None].repeat(B, 1, 1, 1).to(cuda)
    intr = torch.tensor([131.25, 131.25, 64.0, 64.0]).expand(B, V, 4).contiguous().to(cuda)
    code = (torch.randn(B, 3, 6, 128, 128, generator=g) * 0.5).to(cuda)
    with torch.no_grad():
        _, bits = model.get_density(model.decoder, code, cfg=dict(density_thresh=0.1))
        imgs, _ = model.render(model.decoder, code, bits, res, res, intr, poses, cfg=dict(dt_gamma_scale=0.5))
    imgs = imgs.clamp(0, 1)
    opt = dict(decoder=torch.optim.Adam(model.decoder.parameters(), lr=1e-3))
    data = dict(scene_id=list(range(B)), scene_name=[f's{i}' for i in range(B)], cond_imgs=imgs, cond_poses=poses, cond_intrinsics=intr)
    for _ in range(3):
        out = model.train_step(data, opt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = model.train_step(data, opt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res_d = dict(workload=f'stage1_cars_recons16v: {B} scenes x {V} views {res}x{res}, extra_scene_step {train_cfg["extra_scene_step"]}, '
                          f'{train_cfg["n_decoder_rays"]} rays/scene/step', ms_per_train_step=ms, scenes_per_sec=B / ms * 1e3,
                 log_vars=out['log_vars'])
    # fused backward alone, with / without decoder-weight gradients (same rays)
    n = train_cfg['n_decoder_rays']
    ro, rd = R.get_cam_rays(poses, intr, res, res)
    sel = torch.randint(0, V * res * res, (B, n), device=cuda)
    ro = ro.reshape(B, -1, 3).gather(1, sel[..., None].expand(-1, -1, 3)).contiguous()
    rd = rd.reshape(B, -1, 3).gather(1, sel[..., None].expand(-1, -1, 3)).contiguous()
    planes = R.pack_planes(code, R.DEC_P)
    blob = model.decoder.packed_blob()
    dtg = torch.full((B,), 0.5 / 131.25, device=cuda)
    fw = R.render_train_fwd(planes, (128, 128), bits, blob, ro, rd, noises=None, dt_gamma=dtg, want_counts=True)
    gi = torch.randn(B, n, 3, device=cuda)
    for want in (False, True):
        for _ in range(3):
            R.render_train_bwd(planes, (128, 128), bits, blob, ro, rd, fw['weights_sum'], fw['image'], None, gi, dt_gamma=dtg, want_decoder_grad=want)
        e0.record()
        for _ in range(20):
            R.render_train_bwd(planes, (128, 128), bits, blob, ro, rd, fw['weights_sum'], fw['image'], None, gi, dt_gamma=dtg, want_decoder_grad=want)
        e1.record(); torch.cuda.synchronize()
        res_d['bwd_ms_weight_grads' if want else 'bwd_ms_code_only'] = e0.elapsed_time(e1) / 20
    res_d['samples_per_bwd'] = int(fw['num_samples'].sum())
    print(json.dumps(res_d))


if __name__ == '__main__':
    main()
