"""Per-scene training state kept between visits of a scene (stage-1 auto-decoder training).

Behaviour of lib/models/autodecoders/multiscene_nerf.py:19-31,46-56,74-183 (`out_dict_to`, the RAM cache, `cache_load_from`,
`save_dir` + file writers): every scene owns its pre-activation latent `code_`, its occupancy state (`density_grid`,
`density_bitfield`) and the state of its private code optimizer.  Entries use the reference's on-disk record
    dict(scene_id, scene_name, param=dict(code_, density_grid, density_bitfield), optimizer=<optimizer.state_dict()>)
so directories written by either implementation can be loaded by the other.

Layout choices made for this build: an entry's host tensors are allocated once (pinned when CUDA is present) and
overwritten in place on every visit, so a step's write-back is a handful of async D2H copies on the current stream; scene
files are written by a small thread pool rather than forked processes (torch.save releases the GIL in its I/O).
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

_UNCAST = ('density_grid', 'density_bitfield', 'step')        # never converted to the storage float type


def _finite_cast(t, dtype):
    """float tensor -> `dtype`, saturating instead of overflowing to inf (fp16 storage of fp32 latents / Adam moments)"""
    if t.dtype == dtype or not t.is_floating_point():
        return t
    fi = torch.finfo(dtype)
    return t.clamp(min=fi.min, max=fi.max).to(dtype)


def _host_like(t, dtype):
    pin = torch.cuda.is_available()
    return torch.empty(t.shape, dtype=dtype, device='cpu', pin_memory=pin)


def _store(dst, key, val, dtype):
    """write `val` into dst[key]: in place when a host tensor already exists, else allocate it"""
    if not isinstance(val, torch.Tensor):
        dst[key] = val
        return
    want = val.dtype if (key in _UNCAST or not val.is_floating_point()) else dtype
    src = _finite_cast(val.detach(), want)
    if isinstance(dst.get(key), torch.Tensor) and dst[key].shape == src.shape and dst[key].dtype == src.dtype:
        dst[key].copy_(src, non_blocking=True)
    else:
        buf = _host_like(src, src.dtype)
        buf.copy_(src, non_blocking=True)
        dst[key] = buf


def optimizer_state_record(state_dict, dtype, into=None):
    """optimizer.state_dict() -> host record with float state in `dtype` (`step` untouched); reuses the tensors of `into`"""
    out = dict(state=dict(), param_groups=state_dict['param_groups']) if into is None else into
    out['param_groups'] = state_dict['param_groups']
    for pid, st in state_dict['state'].items():
        slot = out['state'].setdefault(pid, dict())
        for k, v in st.items():
            _store(slot, k, v, dtype)
    return out


def restore_optimizer_state(optimizer, record):
    """Put a cached per-parameter state (moments, step) back into a freshly built optimizer.  Unlike `Optimizer.load_state_dict` the
    hyper-parameters of the record's `param_groups` are NOT restored: the optimizer keeps the learning rate etc. of the CURRENT
    train_cfg, which is what the reference's `optimizer_set_state` (lib/core/utils/misc.py:85-126) does.  Float state is cast to the
    parameter's dtype and device; `step` keeps its own."""
    saved = [pid for g in record['param_groups'] for pid in g['params']]
    live = [p for g in optimizer.param_groups for p in g['params']]
    if len(saved) != len(live):
        raise ValueError('cached optimizer state does not match the optimizer (different number of parameters)')
    for pid, param in zip(saved, live):
        st = record['state'].get(pid)
        if st is None:
            continue
        new = {}
        for k, v in st.items():
            if isinstance(v, torch.Tensor):
                if k == 'step':          # Adam keeps its step counter on the host unless capturable / fused: leave it where it is
                    v = v.clone()
                else:
                    v = v.to(device=param.device, dtype=param.dtype if v.is_floating_point() else None, copy=True)
            new[k] = v
        optimizer.state[param] = new


def shard_bounds(cache_size, world_size):
    """scene-index split points of the reference (multiscene_nerf.py:49): np.round(np.linspace(0, cache_size, ws + 1))"""
    return np.round(np.linspace(0, cache_size, num=world_size + 1)).astype(np.int64)


class SceneCache:
    """RAM cache of the scenes this rank owns (+ optional write-through to `save_dir`)."""

    def __init__(self, cache_size=0, rank=0, world_size=1, half=False, num_file_writers=0):
        self.cache_size, self.half = int(cache_size), bool(half)
        self.code_dtype = torch.float16 if half else torch.float32
        self.optimizer_dtype = torch.bfloat16 if half else torch.float32
        if cache_size > 0:
            b = shard_bounds(cache_size, world_size)
            self.entries = {int(i): None for i in range(b[rank], b[rank + 1])}
        else:
            self.entries = None
        self.loaded = False
        self._pool = ThreadPoolExecutor(num_file_writers) if num_file_writers > 0 else None
        self._pending = []

    # -- reads
    def load_dir(self, path):
        """`cache_load_from` (multiscene_nerf.py:80-95): sorted file list, file index == scene id; True if anything was loaded"""
        files = sorted(os.listdir(path))
        if not files:
            return False
        if len(files) != self.cache_size:
            raise ValueError(f'{path} holds {len(files)} scene files, cache_size is {self.cache_size}')
        for ind in self.entries:
            self.entries[ind] = torch.load(os.path.join(path, files[ind]), map_location='cpu')
        return True

    def get(self, scene_id):
        if self.entries is None:
            return None
        return self.entries[int(scene_id)]           # KeyError: the sampler handed this rank a scene it does not own

    # -- writes
    def record(self, scene_id, scene_name, code_, density_grid, density_bitfield, optimizer_state, into=None):
        out = into if into is not None else dict(param=dict(), optimizer=None)
        out.setdefault('scene_id', scene_id)
        out.setdefault('scene_name', scene_name)
        out['param'].pop('code', None)               # an activated code loaded from a stored scene is superseded by code_
        _store(out['param'], 'code_', code_, self.code_dtype)
        _store(out['param'], 'density_grid', density_grid, None)
        _store(out['param'], 'density_bitfield', density_bitfield, None)
        out['optimizer'] = optimizer_state_record(optimizer_state, self.optimizer_dtype, into=out.get('optimizer'))
        return out

    def put(self, scene_id, scene_name, code_, density_grid, density_bitfield, optimizer_state, save_dir=None):
        rec = None
        if self.entries is not None:
            sid = int(scene_id)
            rec = self.record(scene_id, scene_name, code_, density_grid, density_bitfield, optimizer_state, into=self.entries[sid])
            self.entries[sid] = rec
        if save_dir is not None:
            # a private copy: the cached entry is overwritten in place on the scene's next visit
            snap = self.record(scene_id, scene_name, code_, density_grid, density_bitfield, optimizer_state)
            self._write(snap, os.path.join(save_dir, f'{scene_name}.pth'))
        return rec

    def _write(self, rec, path):
        def job():
            if torch.cuda.is_available():
                torch.cuda.synchronize()             # the async D2H copies into `rec`
            torch.save(rec, path)
        if self._pool is None:
            job()
        else:
            self._pending = [f for f in self._pending if not f.done()]
            self._pending.append(self._pool.submit(job))

    def flush(self):
        for f in self._pending:
            f.result()
        self._pending = []
