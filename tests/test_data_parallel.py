"""CPU, world_size 2 over gloo: sharding the replay batch across ranks with one
gradient all-reduce per optimizer (reference tfutils.py:221-223) reproduces the
single-process update of the whole batch.  Runs the product's engine on the
primitive simulator; the NCCL path only swaps the collective backend."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dreamer_oracle as O
from daydreamer_b200 import engine as E
from tests import util
from tests.hostsim import HostSim

B, L, WORLD = 4, 5, 2
SPECS = {
    'proprio': {},
    # image + one-hot actions (xarm/ur5 style): conv weight grads and the scalar entropy controller
    'image_onehot': dict(discrete=True, act_dim=5, cnn_keys={'image': (10, 10, 3)}, cnn_depth=4,
                         enc_kernels=(4, 4), dec_kernels=(4, 4)),
}


def shard(batch, noise, spec, rank):
  per = B // WORLD
  sl = slice(rank * per, (rank + 1) * per)
  b = {k: v[sl] for k, v in batch.items()}
  H1 = spec.horizon + 1
  n = {'obs_post': noise['obs_post'][:, sl],
       'img': noise['img'].reshape(spec.horizon, B, L, spec.stoch, spec.classes)[:, sl].reshape(
           spec.horizon, per * L, spec.stoch, spec.classes),
       'act': noise['act'].reshape(H1, B, L, -1)[:, sl].reshape(H1, per * L, -1)}
  return b, n


def worker(rank, port, path, name):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=WORLD)
  torch.set_num_threads(1)
  ospec = O.tiny_spec(**SPECS[name])
  spec = util.to_engine_spec(ospec)
  eng = E.Engine(spec, HostSim(), B // WORLD, L, world=WORLD,
                 allreduce=lambda t: dist.all_reduce(t))
  eng.load_params(O.init_params(ospec, seed=1))
  state = None
  for it in range(2):
    batch, noise = shard(util.make_batch(ospec, B, L, seed=10 + it),
                         util.make_noise(ospec, B, L, seed=20 + it), ospec, rank)
    util.load_batch(eng, spec, batch, noise, state)
    eng.launch_step(draw=False)
    state = {'deter': eng.prev_deter.numpy().copy(), 'stoch': eng.prev_stoch.numpy().copy()}
  out = eng.get_params()
  out['__wmkl'] = eng.wmkl_scale.numpy().copy()
  out['__actent'] = eng.actent_scale.numpy().copy()
  out['__norm'] = eng.norm_state.numpy().copy()
  out['__gnorm'] = np.array([eng.read_metrics()['model_grad_norm']])
  np.savez(os.path.join(path, f'rank{rank}.npz'), **out)
  dist.destroy_process_group()


@pytest.mark.parametrize('name', list(SPECS))
def test_two_ranks_match_single_process(name):
  ospec = O.tiny_spec(**SPECS[name])
  spec = util.to_engine_spec(ospec)
  ref = E.Engine(spec, HostSim(), B, L)
  ref.load_params(O.init_params(ospec, seed=1))
  state = None
  for it in range(2):
    util.load_batch(ref, spec, util.make_batch(ospec, B, L, seed=10 + it),
                    util.make_noise(ospec, B, L, seed=20 + it), state)
    ref.launch_step(draw=False)
    state = {'deter': ref.prev_deter.numpy().copy(), 'stoch': ref.prev_stoch.numpy().copy()}
  want = ref.get_params()
  with tempfile.TemporaryDirectory() as tmp:
    port = 29500 + os.getpid() % 2000
    mp.spawn(worker, args=(port, tmp, name), nprocs=WORLD, join=True)
    ranks = [dict(np.load(os.path.join(tmp, f'rank{r}.npz'))) for r in range(WORLD)]
  for k in want:
    # replicas stay bit-identical to each other, and match the single-process update
    assert np.array_equal(ranks[0][k], ranks[1][k]), k
    d = np.abs(ranks[0][k] - want[k])
    assert d.max() <= 2.2 * ospec.lr and d.mean() <= 5e-3 * ospec.lr, (k, d.max(), d.mean())
  # controllers see global-batch statistics on every rank
  assert np.allclose(ranks[0]['__wmkl'], ref.wmkl_scale.numpy())
  assert np.allclose(ranks[0]['__actent'], ref.actent_scale.numpy())
  assert np.allclose(ranks[0]['__norm'], ref.norm_state.numpy(), rtol=1e-5, atol=1e-8)
  assert abs(ranks[0]['__gnorm'][0] - ref.read_metrics()['model_grad_norm']) < 1e-3 * ranks[0]['__gnorm'][0]
