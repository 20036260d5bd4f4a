"""CPU: config -> Spec resolution of the Agent (daydreamer_b200.agent.spec_from_config) for the
BASELINE.json configurations, against the sizes SURVEY.md 8 / A.8 derives from the reference's
configs.yaml, and loud rejection of settings whose kernels are not built."""
import numpy as np
import pytest

from daydreamer_b200.agent import spec_from_config
from daydreamer_b200.config import make_config
from daydreamer_b200 import engine as E


class Space:
  def __init__(self, shape, dtype=np.float32, discrete=False):
    self.shape, self.dtype, self.discrete = tuple(shape), np.dtype(dtype), discrete


def robot_spaces(joints, depth, act=6):
  obs = {'image': Space((64, 64, 3), np.uint8), 'cartesian_position': Space((6,)),
         'joint_positions': Space((joints,)), 'gripper_pos': Space((1,)),
         'gripper_side': Space((3,)), 'grasped_side': Space((3,)), 'reward': Space(()),
         'is_first': Space((), bool), 'is_last': Space((), bool), 'is_terminal': Space((), bool)}
  if depth:
    obs['depth'] = Space((64, 64, 1), np.uint8)
  return obs, {'action': Space((act,), np.float32, True), 'reset': Space((), bool, True)}


def a1_spaces():
  obs = {'vector': Space((16,)), 'image': Space((64, 64, 3), np.uint8), 'reward': Space(()),
         'is_first': Space((), bool), 'is_last': Space((), bool), 'is_terminal': Space((), bool)}
  return obs, {'action': Space((12,)), 'reset': Space((), bool, True)}


def test_a1_bench_config():
  spec = spec_from_config(make_config('a1', **{'rssm.deter': 512, 'rssm.units': 512}), *a1_spaces())
  assert (spec.deter, spec.units, spec.stoch, spec.classes) == (512, 512, 32, 32)
  assert spec.obs_dims == {'vector': 16} and not spec.cnn_keys and spec.embed == 512
  assert not spec.discrete and spec.act_dim == 12
  assert spec.discount == 0.995 and spec.actent_target == 0.5 and spec.actor_minstd == 0.1
  assert (spec.lr, spec.adam_eps, spec.clip, spec.wd) == (1e-4, 1e-6, 100.0, 1e-2)
  shapes = E.param_shapes(spec)
  assert shapes['rssm/img_in/kernel'] == (1024 + 12, 512)
  assert shapes['rssm/gru_out/kernel'] == (1024, 1536)
  assert shapes['rssm/obs_out/kernel'] == (512 + 512, 512)
  n = lambda pre: sum(int(np.prod(s)) for k, s in shapes.items() if k.startswith(pre))
  # model 10.01 M, critic 1.58 M, actor 1.59 M parameters
  assert abs(n(('enc/', 'rssm/', 'dec/', 'reward/', 'cont/')) - 10.01e6) < 0.02e6
  assert abs(n('critic/') - 1.58e6) < 0.01e6 and abs(n('actor/') - 1.59e6) < 0.01e6


def test_a1_shipped_block_is_256():
  spec = spec_from_config(make_config('a1'), *a1_spaces())       # A/configs.yaml:238
  assert (spec.deter, spec.units) == (256, 256)


@pytest.mark.parametrize('name,joints,depth,vec,channels', [('xarm', 7, True, 20, 4), ('ur5', 6, False, 19, 3)])
def test_robot_image_configs(name, joints, depth, vec, channels):
  spec = spec_from_config(make_config(name), *robot_spaces(joints, depth))
  assert spec.discrete and spec.act_dim == 6 and spec.actor_unimix == 0.1
  assert spec.actent_target == 0.4 and spec.discount == 0.995
  assert spec.obs_total == vec and spec.img_shape == (64, 64, channels)
  assert list(spec.cnn_keys) == (['image', 'depth'] if depth else ['image'])
  assert [s[4] for s in spec.enc_conv_shapes] == [31, 14, 6, 2]
  assert [s[4] for s in spec.dec_conv_shapes] == [5, 13, 30, 64]
  assert spec.cnn_out == 2048 and spec.embed == 2560 and spec.deter == 512
  shapes = E.param_shapes(spec)
  assert shapes['enc/conv0/kernel'] == (4, 4, channels, 64)
  assert shapes['dec/conv_out/kernel'] == (6, 6, channels, 64)
  assert shapes['actor/dist_out/out/kernel'] == (512, 6) and 'actor/dist_out/std/kernel' not in shapes
  init = E.init_params(spec, seed=3)
  lim = np.sqrt(3.0 / (25 * (256 + 1536) / 2.0))                  # A/nets.py:524 for dec/conv0
  assert np.abs(init['dec/conv0/kernel']).max() <= lim and np.abs(init['dec/conv0/kernel']).max() > 0.9 * lim


def test_large_rssm_config():
  """BASELINE.json configs[4]: deter 2048, 48x48 latent."""
  cfg = make_config('a1', **{'rssm.deter': 2048, 'rssm.units': 2048, 'rssm.stoch': 48,
                            'rssm.classes': 48, 'imag_horizon': 20})
  spec = spec_from_config(cfg, *a1_spaces())
  assert spec.sc == 2304 and spec.feat == 2048 + 2304 and spec.horizon == 20


@pytest.mark.parametrize('override', [
    {'critic_type': 'qfunction'}, {'expl_behavior': 'Explore'}, {'task_behavior': 'Random'}, {'rssm.classes': 0},
    {'actor_dist_cont': 'trunc_normal'}, {'rssm.act': 'gelu'}, {'reward_head.dist': 'mse'},
    {'model_opt.lr': 3e-4}, {'actor.layers': 3}, {'retnorm.impl': 'mean_std'}, {'priority_correct': 0.5},
    {'loss_scales.kl': 0.5}, {'grad_heads': ['decoder', 'reward']}, {'actent.vel': 0.2},
    {'encoder.cnn_keys': 'image', 'decoder.cnn_keys': 'image', 'encoder.cnn': 'resnet'},
    {'encoder.mlp_keys': '$^', 'decoder.mlp_keys': '$^'},      # nothing left to encode
])
def test_unbuilt_settings_raise(override):
  with pytest.raises(NotImplementedError):
    spec_from_config(make_config('a1', **override), *a1_spaces())


def test_batcher_drops_unused_keys_and_surfaces_errors():
  """Agent.dataset(): zip over batch_size replay generators (the role of embodied.Prefetch,
  E/core/prefetch.py:6-67); keys the learner never reads are dropped before stacking."""
  from daydreamer_b200.agent import Batcher

  def gen():
    i = 0
    while True:
      yield {'vector': np.full((5, 16), i, np.float32), 'image': np.zeros((5, 64, 64, 3), np.uint8),
             'action': np.zeros((5, 12), np.float32), 'log_extra': np.zeros((5,), np.float32)}
      i += 1
  it = iter(Batcher(gen, 4, keys={'vector', 'action', 'reward'}))
  b = next(it)
  assert set(b) == {'vector', 'action'} and b['vector'].shape == (4, 5, 16)
  assert set(next(iter(Batcher(gen, 3)))) == {'vector', 'image', 'action', 'log_extra'}

  def bad():
    yield {'vector': np.zeros((5, 16), np.float32)}
    raise RuntimeError('replay failed')
  it = iter(Batcher(bad, 2, workers=1))
  next(it)
  with pytest.raises(RuntimeError):
    next(it)


def test_default_key_regexes_give_the_encoder_more_keys_than_the_decoder():
  """Reference defaults `mlp_keys: '.*'` (A/configs.yaml:54-55): the encoder MLP takes every
  rank-0/1 key except is_first / is_last (A/nets.py:192-201: vector, reward, is_terminal), the
  decoder only rank-1 keys outside its exclusion list (A/nets.py:241-250: vector)."""
  obs, act = a1_spaces()
  cfg = make_config('a1', **{'encoder.mlp_keys': '.*', 'decoder.mlp_keys': '.*'})
  spec = spec_from_config(cfg, obs, act)
  assert spec.obs_dims == {'vector': 16, 'reward': 1, 'is_terminal': 1}
  assert spec.dec_dims == {'vector': 16} and spec.dec_total == 16 and spec.obs_total == 18
  shapes = E.param_shapes(spec)
  assert shapes['enc/dense0/kernel'] == (18, 512) and shapes['dec/dist_out/out/kernel'] == (512, 16)
