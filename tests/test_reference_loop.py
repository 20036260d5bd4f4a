"""CPU: the REFERENCE's own run loop (embodied/run/train.py: Driver + replay + Checkpoint +
Timer.wrap + Logger, read from /root/reference, never copied) drives daydreamer_b200.Agent
unchanged -- the drop-in claim of INTEGRATION.md, exercised where there is no GPU by handing the
Agent the torch-CPU primitive simulator (tests/hostsim.py) through its `_backend_for_tests` seam.
Covers: Agent.configs through embodied.Config + flags, spaces from the reference's Dummy env
(continuous and one-hot actions, image + vector observations), dataset() over the reference's
FixedLength replay, train/report/policy through Timer.wrap re-binding, report() outputs accepted
by embodied.Logger, save()/load() through embodied.Checkpoint (pickle).
Skipped when the reference checkout is absent (it is on the GPU box)."""
import pathlib
import sys
import types

import numpy as np
import pytest

REF = pathlib.Path('/root/reference')
pytestmark = pytest.mark.skipif(not (REF / 'embodied').exists(), reason='reference checkout absent')


@pytest.fixture(scope='module')
def embodied():
  if 'gym' not in sys.modules:                       # the reference imports gym at module level only
    gym = types.ModuleType('gym')
    gym.spaces = types.ModuleType('gym.spaces')
    sys.modules['gym'], sys.modules['gym.spaces'] = gym, gym.spaces
  sys.path.insert(0, str(REF))
  try:
    import embodied as emb
  finally:
    sys.path.remove(str(REF))
  return emb


@pytest.mark.parametrize('task', ['dummy_continuous', 'dummy_discrete'])
def test_reference_train_loop_drives_the_agent(embodied, task, tmp_path):
  from daydreamer_b200.agent import Agent
  from tests.hostsim import HostSim
  # exactly what the reference's train.py does with Agent.configs and command-line flags
  config = embodied.Config(Agent.configs['defaults'])
  config = config.update(Agent.configs['debug'])
  config = embodied.Flags(config).parse([
      '--task', task, '--logdir', str(tmp_path), '--run', 'train',
      '--encoder.mlp_keys', 'vector', '--decoder.mlp_keys', 'vector',
      '--encoder.cnn_keys', 'image', '--decoder.cnn_keys', 'image',
      '--rssm.deter', '32', '--encoder.cnn_depth', '4', '--decoder.cnn_depth', '4',
      '--batch_size', '2', '--replay_chunk', '6', '--imag_horizon', '3', '--env.length', '12',
      '--train.train_fill', '40', '--train.train_every', '8', '--train.log_every', '16',
      '--train.eval_every', '24', '--train.steps', '48', '--train.log_keys_video', 'image',
      # run/train.py asks for mode='explore' while step < expl_until: exercise behaviors.Random too
      '--expl_behavior', 'Random', '--train.expl_until', '20'])
  args = embodied.Config(logdir=config.logdir, **config.train)
  args = args.update(expl_until=args.expl_until // config.env.repeat)
  step = embodied.Counter()
  records = []

  class Collect:                                      # an embodied.Logger output
    def __call__(self, summaries):
      records.extend(summaries)

  logger = embodied.Logger(step, [Collect()], multiplier=1)
  env = embodied.envs.load_env(
      config.task, mode='train', logdir=str(tmp_path), amount=1, parallel='none',
      size=(64, 64), length=config.env.length)
  agent = Agent(env.obs_space, env.act_space, step, config, _backend_for_tests=HostSim())
  assert agent.spec.discrete == (task == 'dummy_discrete')
  assert list(agent.spec.cnn_keys) == ['image'] and list(agent.spec.obs_dims) == ['vector']
  store = embodied.replay.Stats(embodied.replay.RAMStore(1000))
  replay = embodied.replay.FixedLength(store, config.replay_chunk, **config.replay_fixed)
  try:
    embodied.run.train(agent, env, replay, logger, args)
  finally:
    env.close()
  assert int(step) >= 48
  names = {name for _, name, _ in records}
  for want in ('train/model_loss', 'train/actor_loss', 'train/image_loss_mean', 'report/openl_image',
               'report/task_imag_image', 'report/model_loss_mean', 'episode/score'):
    assert want in names, (want, sorted(names)[:40])
  vals = [v for _, n, v in records if n == 'train/model_loss']
  assert np.isfinite(vals).all()
  vid = [v for _, n, v in records if n == 'report/openl_image'][-1]
  assert np.asarray(vid).shape == (6, 3 * 64, 2 * 64, 3)          # [T, 3H, B*W, C]
  # the checkpoint the loop wrote round-trips into a fresh agent (embodied.Checkpoint pickles save())
  other = Agent(env.obs_space, env.act_space, embodied.Counter(), config, _backend_for_tests=HostSim())
  import pickle
  with open(pathlib.Path(str(tmp_path)) / 'checkpoint.pkl', 'rb') as f:     # what Checkpoint.load reads
    other.load(pickle.load(f)['agent'])
  p1, p2 = agent._main.get_params(), other._main.get_params()
  assert all(np.array_equal(p1[k], p2[k]) for k in p1)


def test_reference_learner_process_loop(embodied, tmp_path):
  """embodied/run/learning.py -- the learner PROCESS of the actor/learner split (north_star: "the
  learner process drops in unchanged") -- over a replay pre-filled by the reference's Driver with
  its RandomAgent: dataset(), train(), report() for train and eval replays, agent.pkl checkpoint."""
  from daydreamer_b200.agent import Agent
  from tests.hostsim import HostSim
  config = embodied.Config(Agent.configs['defaults'])
  config = config.update(Agent.configs['debug'])
  config = embodied.Flags(config).parse([
      '--task', 'dummy_discrete', '--logdir', str(tmp_path), '--run', 'learning',
      # key regexes left at the reference defaults ('.*'): the encoder MLP takes vector, step,
      # reward and is_terminal, the decoder reconstructs vector and image only
      '--rssm.deter', '32', '--encoder.cnn_depth', '4', '--decoder.cnn_depth', '4',
      '--batch_size', '2', '--replay_chunk', '6', '--imag_horizon', '3', '--env.length', '12',
      '--train.train_fill', '30', '--train.sync_every', '-1', '--train.eval_every', '3',
      '--train.eval_samples', '1', '--train.steps', '6'])
  args = embodied.Config(logdir=config.logdir, **config.train)
  step = embodied.Counter()
  records = []
  logger = embodied.Logger(step, [lambda summaries: records.extend(summaries)], multiplier=1)
  env = embodied.envs.load_env(config.task, mode='train', logdir=str(tmp_path), amount=1,
                               parallel='none', size=(64, 64), length=config.env.length)
  store = embodied.replay.Stats(embodied.replay.RAMStore(1000))
  replay = embodied.replay.FixedLength(store, config.replay_chunk, **config.replay_fixed)
  driver = embodied.Driver(env)
  driver.on_step(replay.add)
  driver(embodied.RandomAgent(env.act_space).policy, steps=40, episodes=1)
  agent = Agent(env.obs_space, env.act_space, step, config, _backend_for_tests=HostSim())
  assert list(agent.spec.obs_dims) == ['vector', 'step', 'reward', 'is_terminal']
  assert agent.spec.dec_dims == {'vector': 7} and list(agent.spec.cnn_keys) == ['image']
  env.close()
  embodied.run.learning(agent, replay, replay, logger, args)
  assert int(step) == 6
  names = {name for _, name, _ in records}
  for want in ('train/model_loss', 'train/actor_grad_norm', 'report/model_loss_mean', 'eval/model_loss_mean'):
    assert want in names, (want, sorted(names)[:60])
  assert (pathlib.Path(str(tmp_path)) / 'agent.pkl').exists()
