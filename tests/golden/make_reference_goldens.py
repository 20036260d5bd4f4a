"""NOT YET RUN (executing the reference checkout needs an authorisation the build session did not have;
no fixture from this script is committed and no test depends on one).

Would generate tests/golden/reference_shim_steps.npz: the REFERENCE's own learner code
(/root/reference/embodied/agents/dreamerv2plus: agent.py, nets.py, tfutils.py, behaviors.py) executing
`Agent.train` on a tiny a1-like configuration, with TensorFlow replaced by the torch-backed stand-in of
tests/golden/tf_shim.py (TensorFlow itself cannot be installed in this image).

What runs is the reference's Python: its wiring of encoder / RSSM / heads, its losses, AutoAdapt and
Normalize controllers, its Adam, its imagination rollout, lambda-returns and actor / critic updates.
What is substituted is the primitive layer (matmul, exp, reductions; torch autograd for GradientTape).
The fixture holds the inputs (parameters, two replay batches, every random draw in call order) and the
reference's outputs (metrics, posterior, gradients, updated parameters of two consecutive train steps);
tests/test_reference_goldens.py holds the oracle (oracle/dreamer_oracle.py) to them.

Run here (needs /root/reference):   python tests/golden/make_reference_goldens.py
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = pathlib.Path('/root/reference')

from tests.golden import tf_shim  # noqa: E402

B, L, STEPS = 3, 5, 2
OUT = pathlib.Path(__file__).with_name('reference_shim_steps.npz')


def reference_agent():
  tf = tf_shim.install()
  sys.path.insert(0, str(REF))
  import embodied
  agent_dir = REF / 'embodied' / 'agents' / 'dreamerv2plus'
  with tf_shim.argv0(agent_dir / 'train.py'):
    from embodied.agents.dreamerv2plus import tfagent
    tfagent.Wrapper = lambda cls: cls              # the bare Agent, not the tf.distribute wrapper
    from embodied.agents.dreamerv2plus import agent as refagent
  config = embodied.Config(refagent.Agent.configs['defaults'])
  config = config.update({                         # the a1 block's settings at toy widths
      'rssm.deter': 24, 'rssm.units': 20, 'rssm.stoch': 4, 'rssm.classes': 6,
      'encoder.mlp_keys': 'vector', 'encoder.cnn_keys': '$^', 'encoder.mlp_layers': 2, 'encoder.mlp_units': 16,
      'decoder.mlp_keys': 'vector', 'decoder.cnn_keys': '$^', 'decoder.mlp_layers': 2, 'decoder.mlp_units': 16,
      'reward_head.layers': 2, 'reward_head.units': 16, 'cont_head.layers': 2, 'cont_head.units': 16,
      'actor.layers': 2, 'actor.units': 16, 'critic.layers': 2, 'critic.units': 16,
      'actor.minstd': 0.1, 'actent.target': 0.5, 'discount': 0.995, 'imag_horizon': 3,
      'tf.jit': False, 'tf.precision': 'float32', 'tf.platform': 'cpu'})
  obs_space = {'vector': embodied.Space(np.float32, (5,)), 'reward': embodied.Space(np.float32),
               'is_first': embodied.Space(bool), 'is_last': embodied.Space(bool),
               'is_terminal': embodied.Space(bool)}
  act_space = {'action': embodied.Space(np.float32, (3,), -1, 1), 'reset': embodied.Space(bool)}
  agent = refagent.Agent(obs_space, act_space, embodied.Counter(), config)
  return tf, agent


def make_batch(seed):
  rng = np.random.default_rng(seed)
  first = np.zeros((B, L), bool)
  first[:, 0] = True
  first[1, 3] = True                                # an episode boundary inside a sequence
  term = np.zeros((B, L), bool)
  term[2, 2] = True
  return {'vector': rng.standard_normal((B, L, 5)).astype(np.float32),
          'reward': rng.standard_normal((B, L)).astype(np.float32),
          'is_first': first, 'is_last': term.copy(), 'is_terminal': term,
          'action': rng.uniform(-1, 1, (B, L, 3)).astype(np.float32)}


def named_variables(tf, agent):
  """oracle parameter name -> reference variable, by the reference's own module structure."""
  out = {}

  def walk(mod, prefix):
    for key, val in mod._modules.items():
      key = 'dist_out' if key.startswith('dist_') else key
      if isinstance(val, tf.Variable):
        out[f'{prefix}/{key}'] = val
      else:
        walk(val, f'{prefix}/{key}')
  wm, ac = agent.wm, agent.task_behavior.ac
  walk(wm.encoder._mlp, 'enc')
  walk(wm.rssm, 'rssm')
  walk(wm.heads['decoder']._mlp, 'dec')
  walk(wm.heads['reward'], 'reward')
  walk(wm.heads['cont'], 'cont')
  walk(ac.actor, 'actor')
  walk(ac.critics['extr'].net, 'critic')
  walk(ac.critics['extr'].target_net, 'slow_critic')
  return out


def reset_state(tf, agent):
  """Everything that is not a parameter back to its initial value (after the variable-creating run)."""
  wm, ac = agent.wm, agent.task_behavior.ac
  critic = ac.critics['extr']
  for opt in (wm.model_opt, ac.opt, critic.opt):
    opt._step.assign(0)
    for v in list(opt._m.values()) + list(opt._v.values()):
      v.assign(tf.zeros_like(v))
  wm.wmkl._scale.assign(tf.ones_like(wm.wmkl._scale))
  ac.actent._scale.assign(tf.ones_like(ac.actent._scale))
  for norm in [ac.advnorm, *ac.retnorms.values(), *ac.scorenorms.values()]:
    norm._mean.assign(0.0)
    norm._sqrs.assign(0.0)
    norm._step.assign(0)
  critic.updates.assign(-1)


def main():
  from oracle import dreamer_oracle as O
  tf, agent = reference_agent()
  to_tf = lambda batch: {k: tf.convert_to_tensor(v) for k, v in batch.items()}
  tf_shim.NOISE.reset(0)
  agent.train(to_tf(make_batch(99)), None)          # creates every variable (values are replaced below)
  refvars = named_variables(tf, agent)

  spec = O.tiny_spec()                               # the same toy widths, a1 hyper-parameters
  params = O.init_params(spec, seed=3)
  rng = np.random.default_rng(4)
  for k, v in params.items():                        # non-trivial LayerNorm parameters and biases
    if not k.endswith('/kernel'):
      params[k] = (v + 0.1 * rng.standard_normal(v.shape)).astype(np.float32)
  assert set(refvars) == set(params), (sorted(set(refvars) ^ set(params)))
  for k, var in refvars.items():
    assert tuple(var.shape) == params[k].shape, (k, tuple(var.shape), params[k].shape)
    var.assign(params[k])
  reset_state(tf, agent)

  # record every gradient the reference's optimizers ask for
  byid = {id(v): k for k, v in refvars.items()}
  grad_log = []
  plain_gradient = tf.GradientTape.gradient

  def gradient(self, target, sources):
    grads = plain_gradient(self, target, sources)
    grad_log.append({byid[id(s)]: g.numpy().copy() for s, g in zip(sources, grads) if id(s) in byid})
    return grads
  tf.GradientTape.gradient = gradient

  fix = {f'param0/{k}': v for k, v in params.items()}
  state = None
  for step in range(STEPS):
    batch = make_batch(10 + step)
    tf_shim.NOISE.reset(20 + step)
    grad_log.clear()
    outs, state, metrics = agent.train(to_tf(batch), state)
    draws = list(tf_shim.NOISE.draws)
    # the draws in call order: per observe step (prior sample, posterior sample); the actor's first
    # action; per imagined step (prior sample, action)
    S, C, A, H, N = spec.stoch, spec.classes, spec.act_dim, spec.horizon, B * L
    assert len(draws) == 2 * L + 1 + 2 * H, len(draws)
    obs_post = np.stack([draws[2 * t + 1][1].reshape(B, S, C) for t in range(L)])
    rest = draws[2 * L:]
    act = np.stack([rest[0][1]] + [rest[2 + 2 * h][1] for h in range(H)])
    img = np.stack([rest[1 + 2 * h][1].reshape(N, S, C) for h in range(H)])
    assert act.shape == (H + 1, N, A) and img.shape == (H, N, S, C)
    for k, v in batch.items():
      fix[f'step{step}/batch/{k}'] = v
    fix[f'step{step}/noise/obs_post'], fix[f'step{step}/noise/img'], fix[f'step{step}/noise/act'] = obs_post, img, act
    for k, v in metrics.items():
      fix[f'step{step}/metric/{k}'] = np.float64(v.numpy())
    for k in ('deter', 'stoch', 'logit'):
      fix[f'step{step}/state/{k}'] = state[k].numpy().copy()
    assert len(grad_log) == 3, len(grad_log)        # model, critic, actor
    for group, g in zip(('model', 'critic', 'actor'), grad_log):
      for k, v in g.items():
        fix[f'step{step}/grad/{group}/{k}'] = v
    for k, var in refvars.items():
      fix[f'step{step}/param/{k}'] = var.numpy().copy()
  np.savez_compressed(OUT, **fix)
  print(f'wrote {OUT} ({OUT.stat().st_size / 1024:.0f} KB, {len(fix)} arrays)')

  # the comparison the test makes, for a first look
  ol = O.Learner(spec, {k: v.copy() for k, v in params.items()})
  ostate = None
  for step in range(STEPS):
    batch = {k[len(f'step{step}/batch/'):]: v for k, v in fix.items() if k.startswith(f'step{step}/batch/')}
    noise = {k[len(f'step{step}/noise/'):]: v for k, v in fix.items() if k.startswith(f'step{step}/noise/')}
    _, ostate, omets, aux = ol.train_step(batch, noise, ostate, keep=True)
    worst = max(((abs(float(omets[k]) - float(fix[f'step{step}/metric/{k}'])) /
                  (abs(float(fix[f'step{step}/metric/{k}'])) + 1e-6), k)
                 for k in omets if f'step{step}/metric/{k}' in fix and np.isfinite(fix[f'step{step}/metric/{k}'])))
    pdiff = max(float(np.abs(ol.params[k].numpy() - fix[f'step{step}/param/{k}']).max()) for k in params)
    print(f'step {step}: worst metric rel diff {worst[0]:.2e} ({worst[1]}), max |param diff| {pdiff:.2e} '
          f'(lr = {spec.lr})')


if __name__ == '__main__':
  main()
