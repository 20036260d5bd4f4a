"""Regenerates tests/golden/oracle_steps.npz: outputs of the CPU oracle (oracle/dreamer_oracle.py)
on seeded inputs, for three small configurations.  The oracle itself is anchored by the
known-answer tests in tests/test_oracle.py (the TensorFlow reference cannot be imported here, so
these vectors pin the ORACLE against drift between rounds -- they are not reference outputs).

  python tests/golden/make_oracle_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import dreamer_oracle as O      # noqa: E402
from tests import util                        # noqa: E402

CASES = {
    'proprio_normal': (dict(), 3, 5),
    'onehot': (dict(discrete=True, act_dim=5, actent_target=0.4), 2, 4),
    'image_onehot': (dict(discrete=True, act_dim=5, cnn_keys={'image': (18, 18, 3), 'depth': (18, 18, 1)},
                          cnn_depth=4, enc_kernels=(4, 4), dec_kernels=(3, 4, 4)), 2, 4),
}
METRICS = ('model_loss', 'actor_loss', 'extr_critic_loss', 'model_grad_norm', 'actor_grad_norm',
           'extr_critic_grad_norm', 'wmkl_mean', 'post_ent_mean', 'prior_ent_mean', 'actent_mean',
           'extr_imag_return_mean', 'extr_score_mean')


def run(name):
  kw, B, L = CASES[name]
  spec = O.tiny_spec(**kw)
  learner = O.Learner(spec, O.init_params(spec, seed=1))
  out = {}
  state = None
  for it in range(2):
    batch = util.make_batch(spec, B, L, seed=10 + it)
    noise = util.make_noise(spec, B, L, seed=20 + it)
    _, state, mets, aux = learner.train_step(batch, noise, state, keep=True)
    out[f'{name}/step{it}/metrics'] = np.array([mets[k] for k in METRICS], np.float64)
    out[f'{name}/step{it}/post_idx'] = aux['post']['idx'].numpy().astype(np.int8)
    out[f'{name}/step{it}/traj_idx'] = aux['traj']['stoch'].argmax(-1).numpy().astype(np.int8)
    for g in ('model', 'critic', 'actor'):
      grads = aux[f'{g}_grads']
      out[f'{name}/step{it}/{g}_grad_sums'] = np.array(
          [float(v.double().sum()) for _, v in sorted(grads.items())], np.float64)
  out[f'{name}/param_sums'] = np.array(
      [float(v.double().sum()) for _, v in sorted(learner.params.items())], np.float64)
  return out


def build():
  out = {}
  for name in CASES:
    out.update(run(name))
  return out


if __name__ == '__main__':
  data = build()
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle_steps.npz')
  np.savez_compressed(path, **data)
  print('wrote', path, len(data), 'arrays', os.path.getsize(path), 'bytes')
