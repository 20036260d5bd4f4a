"""Minimal nested-config helper for running the Agent without the `embodied`
package (bench.py, tests, smoke).  With the reference's learner the Agent gets a
real `embodied.Config` (embodied/core/config.py) instead; both are read through
dotted paths only."""
import copy
import re

from . import configs as _configs


def _flatten(d, prefix=''):
  out = {}
  for k, v in d.items():
    if isinstance(v, dict):
      out.update(_flatten(v, f'{prefix}{k}.'))
    else:
      out[f'{prefix}{k}'] = v
  return out


def _unflatten(flat):
  out = {}
  for k, v in flat.items():
    node = out
    parts = k.split('.')
    for p in parts[:-1]:
      node = node.setdefault(p, {})
    node[parts[-1]] = v
  return out


def make_config(*names, **overrides):
  """defaults -> named blocks -> overrides; keys may be dotted paths or regexes
  over dotted paths (like embodied.Config.update)."""
  flat = _flatten(copy.deepcopy(_configs.DEFAULTS))
  updates = []
  for n in names:
    updates.append(_flatten(_configs.CONFIGS[n]))
  updates.append(_flatten(overrides))
  for upd in updates:
    for key, value in upd.items():
      if key in flat:
        keys = [key]
      else:
        pat = re.compile(key)
        keys = [k for k in flat if pat.fullmatch(k) or pat.match(k) and pat.match(k).end() == len(k)]
        if not keys:
          raise KeyError(f'unknown config key or pattern: {key}')
      for k in keys:
        old = flat[k]
        flat[k] = type(old)(value) if isinstance(old, (int, float, str, bool)) and not isinstance(value, type(old)) else value
  return _unflatten(flat)
