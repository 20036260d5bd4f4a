"""NumPy Philox4x32-10 restatement  --  TEST INFRASTRUCTURE ONLY.

Mirrors the counter-based generator the CUDA kernels use for the categorical
(Gumbel) and action (normal) noise, so parity tests can materialise the very
same noise tensors on the host (SURVEY.md section 7 "Sampling parity": identical
noise injected -> identical indices).  Philox is Salmon et al., "Parallel
random numbers: as easy as 1, 2, 3" (SC'11); constants are the published ones.

Stream layout (shared with daydreamer_b200/csrc/dd_rng.cuh):
  element e of a buffer uses block = e // 4, lane = e % 4;
  counter = (block_lo, block_hi, stream_id, step), key = (seed_lo, seed_hi);
  u = ((x >> 8) + 0.5) * 2^-24 in (0, 1).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
  c0, c1, c2, c3 = (np.asarray(c, np.uint32).copy() for c in (c0, c1, c2, c3))
  k0 = np.uint32(k0)
  k1 = np.uint32(k1)
  with np.errstate(over='ignore'):
    for _ in range(rounds):
      p0 = M0 * c0.astype(np.uint64)
      p1 = M1 * c2.astype(np.uint64)
      hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
      lo0 = (p0 & MASK).astype(np.uint32)
      hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
      lo1 = (p1 & MASK).astype(np.uint32)
      c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
      k0 = np.uint32(k0 + W0)
      k1 = np.uint32(k1 + W1)
  return c0, c1, c2, c3


def _blocks(n, seed, stream, step):
  nb = (n + 3) // 4
  blk = np.arange(nb, dtype=np.uint64)
  c0 = (blk & MASK).astype(np.uint32)
  c1 = (blk >> np.uint64(32)).astype(np.uint32)
  c2 = np.full(nb, stream, np.uint32)
  c3 = np.full(nb, step & 0xFFFFFFFF, np.uint32)
  return philox4x32(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def _u01(x):
  return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def uniform(n, seed, stream, step):
  x = np.stack(_blocks(n, seed, stream, step), 1).reshape(-1)[:n]
  return _u01(x)


def gumbel(n, seed, stream, step):
  u = uniform(n, seed, stream, step)
  return (-np.log(-np.log(u))).astype(np.float32)


def normal(n, seed, stream, step):
  x0, x1, x2, x3 = _blocks(n, seed, stream, step)
  u0, u1, u2, u3 = _u01(x0), _u01(x1), _u01(x2), _u01(x3)
  two_pi = np.float32(2.0 * np.pi)
  r0 = np.sqrt(np.float32(-2.0) * np.log(u0))
  r1 = np.sqrt(np.float32(-2.0) * np.log(u2))
  z = np.stack([r0 * np.cos(two_pi * u1), r0 * np.sin(two_pi * u1),
                r1 * np.cos(two_pi * u3), r1 * np.sin(two_pi * u3)], 1)
  return z.reshape(-1)[:n].astype(np.float32)
