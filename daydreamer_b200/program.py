"""Layer programs for the layer machine (csrc/dd_machine.cu, include/dd_kernels.h `DdLayer`).

A program is a list of `Layer`s sharing one row count M; every 128-row tile of the batch runs
through all of them inside ONE persistent kernel (one thread-block cluster per tile).  A layer is

    acc[M, N] = sum_seg  A_seg[M, k_seg] @ B[wn0:wn0+N, wk0_seg:wk0_seg+k_seg].T
    out       = epilogue(acc (+ add) (+ bias))

with epilogue PLAIN, LN (LayerNorm forward + optional elu; A/nets.py:585-602) or LNBWD (the
reverse of LN).  Every GEMM operand is a pair (x, x_lo) -- see `lo_of` -- so the tensor core
reaches fp32 accuracy with three tf32 products; an A segment with `a_lo=None` is exact in tf32
(the one-hot latent) and skips its lo terms.

This module only DESCRIBES programs with torch views; backends compile and run them
(`be.program_build(layers, rows)` / `be.program_run(handle)`).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import torch

EPI_PLAIN, EPI_LN, EPI_LNBWD, EPI_GRU, EPI_STATS, EPI_ACTOR, EPI_STATSBWD, EPI_GRUBWD = range(8)


@dataclasses.dataclass
class Seg:
  a: torch.Tensor                      # [M, k] view (row stride a multiple of 4 floats)
  a_lo: Optional[torch.Tensor]         # same layout, or None for an exact (one-hot) operand
  wk0: int = 0                         # first column of the B operand this segment multiplies


@dataclasses.dataclass
class Layer:
  segs: List[Seg]                      # empty: epilogue-only layer, its input is `add`
  w: Optional[torch.Tensor]            # B operand: row-major [wrows, wcols] view
  w_lo: Optional[torch.Tensor]
  N: int
  epi: int = EPI_PLAIN
  act: int = 0
  wn0: int = 0
  add: Optional[torch.Tensor] = None   # [M, N] added to the accumulator
  bias: Optional[torch.Tensor] = None  # [N]
  gamma: Optional[torch.Tensor] = None
  beta: Optional[torch.Tensor] = None
  z: Optional[torch.Tensor] = None     # LN: pre-norm output; LNBWD: saved pre-norm input
  st: Optional[torch.Tensor] = None    # [M, 2] (mean, rstd)
  out: Optional[torch.Tensor] = None
  out_lo: Optional[torch.Tensor] = None
  dgamma: Optional[torch.Tensor] = None
  dbeta: Optional[torch.Tensor] = None
  # epilogue-specific extras -- GRU: x0 = previous deter [M, D], x1 = post-norm gates [M, 3D] (saved);
  # STATS: x0 = Gumbel noise (None: logits only), x1 = one-hot sample, f0 = unimix, ipar = classes;
  # ACTOR: x0 = eps, x1 = mean, x2 = std, f0 / f1 = min / max std, ipar = action dim;
  # STATSBWD: x0 = forward logits, f0 = unimix, ipar = classes; GRUBWD: x0 = previous deter, x1 = y,
  # x2 = buffer that accumulates d(previous deter), z / st / gamma of the forward GRU layer
  x0: Optional[torch.Tensor] = None
  x1: Optional[torch.Tensor] = None
  x2: Optional[torch.Tensor] = None
  f0: float = 0.0
  f1: float = 0.0
  ipar: int = 0


def eligible_width(n: int, epi: int = EPI_LN) -> bool:
  """Column counts the machine tiles across its 8 CTAs (dd_program_build)."""
  return n % 256 == 0 or (epi in (EPI_PLAIN, EPI_STATS, EPI_ACTOR) and 1 <= n <= 64)


def trunc19(x: torch.Tensor) -> torch.Tensor:
  """What the tf32 tensor core reads of an fp32 word: sign, exponent, top 10 mantissa bits."""
  return (x.contiguous().view(torch.int32) & -8192).view(torch.float32)


def lo_of(x: torch.Tensor) -> torch.Tensor:
  """tf32 remainder of x (csrc/dd_machine.cu lo_of): round-to-nearest tf32 of x - trunc19(x)."""
  x = x.contiguous()
  lo = x - trunc19(x)
  return ((lo.view(torch.int32) + 4096) & -8192).view(torch.float32)
