"""CPU: the torch statement of layer programs (tests/hostsim.py program_run) against plain
autograd, and the (x, x_lo) operand split the tensor core relies on."""
import numpy as np
import torch
import torch.nn.functional as F

from daydreamer_b200 import program as P
from tests.hostsim import HostSim


def rnd(*shape, seed=0, scale=1.0):
  g = torch.Generator().manual_seed(seed)
  return torch.randn(*shape, generator=g) * scale


def test_lo_split_is_tf32_exact_and_accurate():
  x = rnd(4096, seed=1) * torch.exp(rnd(4096, seed=2) * 3)
  hi, lo = P.trunc19(x), P.lo_of(x)
  for t in (hi, lo):           # both halves survive the tensor core's 19-bit read unchanged
    assert torch.equal(P.trunc19(t), t)
  rel = ((hi.double() + lo.double() - x.double()).abs() / x.double().abs()).max().item()
  assert rel <= 2.0 ** -21, rel
  oh = F.one_hot(torch.arange(8) % 4, 4).float()
  assert torch.equal(P.trunc19(oh), oh) and float(P.lo_of(oh).abs().max()) == 0.0


def test_trunk_program_matches_autograd():
  """4-layer Linear->LayerNorm->elu trunk forward + the reverse chain as two programs."""
  be = HostSim()
  M, D, U = 96, 40, 256
  X = rnd(M, D, seed=1)
  Ws = [rnd(D if i == 0 else U, U, seed=10 + i, scale=0.1).requires_grad_() for i in range(3)]
  gam = [(1.0 + 0.1 * rnd(U, seed=30 + i)).requires_grad_() for i in range(3)]
  bet = [(0.1 * rnd(U, seed=40 + i)).requires_grad_() for i in range(3)]
  Xg = X.clone().requires_grad_()
  h = Xg
  for i in range(3):
    h = F.elu(F.layer_norm(h @ Ws[i], (U,), gam[i], bet[i], eps=1e-3))
  dH = rnd(M, U, seed=50)
  h.backward(dH)
  # programs
  Z = [torch.zeros(M, U) for _ in range(3)]
  H = [torch.zeros(M, U) for _ in range(3)]
  Hlo = [torch.zeros(M, U) for _ in range(3)]
  st = [torch.zeros(M, 2) for _ in range(3)]
  layers = []
  for i in range(3):
    wt = Ws[i].detach().t().contiguous()
    src = P.Seg(X, P.lo_of(X)) if i == 0 else P.Seg(H[i - 1], Hlo[i - 1])
    layers.append(P.Layer([src], wt, P.lo_of(wt), U, P.EPI_LN, act=1, gamma=gam[i].detach(),
                          beta=bet[i].detach(), z=Z[i], st=st[i], out=H[i], out_lo=Hlo[i]))
  be.program_run(be.program_build(layers, M))
  np.testing.assert_allclose(H[2].numpy(), h.detach().numpy(), rtol=2e-5, atol=2e-5)
  dZ = [torch.zeros(M, U) for _ in range(3)]
  dZlo = [torch.zeros(M, U) for _ in range(3)]
  dg = [torch.zeros(U) for _ in range(3)]
  db = [torch.zeros(U) for _ in range(3)]
  dX = torch.zeros(M, D)
  back = [P.Layer([], None, None, U, P.EPI_LNBWD, act=1, add=dH,
                  gamma=gam[2].detach(), beta=bet[2].detach(), z=Z[2], st=st[2], out=dZ[2],
                  out_lo=dZlo[2], dgamma=dg[2], dbeta=db[2])]
  for i in (1, 0):
    w = Ws[i + 1].detach()
    back.append(P.Layer([P.Seg(dZ[i + 1], dZlo[i + 1])], w, P.lo_of(w), U, P.EPI_LNBWD, act=1,
                        gamma=gam[i].detach(), beta=bet[i].detach(), z=Z[i], st=st[i], out=dZ[i],
                        out_lo=dZlo[i], dgamma=dg[i], dbeta=db[i]))
  w = Ws[0].detach()
  back.append(P.Layer([P.Seg(dZ[0], dZlo[0])], w, P.lo_of(w), D, P.EPI_PLAIN, out=dX))
  be.program_run(be.program_build(back, M))
  np.testing.assert_allclose(dX.numpy(), Xg.grad.numpy(), rtol=1e-4, atol=1e-5)
  for i in range(3):
    np.testing.assert_allclose(dg[i].numpy(), gam[i].grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(db[i].numpy(), bet[i].grad.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose((H[i - 1] if i else X).t().numpy() @ dZ[i].numpy(),
                               Ws[i].grad.numpy(), rtol=1e-3, atol=1e-4)


def test_stale_lo_is_detected():
  """The simulator reconstructs operands from (x, x_lo): a remainder that was not refreshed
  after x changed must change the result (this is what guards the engine's wiring on CPU)."""
  be = HostSim()
  M, K, N = 32, 64, 256
  X, W = rnd(M, K, seed=1), rnd(K, N, seed=2)
  wt = W.t().contiguous()
  out = torch.zeros(M, N)
  L = P.Layer([P.Seg(X, torch.zeros(M, K))], wt, P.lo_of(wt), N, P.EPI_PLAIN, out=out)
  be.program_run(be.program_build([L], M))
  err = (out - X @ W).abs().max().item()
  assert err > 1e-4, err


def test_transpose_jobs():
  be = HostSim()
  shapes = [(8, 12), (5, 3)]
  p = rnd(sum(r * c for r, c in shapes), seed=3)
  pt, ptlo = torch.zeros(200), torch.zeros(200)
  be.transpose_lo(p, pt, ptlo, [(0, 8, 12, 0, 8, 0), (96, 5, 3, 100, 4, 1)])
  assert torch.equal(pt[:96].view(12, 8), p[:96].view(8, 12).t())
  got = pt[100:100 + 4 * 4 + 3].as_strided((5, 3), (4, 1))
  assert torch.equal(got, p[96:111].view(5, 3))
  assert torch.equal(ptlo[:96].view(12, 8), P.lo_of(p[:96].view(8, 12).t().contiguous()))


# ---- dd_program_deps: which earlier layers write an A segment (host-only entry point of the .so) ----
def _brute_force_deps(layers, rows):
  """The same question answered element by element: addresses (in floats) every layer writes, the
  last earlier layer whose set meets a segment's."""
  def cells(t, width=None):
    if t is None:
      return set()
    t2 = t if t.dim() == 2 else t.view(-1, 1) if t.dim() == 1 else t.reshape(-1, t.shape[-1])
    w = t2.shape[1] if width is None else width
    base, ld = t2.data_ptr() // 4, (t2.stride(0) if t2.shape[0] > 1 else w)
    return {base + r * ld + c for r in range(rows) for c in range(w)}
  written = []
  for L in layers:
    ow = L.N // 3 if L.epi == P.EPI_GRU else 3 * L.N if L.epi == P.EPI_GRUBWD else L.ipar if L.epi == P.EPI_ACTOR else L.N
    w = cells(L.out, ow) | cells(L.out_lo, ow)
    if L.epi in (P.EPI_LN, P.EPI_GRU, P.EPI_PLAIN):
      w |= cells(L.z, L.N if L.z is not None else None)
    if L.epi in (P.EPI_LN, P.EPI_GRU):
      w |= cells(L.st)
    if L.epi in (P.EPI_GRU, P.EPI_STATS, P.EPI_ACTOR):
      w |= cells(L.x1, L.ipar if L.epi == P.EPI_ACTOR else L.N)
    if L.epi in (P.EPI_ACTOR, P.EPI_GRUBWD):
      w |= cells(L.x2, L.ipar if L.epi == P.EPI_ACTOR else L.N)
    written.append(w)
  out = []
  for l, L in enumerate(layers):
    deps = [0, 0]
    for g, sg in enumerate(L.segs):
      need = cells(sg.a) | cells(sg.a_lo)
      for j in range(l - 1, -1, -1):
        if need & written[j]:
          deps[g] = j + 1
          break
    out.append(tuple(deps))
  return out


def test_segment_dependencies_of_the_rollout_programs():
  """The engine's real rollout / reverse programs (small widths, CPU tensors): dd_program_deps agrees
  with the element-wise answer, and the early operands really are early: the latent of img_in does not
  wait for the actor head before it, the carried deter of the GRU does not wait for img_in."""
  import os
  from daydreamer_b200 import backend, engine as E
  from oracle import dreamer_oracle as O
  from tests import util
  lib = backend.load_library()
  os.environ['DD_REVERSE'] = '1'
  ospec = O.tiny_spec(obs_dims={'vector': 16}, act_dim=12, deter=256, units=256, stoch=8, classes=32,
                      enc_units=256, head_units=256, horizon=2)
  eng = E.Engine(util.to_engine_spec(ospec), HostSim(), 2, 3)
  eng.load_params(O.init_params(ospec, seed=1))
  util.load_batch(eng, util.to_engine_spec(ospec), util.make_batch(ospec, 2, 3, seed=1), util.make_noise(ospec, 2, 3, seed=2))
  eng.launch_step(draw=False, slow_refreshed=True)
  fwd = eng._rollout_layers
  rows = eng.N
  got = backend.program_deps(lib, fwd, rows)
  assert got == _brute_force_deps(fwd, rows)
  rev = eng._reverse_layers
  assert backend.program_deps(lib, rev, rows) == _brute_force_deps(rev, rows)
  kinds = [L.epi for L in fwd]
  for l, L in enumerate(fwd):
    if len(L.segs) == 1:
      assert got[l][0] in (0, l)                 # a chain: the layer before (or a program input)
      continue
    d0, d1 = got[l]
    if L.epi == P.EPI_GRU:                       # [carried deter | x_in]: x_in is the layer before
      assert d1 == l and d0 < l
    elif l > 0 and kinds[l - 1] == P.EPI_ACTOR:  # img_in: [latent | action]: the action is the layer before
      assert d1 == l and d0 <= l - 2           # ... and the latent was sampled before the actor ran
    elif l == 0:                                 # first actor call: the start states are inputs
      assert (d0, d1) == (0, 0)
    else:                                        # actor layer 1: [deter (GRU, 5 back) | latent (STATS, before)]
      assert d1 == l and 0 < d0 < l
