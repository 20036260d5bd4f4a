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
