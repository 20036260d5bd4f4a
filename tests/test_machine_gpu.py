"""-m gpu: the layer machine (csrc/dd_machine.cu) through the C-ABI against the torch-CPU statement
of the same programs (tests/hostsim.py program_run): single layers of every epilogue, segment /
chunk / tile edge cases, multi-layer chains, and fp32-level accuracy of the TMA-fed 3xTF32 product
against float64."""
import numpy as np
import pytest
import torch

from daydreamer_b200 import program as P
from tests.hostsim import HostSim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', params=[1, 2], ids=['cluster8', 'splitk16'])
def backends(request):
  """Both geometries of the machine: clusters of 8 CTAs (default) and split-K clusters of 16."""
  from daydreamer_b200.backend import CudaBackend
  cu = CudaBackend()
  cu.lib.dd_set_machine_ks(request.param)
  yield HostSim(), cu
  cu.lib.dd_set_machine_ks(1)


def rnd(*shape, seed=0, scale=1.0):
  g = torch.Generator().manual_seed(seed)
  return torch.randn(*shape, generator=g) * scale


def run_both(backends, case, rtol=2e-5, atol=2e-5):
  hs, cu = backends
  ref = case(hs, lambda x: x.clone())
  try:
    out = case(cu, lambda x: x.clone().to(cu.device))
    cu.sync()
  except Exception:
    print('machine debug:', cu.program_debug())
    raise
  for k in ref:
    if k == 'lo':      # a remainder is only defined relative to the value it belongs to
      for res in (ref, out):
        val = next((res[n] for n in ('out', 'dz', 'x') if n in res), None)
        assert val is not None and torch.equal(P.lo_of(val.cpu()).view(val.shape), res['lo'].cpu()), \
            'x_lo is not the tf32 remainder of x'
      continue
    a, b = ref[k].double().cpu().numpy(), out[k].double().cpu().numpy()
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(a)
    assert (err <= tol).all(), (k, float(err.max()), float(np.abs(a).max()),
                                int((err > tol).sum()), a.shape)
  return ref, out


def wt_pair(T, W):
  """Forward B operand of x @ W: the transposed kernel and its remainder."""
  wt = torch.empty(W.shape[1], W.shape[0]).copy_(W.t())
  pair = T(torch.stack([wt, P.lo_of(wt).view(wt.shape)]))     # two planes: one 3-D TMA box per tile
  return pair[0], pair[1]


def act_pair(T, X):
  return T(X), T(P.lo_of(X))


@pytest.mark.parametrize('M,N,K', [(1024, 512, 512), (128, 512, 64), (1000, 256, 100), (16384, 512, 512),
                                   (300, 1024, 36), (1024, 24, 512), (1024, 1, 512), (777, 64, 512)])
def test_plain_layer(backends, M, N, K):
  def case(be, T):
    X, Xlo = act_pair(T, rnd(M, K, seed=1))
    Wt, Wtlo = wt_pair(T, rnd(K, N, seed=2, scale=K ** -0.5))
    bias = T(rnd(N, seed=3))
    add = T(rnd(M, N, seed=4))
    out, out_lo = T(torch.zeros(M, N)), T(torch.zeros(M, N))
    L = P.Layer([P.Seg(X, Xlo)], Wt, Wtlo, N, P.EPI_PLAIN, bias=bias, add=add, out=out,
                out_lo=out_lo if N % 4 == 0 else None)
    be.program_run(be.program_build([L], M))
    res = {'out': out}
    if N % 4 == 0:
      res['lo'] = out_lo
    return res
  run_both(backends, case, rtol=2e-5, atol=2e-5)


def test_accuracy_vs_float64(backends):
  """hi = the raw fp32 word (the tensor core reads its top 19 bits), lo = tf32 remainder: the
  three-product sum must be fp32-accurate, far beyond plain tf32 (1e-3)."""
  _, cu = backends
  M, N, K = 1024, 512, 1536
  X, W = rnd(M, K, seed=5), rnd(K, N, seed=6, scale=K ** -0.5)
  T = lambda x: x.clone().to(cu.device)
  Xd, Xlo = act_pair(T, X)
  Wt, Wtlo = wt_pair(T, W)
  out = torch.zeros(M, N, device=cu.device)
  L = P.Layer([P.Seg(Xd, Xlo)], Wt, Wtlo, N, P.EPI_PLAIN, out=out)
  cu.program_run(cu.program_build([L], M))
  cu.sync()
  ref = X.double() @ W.double()
  err = (out.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
  print(f'3xTF32 via TMA: max err / max|C| = {err:.3e} at K={K}')
  assert err < 5e-6, err


@pytest.mark.parametrize('M,N,K1,K2,act', [(1024, 512, 512, 0, 1), (1024, 1536, 512, 512, 0),
                                           (1000, 1024, 512, 12, 1), (256, 256, 96, 0, 1),
                                           (4096, 512, 1024, 512, 1)])
def test_ln_layer(backends, M, N, K1, K2, act):
  """LayerNorm(+elu) epilogue, one or two A segments (the second exact one-hot when K2 = 1024-like),
  one to three column chunks per CTA."""
  def case(be, T):
    K = K1 + K2
    X1, X1lo = act_pair(T, rnd(M, K1, seed=1))
    segs = [P.Seg(X1, X1lo, 0)]
    if K2:
      X2, X2lo = act_pair(T, rnd(M, K2, seed=2))
      segs.append(P.Seg(X2, X2lo, K1))
    Wt, Wtlo = wt_pair(T, rnd(K, N, seed=3, scale=K ** -0.5))
    gamma, beta = T(1.0 + 0.1 * rnd(N, seed=4)), T(0.1 * rnd(N, seed=5))
    z, out, out_lo, st = (T(torch.zeros(M, N)), T(torch.zeros(M, N)), T(torch.zeros(M, N)),
                          T(torch.zeros(M, 2)))
    L = P.Layer(segs, Wt, Wtlo, N, P.EPI_LN, act=act, gamma=gamma, beta=beta, z=z, st=st, out=out,
                out_lo=out_lo)
    be.program_run(be.program_build([L], M))
    return {'z': z, 'out': out, 'lo': out_lo, 'st': st}
  run_both(backends, case, rtol=5e-5, atol=5e-5)


def test_onehot_segment(backends):
  """An exact (one-hot) operand skips its lo products: [deter | stoch] @ W as two segments."""
  def case(be, T):
    M, D, S, C, N = 1024, 512, 32, 32, 512
    X1, X1lo = act_pair(T, rnd(M, D, seed=1))
    g = torch.Generator().manual_seed(7)
    idx = torch.randint(0, C, (M, S), generator=g)
    oh = T(torch.nn.functional.one_hot(idx, C).float().reshape(M, S * C))
    Wt, Wtlo = wt_pair(T, rnd(D + S * C, N, seed=3, scale=0.03))
    gamma, beta = T(1.0 + 0.1 * rnd(N, seed=4)), T(0.1 * rnd(N, seed=5))
    out, st = T(torch.zeros(M, N)), T(torch.zeros(M, 2))
    L = P.Layer([P.Seg(X1, X1lo, 0), P.Seg(oh, None, D)], Wt, Wtlo, N, P.EPI_LN, act=1,
                gamma=gamma, beta=beta, st=st, out=out)
    be.program_run(be.program_build([L], M))
    return {'out': out, 'st': st}
  run_both(backends, case, rtol=5e-5, atol=5e-5)


@pytest.mark.parametrize('M,N,K,act,wgrad', [(1024, 512, 512, 1, True), (1024, 1536, 512, 0, False),
                                             (900, 512, 1024, 1, True), (2048, 256, 64, 1, True)])
def test_ln_bwd_layer(backends, M, N, K, act, wgrad):
  """dZ of h = act(LN(z)) from dH = dY @ W^T computed by the same layer; dgamma / dbeta."""
  def case(be, T):
    dY, dYlo = act_pair(T, rnd(M, K, seed=1))
    W = rnd(N, K, seed=2, scale=K ** -0.5)              # kernel [N_in = N, N_out = K]: dH = dY @ W^T
    Wd, Wlo = T(W), T(P.lo_of(W))                      # separate allocations: the 2-descriptor path
    z = T(rnd(M, N, seed=3))
    st = T(torch.zeros(M, 2))
    gamma, beta = T(1.0 + 0.1 * rnd(N, seed=4)), T(0.1 * rnd(N, seed=5))
    HostSim().ln_fwd(z.cpu(), gamma.cpu(), beta.cpu(), torch.zeros(M, N), (stc := torch.zeros(M, 2)), act)
    st.copy_(stc)
    add = T(rnd(M, N, seed=6))
    dz, dz_lo = T(torch.zeros(M, N)), T(torch.zeros(M, N))
    dg, db = T(rnd(N, seed=7)), T(rnd(N, seed=8))
    L = P.Layer([P.Seg(dY, dYlo)], Wd, Wlo, N, P.EPI_LNBWD, act=act, gamma=gamma, beta=beta, z=z,
                st=st, add=add, out=dz, out_lo=dz_lo, dgamma=dg if wgrad else None,
                dbeta=db if wgrad else None)
    be.program_run(be.program_build([L], M))
    return {'dz': dz, 'lo': dz_lo, 'dg': dg, 'db': db}
  run_both(backends, case, rtol=1e-4, atol=1e-4 * np.sqrt(M / 128))


@pytest.mark.parametrize('M', [1024, 16384, 1600, 4800])    # 16384 / 4800 rows: two interleaved tiles per cluster
def test_trunk_chain(backends, M):
  """A 4-layer MLP trunk + its output layer as ONE program: layer i+1 reads what layer i wrote
  (x and x_lo) after a cluster-local handshake; then the reverse chain as a second program."""
  def case(be, T):
    D, U, NOUT = 1536, 512, 24
    X, Xlo = act_pair(T, rnd(M, D, seed=1))
    Ws = [rnd(D if i == 0 else U, U, seed=10 + i, scale=(D if i == 0 else U) ** -0.5) for i in range(4)]
    Wo = rnd(U, NOUT, seed=20, scale=U ** -0.5)
    gam = [T(1.0 + 0.1 * rnd(U, seed=30 + i)) for i in range(4)]
    bet = [T(0.1 * rnd(U, seed=40 + i)) for i in range(4)]
    Z = [T(torch.zeros(M, U)) for _ in range(4)]
    H = [T(torch.zeros(M, U)) for _ in range(4)]
    Hlo = [T(torch.zeros(M, U)) for _ in range(4)]
    st = [T(torch.zeros(M, 2)) for _ in range(4)]
    out = T(torch.zeros(M, NOUT))
    layers = []
    for i in range(4):
      Wt, Wtlo = wt_pair(T, Ws[i])
      src = P.Seg(X, Xlo) if i == 0 else P.Seg(H[i - 1], Hlo[i - 1])
      layers.append(P.Layer([src], Wt, Wtlo, U, P.EPI_LN, act=1, gamma=gam[i], beta=bet[i], z=Z[i],
                            st=st[i], out=H[i], out_lo=Hlo[i]))
    Wt, Wtlo = wt_pair(T, Wo)
    layers.append(P.Layer([P.Seg(H[3], Hlo[3])], Wt, Wtlo, NOUT, P.EPI_PLAIN, bias=T(rnd(NOUT, seed=50)),
                          out=out))
    be.program_run(be.program_build(layers, M))
    # reverse: dH3 given -> dZ3 -> dH2 = dZ3 @ W3^T -> ... -> dX = dZ0 @ W0^T
    dH, dHlo = act_pair(T, rnd(M, U, seed=60))
    Wb = [(T(w), T(P.lo_of(w))) for w in Ws]
    dZ = [T(torch.zeros(M, U)) for _ in range(4)]
    dZlo = [T(torch.zeros(M, U)) for _ in range(4)]
    dg = [T(torch.zeros(U)) for _ in range(4)]
    db = [T(torch.zeros(U)) for _ in range(4)]
    dX = T(rnd(M, D, seed=61))
    back = [P.Layer([], None, None, U, P.EPI_LNBWD, add=dH, act=1, gamma=gam[3], beta=bet[3],
                    z=Z[3], st=st[3], out=dZ[3], out_lo=dZlo[3], dgamma=dg[3], dbeta=db[3])]
    for i in (2, 1, 0):
      back.append(P.Layer([P.Seg(dZ[i + 1], dZlo[i + 1])], Wb[i + 1][0], Wb[i + 1][1], U, P.EPI_LNBWD,
                          act=1, gamma=gam[i], beta=bet[i], z=Z[i], st=st[i], out=dZ[i],
                          out_lo=dZlo[i], dgamma=dg[i], dbeta=db[i]))
    back.append(P.Layer([P.Seg(dZ[0], dZlo[0])], Wb[0][0], Wb[0][1], D, P.EPI_PLAIN, add=dX, out=dX))
    be.program_run(be.program_build(back, M))
    return {'h3': H[3], 'out': out, 'st3': st[3], 'dz0': dZ[0], 'dx': dX, 'dg0': dg[0], 'db3': db[3]}
  run_both(backends, case, rtol=2e-4, atol=2e-4 * np.sqrt(M / 128))


def test_operand_prep(backends):
  """dd_split_lo on a strided view and dd_transpose_lo over several matrices of a flat buffer."""
  def case(be, T):
    X = T(rnd(333, 72, seed=1))
    lo = T(torch.zeros(333, 64))
    be.split_lo(X[:, 4:68], lo)
    shapes = [(512, 512), (1036, 512), (512, 12), (40, 24)]
    n = sum(r * c for r, c in shapes)
    p = T(rnd(n, seed=2))
    pt, ptlo = T(torch.zeros(n)), T(torch.zeros(n))
    jobs, off = [], 0
    for r, c in shapes:
      jobs.append((off, r, c, off, r if len(jobs) % 2 == 0 else c, len(jobs) % 2))
      off += r * c
    be.transpose_lo(p, pt, ptlo, jobs)
    return {'x': X[:, 4:68].contiguous(), 'lo': lo, 'pt': pt, 'ptlo': ptlo}
  ref, out = run_both(backends, case, rtol=0, atol=0)


@pytest.mark.parametrize('M,D', [(1024, 512), (300, 256), (4800, 256)])
def test_gru_layer(backends, M, D):
  """[deter | x] @ W_gru -> joint LayerNorm over 3 D -> gates -> new deter, as one layer."""
  def case(be, T):
    X1, X1lo = act_pair(T, rnd(M, D, seed=1))
    X2, X2lo = act_pair(T, rnd(M, D, seed=2))
    Wt, Wtlo = wt_pair(T, rnd(2 * D, 3 * D, seed=3, scale=(2 * D) ** -0.5))
    gamma, beta = T(1.0 + 0.1 * rnd(3 * D, seed=4)), T(0.1 * rnd(3 * D, seed=5))
    z, y, st = T(torch.zeros(M, 3 * D)), T(torch.zeros(M, 3 * D)), T(torch.zeros(M, 2))
    feat = T(torch.zeros(M, D + 64))
    out, out_lo = feat[:, :D], T(torch.zeros(M, D))
    L = P.Layer([P.Seg(X1, X1lo, 0), P.Seg(X2, X2lo, D)], Wt, Wtlo, 3 * D, P.EPI_GRU, gamma=gamma,
                beta=beta, z=z, st=st, out=out, out_lo=out_lo, x0=X1, x1=y)
    be.program_run(be.program_build([L], M))
    return {'z': z, 'y': y, 'st': st, 'out': feat, 'lo': out_lo}
  def fix(res):     # 'lo' belongs to the strided view
    res['out'] = res['out'][:, :D].contiguous()
    return res
  hs, cu = backends
  ref = fix(case(hs, lambda x: x.clone()))
  out = fix(case(cu, lambda x: x.clone().to(cu.device)))
  cu.sync()
  for k in ('z', 'y', 'st', 'out'):
    np.testing.assert_allclose(out[k].cpu().numpy(), ref[k].numpy(), rtol=5e-5, atol=5e-5, err_msg=k)
  assert torch.equal(P.lo_of(out['out'].cpu()), out['lo'].cpu())


@pytest.mark.parametrize('M,K,S,C,noise', [(1024, 512, 32, 32, True), (1000, 256, 8, 32, False),
                                           (1024, 512, 1, 6, True)])
def test_stats_layer(backends, M, K, S, C, noise):
  """Linear + bias -> unimix logits -> Gumbel-argmax one-hot per group, fused; sampled indices must
  agree with the reference except where logit + noise ties to rounding."""
  def case(be, T):
    N = S * C
    X, Xlo = act_pair(T, rnd(M, K, seed=1))
    Wt, Wtlo = wt_pair(T, rnd(K, N, seed=2, scale=K ** -0.5))
    bias = T(0.1 * rnd(N, seed=3))
    g = torch.Generator().manual_seed(4)
    G = T(-torch.log(-torch.log(torch.rand(M, N, generator=g).clamp(1e-6, 1 - 1e-6))))
    logit, oh = T(torch.zeros(M, N)), T(torch.zeros(M, N))
    L = P.Layer([P.Seg(X, Xlo)], Wt, Wtlo, N, P.EPI_STATS, bias=bias, out=logit, x0=G if noise else None,
                x1=oh, f0=0.01, ipar=C)
    be.program_run(be.program_build([L], M))
    return {'logit': logit, 'oh': oh}
  hs, cu = backends
  ref = case(hs, lambda x: x.clone())
  out = case(cu, lambda x: x.clone().to(cu.device))
  cu.sync()
  np.testing.assert_allclose(out['logit'].cpu().numpy(), ref['logit'].numpy(), rtol=2e-5, atol=2e-5)
  a, b = out['oh'].cpu().numpy(), ref['oh'].numpy()
  assert np.all(a.reshape(M, S, C).sum(-1) == 1.0)
  bad = int((a != b).any(-1).sum())
  assert bad <= 2, bad


def test_actor_layer(backends):
  def case(be, T):
    M, K, A = 1024, 512, 12
    X, Xlo = act_pair(T, rnd(M, K, seed=1))
    Wt, Wtlo = wt_pair(T, rnd(K, 2 * A, seed=2, scale=K ** -0.5))
    bias = T(0.1 * rnd(2 * A, seed=3))
    eps = T(rnd(M, A, seed=4))
    act, act_lo, mean, std = (T(torch.zeros(M, A)) for _ in range(4))
    L = P.Layer([P.Seg(X, Xlo)], Wt, Wtlo, 2 * A, P.EPI_ACTOR, bias=bias, out=act, out_lo=act_lo, x0=eps,
                x1=mean, x2=std, f0=0.1, f1=1.0, ipar=A)
    be.program_run(be.program_build([L], M))
    return {'out': act, 'lo': act_lo, 'mean': mean, 'std': std}
  run_both(backends, case, rtol=2e-5, atol=2e-5)


def test_stats_bwd_layer(backends):
  """Epilogue-only layer: straight-through categorical gradient back to the raw logits."""
  def case(be, T):
    M, S, C = 1024, 32, 32
    N = S * C
    g = torch.Generator().manual_seed(3)
    z = torch.randn(M, S, C, generator=g)
    lg = torch.log(0.99 * torch.softmax(z, -1) + 0.01 / C).reshape(M, N)
    feat = T(rnd(M, 512 + N, seed=4))            # d(stoch) is a column slice of a wider buffer
    dzs, dzs_lo = T(torch.zeros(M, N)), T(torch.zeros(M, N))
    L = P.Layer([], None, None, N, P.EPI_STATSBWD, add=feat[:, 512:], out=dzs, out_lo=dzs_lo, x0=T(lg),
                f0=0.01, ipar=C)
    be.program_run(be.program_build([L], M))
    return {'out': dzs, 'lo': dzs_lo}
  run_both(backends, case, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize('M,D', [(1024, 512), (300, 256)])
def test_gru_bwd_layer(backends, M, D):
  """dZ0 @ W0^T + head gradient -> GRU gates backward -> joint LayerNorm backward over 3 D, and the
  gradient wrt the previous deter accumulated into a strided slice."""
  def case(be, T):
    U = D
    dZ0, dZ0lo = act_pair(T, rnd(M, U, seed=1, scale=0.1))
    W0 = rnd(D, U, seed=2, scale=U ** -0.5)                 # kernel [D, U]: d deter = dZ0 @ W0^T
    pw = T(torch.stack([W0, P.lo_of(W0)]))
    z = rnd(M, 3 * D, seed=3)
    gamma, beta = 1.0 + 0.1 * rnd(3 * D, seed=4), 0.1 * rnd(3 * D, seed=5)
    y, st = torch.zeros(M, 3 * D), torch.zeros(M, 2)
    HostSim().ln_fwd(z, gamma, beta, y, st, 0)
    prev = rnd(M, D, seed=6)
    dtf_t = T(rnd(M, D + 64, seed=7, scale=0.1))            # head gradients at step t (strided view)
    dtf_k = T(rnd(M, D + 64, seed=8, scale=0.1))            # accumulates d prev
    dzg, dzg_lo = T(torch.zeros(M, 3 * D)), T(torch.zeros(M, 3 * D))
    L = P.Layer([P.Seg(dZ0, dZ0lo)], pw[0], pw[1], D, P.EPI_GRUBWD, add=dtf_t[:, :D], gamma=T(gamma),
                beta=T(beta), z=T(z), st=T(st), out=dzg, out_lo=dzg_lo, x0=T(prev), x1=T(y),
                x2=dtf_k[:, :D])
    be.program_run(be.program_build([L], M))
    return {'out': dzg, 'lo': dzg_lo, 'dprev': dtf_k}
  run_both(backends, case, rtol=1e-4, atol=2e-5)
