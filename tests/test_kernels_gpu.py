"""-m gpu: every CUDA primitive, called through the C-ABI (ctypes), against the
torch-CPU statement of the same primitive (tests/hostsim.py) on seeded random
inputs, including row-strided views, ragged sizes and both class counts."""
import numpy as np
import pytest
import torch

from tests.hostsim import HostSim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def backends():
  from daydreamer_b200.backend import CudaBackend
  return HostSim(), CudaBackend()


def both(backends, case, rtol=2e-5, atol=2e-6, exact=()):
  """case(be, T) -> dict of tensors; T moves a CPU tensor to be's device."""
  hs, cu = backends
  ref = case(hs, lambda x: x.clone())
  out = case(cu, lambda x: x.clone().to(cu.device))
  cu.sync()
  for k in ref:
    a, b = ref[k].double().cpu().numpy(), out[k].double().cpu().numpy()
    if k in exact:
      assert np.array_equal(a, b), k
    else:
      err = np.abs(a - b)
      tol = atol + rtol * np.abs(a)
      assert (err <= tol).all(), (k, err.max(), np.abs(a).max())


def rnd(*shape, seed=0, scale=1.0):
  g = torch.Generator().manual_seed(seed)
  return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize('M,N,K', [(16, 64, 40), (16, 1536, 1024), (5, 7, 3), (130, 70, 33),
                                   (1024, 512, 512), (300, 1, 512), (2048, 1536, 100)])
@pytest.mark.parametrize('ta,tb', [(False, False), (False, True), (True, False), (True, True)])
def test_gemm(backends, M, N, K, ta, tb):
  def case(be, T):
    A = T(rnd(*((K, M + 3) if ta else (M, K + 3)), seed=1))
    B = T(rnd(*((N, K + 5) if tb else (K, N + 5)), seed=2))
    C = T(rnd(M, N + 2, seed=3))
    bias = T(rnd(N, seed=4))
    a = A[:, :M] if ta else A[:, :K]
    b = B[:, :K] if tb else B[:, :N]
    c = C[:, :N]
    be.gemm(a, b, c, ta=ta, tb=tb, beta=0.0, bias=bias)
    c1 = c.clone()
    be.gemm(a, b, c, ta=ta, tb=tb, beta=1.0)
    return {'c0': c1, 'c1': c.clone(), 'pad': C[:, N:].clone()}
  both(backends, case, rtol=1e-4, atol=1e-4 * np.sqrt(K))


@pytest.mark.parametrize('M,N,K', [(1024, 12, 512), (1024, 24, 512), (1024, 512, 12), (15360, 1, 512),
                                   (15360, 512, 1), (1000, 3, 300), (70, 32, 1536), (333, 5, 20),
                                   (1024, 1536, 12), (900, 2, 16), (257, 31, 33)])
@pytest.mark.parametrize('tb', [False, True])
@pytest.mark.parametrize('level', [1, 2])
def test_gemm_thin(backends, M, N, K, tb, level):
  """Thin shapes (N <= 32 or K <= 32, many rows) on the streaming kernels of dd_gemm_thin.cu,
  contiguous (vector path) and row-strided operands, bias / beta combinations."""
  def case(be, T):
    out = {}
    prev = be.lib.dd_set_gemm_thin(level) if hasattr(be, 'lib') else None     # off by default
    for pad in (0, 3):
      A = T(rnd(M, K + pad, seed=1))[:, :K]
      B = T(rnd(*((N, K + pad) if tb else (K, N + pad)), seed=2))
      B = B[:, :K] if tb else B[:, :N]
      C = T(rnd(M, N + pad, seed=3))[:, :N]
      bias = T(rnd(N + 4, seed=4))[:N] if pad == 0 else None
      be.gemm(A, B, C, tb=tb, beta=0.0, bias=bias)
      out[f'c0_{pad}'] = C.clone()
      be.gemm(A, B, C, tb=tb, beta=1.0)
      out[f'c1_{pad}'] = C.clone()
    if prev is not None:
      be.lib.dd_set_gemm_thin(prev)
    return out
  both(backends, case, rtol=1e-4, atol=1e-4 * np.sqrt(K))


def test_gemm_splitk_wgrad(backends):
  def case(be, T):
    X = T(rnd(16384, 200, seed=1))
    dY = T(rnd(16384, 96, seed=2, scale=0.1))
    G = T(torch.zeros(200, 96))
    be.gemm(X, dY, G, ta=True, beta=1.0)
    return {'g': G}
  both(backends, case, rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize('M,N,act', [(16, 512, 1), (1000, 1536, 0), (3, 20, 1), (70, 6144, 0),
                                     (257, 72, 1), (1024, 512, 1), (777, 256, 0), (16384, 512, 1),
                                     (300, 200, 1), (50, 640, 1)])
@pytest.mark.parametrize('warp_max', [128, 512])
def test_ln(backends, M, N, act, warp_max):
  def case(be, T):
    prev = be.lib.dd_set_ln_warp_max(warp_max) if hasattr(be, 'lib') else None
    Z = T(rnd(M, N + 4, seed=1, scale=2.0))[:, 2:2 + N]
    sc, bi = T(1 + 0.2 * rnd(N, seed=2)), T(0.3 * rnd(N, seed=3))
    H, st = T(torch.zeros(M, N)), T(torch.zeros(M, 2))
    be.ln_fwd(Z, sc, bi, H, st, act)
    dH = T(rnd(M, N, seed=4))
    dZ, ds, db = T(torch.zeros(M, N)), T(0.1 * rnd(N, seed=5)), T(0.1 * rnd(N, seed=6))
    be.ln_bwd(dH, Z, st, sc, bi, dZ, ds, db, act)
    dZ2 = T(torch.zeros(M, N))
    be.ln_bwd(dH, Z, st, sc, bi, dZ2, None, None, act)
    if prev is not None:
      be.lib.dd_set_ln_warp_max(prev)
    return {'H': H, 'st': st, 'dZ': dZ, 'ds': ds, 'db': db, 'dZ2': dZ2}
  both(backends, case, rtol=1e-4, atol=1e-4 * np.sqrt(M))


@pytest.mark.parametrize('M,N,act', [(1024, 512, 1), (15360, 512, 1), (777, 256, 0), (50, 128, 1),
                                     (333, 384, 0), (16, 512, 1)])
def test_ln_vec_experimental(backends, M, N, act):
  """Round-2 work in progress: the float4 warp-per-row LayerNorm kernels (dd_set_ln_vec) on
  contiguous, 16-byte aligned rows (the only layout they accept)."""
  def case(be, T):
    if hasattr(be, 'lib'):
      be.lib.dd_set_ln_vec(1)
    Z = T(rnd(M, N, seed=1, scale=2.0))
    sc, bi = T(1 + 0.2 * rnd(N, seed=2)), T(0.3 * rnd(N, seed=3))
    H, st = T(torch.zeros(M, N)), T(torch.zeros(M, 2))
    be.ln_fwd(Z, sc, bi, H, st, act)
    dH = T(rnd(M, N, seed=4))
    dZ, ds, db = T(torch.zeros(M, N)), T(0.1 * rnd(N, seed=5)), T(0.1 * rnd(N, seed=6))
    be.ln_bwd(dH, Z, st, sc, bi, dZ, ds, db, act)
    dZ2 = T(torch.zeros(M, N))
    be.ln_bwd(dH, Z, st, sc, bi, dZ2, None, None, act)
    if hasattr(be, 'lib'):
      be.lib.dd_set_ln_vec(0)
    return {'H': H, 'st': st, 'dZ': dZ, 'ds': ds, 'db': db, 'dZ2': dZ2}
  both(backends, case, rtol=1e-4, atol=1e-4 * np.sqrt(M))


@pytest.mark.parametrize('M,D', [(16, 512), (1024, 24), (7, 5)])
def test_gru(backends, M, D):
  def case(be, T):
    Y = T(rnd(M, 3 * D, seed=1))
    feat = T(rnd(M, D + 9, seed=2))
    prev = feat[:, :D]
    out = T(torch.zeros(M, D + 3))[:, 1:1 + D]
    be.gru_fwd(Y, prev, out)
    dO = T(rnd(M, D, seed=3))
    dY, dP = T(torch.zeros(M, 3 * D)), T(rnd(M, D, seed=4))
    dP2 = dP.clone()
    be.gru_bwd(dO, Y, prev, dY, dP, True)
    be.gru_bwd(dO, Y, prev, dY, dP2, False)
    return {'out': out.clone(), 'dY': dY, 'dP': dP, 'dP2': dP2}
  both(backends, case)


@pytest.mark.parametrize('M,S,C', [(16, 32, 32), (100, 48, 48), (5, 4, 6), (1024, 32, 32)])
def test_stats(backends, M, S, C):
  def case(be, T):
    SC = S * C
    Z = T(rnd(M, SC, seed=1, scale=1.5))
    G = T(-torch.log(-torch.log(torch.rand(M, SC, generator=torch.Generator().manual_seed(2)).clamp(1e-6, 1 - 1e-6))))
    feat = T(torch.zeros(M, 7 + SC))
    logit, idx = T(torch.zeros(M, SC)), T(torch.zeros(M, S, dtype=torch.int32))
    be.stats_fwd(Z, G, logit, feat[:, 7:], idx, S, C, 0.01, 0)
    st1, idx1 = T(torch.zeros(M, SC)), T(torch.zeros(M, S, dtype=torch.int32))
    lg1 = T(torch.zeros(M, SC))
    be.stats_fwd(Z, None, lg1, st1, idx1, S, C, 0.01, 1)
    lg2 = T(torch.zeros(M, SC))
    be.stats_fwd(Z, None, lg2, None, None, S, C, 0.01, 2)
    dS, dL = T(rnd(M, SC, seed=3)), T(rnd(M, SC, seed=4))
    dZa, dZb, dZc = (T(torch.zeros(M, SC)) for _ in range(3))
    be.stats_bwd(dS, dL, logit, dZa, S, C, 0.01)
    be.stats_bwd(dS, None, logit, dZb, S, C, 0.01)
    be.stats_bwd(None, dL, logit, dZc, S, C, 0.01)
    return {'logit': logit, 'stoch': feat[:, 7:].clone(), 'idx': idx, 'st1': st1, 'idx1': idx1,
            'lg2': lg2, 'dZa': dZa, 'dZb': dZb, 'dZc': dZc}
  both(backends, case, rtol=1e-4, atol=1e-5, exact=('stoch', 'idx', 'st1', 'idx1'))


@pytest.mark.parametrize('M,S,C', [(64, 32, 32), (9, 48, 48), (1024, 4, 6)])
def test_kl(backends, M, S, C):
  def case(be, T):
    SC = S * C
    lsm = lambda x: torch.log_softmax(x.reshape(M, S, C), -1).reshape(M, SC)
    post, prior = T(lsm(rnd(M, SC, seed=1))), T(lsm(rnd(M, SC, seed=2)))
    kl, ea, eb = T(torch.zeros(M)), T(torch.zeros(M)), T(torch.zeros(M))
    be.kl_fwd(post, prior, S, C, 0.8, kl, ea, eb)
    scale = T(torch.tensor([0.37]))
    dp, dq = T(torch.zeros(M, SC)), T(torch.zeros(M, SC))
    be.kl_bwd(post, prior, S, C, 0.8, 1.0 / M, scale, dp, dq)
    return {'kl': kl, 'ea': ea, 'eb': eb, 'dp': dp, 'dq': dq}
  both(backends, case, rtol=1e-4, atol=1e-5)


def test_mask(backends):
  M, D, SC, A = 16, 40, 24, 5
  def case(be, T):
    feat = T(rnd(M, D + SC, seed=1))
    act = T(rnd(M, A, seed=2))
    first = T((torch.arange(M) % 3 == 0).float())
    idd, ids = T(rnd(D, seed=3)), T(rnd(SC, seed=4))
    od, os_, oa = T(torch.zeros(M, D)), T(torch.zeros(M, SC)), T(torch.zeros(M, A))
    be.mask_fwd(feat[:, :D], feat[:, D:], act, first, idd, ids, od, os_, oa)
    dd, ds = T(rnd(M, D, seed=5)), T(rnd(M, SC, seed=6))
    dfeat = T(rnd(M, D + SC, seed=7))
    dinit = T(rnd(D, seed=8))
    be.mask_bwd(dd, ds, first, dfeat[:, :D], dfeat[:, D:], dinit)
    dinit2 = T(torch.zeros(D))
    be.mask_bwd(dd, ds, first, None, None, dinit2)
    return {'od': od, 'os': os_, 'oa': oa, 'dfeat': dfeat, 'dinit': dinit, 'dinit2': dinit2}
  both(backends, case)


def test_elementwise(backends):
  def case(be, T):
    a, b = T(rnd(33, 50, seed=1)), T(rnd(33, 50, seed=2))
    be.add_(a[:, 3:20], b[:, 5:22])
    c = T(torch.zeros(33, 17))
    be.copy_(c, a[:, 10:27])
    f = T(rnd(20, 30, seed=3))
    be.fill_(f[:, 4:9], 2.5)
    s = T(torch.tensor([0.3]))
    be.axpy_(b[:, :7], a[:, :7], s)
    d = T(torch.zeros(33, 50))
    be.sub(d, a, b)
    e = T(torch.zeros(33, 50))
    be.symexp(a, e)
    src = T(rnd(6, 9, 11, seed=4))
    dst = T(torch.zeros(9, 6, 11))
    be.transpose01(src, dst)
    cs, cs2 = T(torch.zeros(17)), T(rnd(17, seed=5))
    X = T(rnd(5000, 30, seed=6))
    be.colsum(X[:, 2:19], cs, acc=False)
    be.colsum(X[:, 2:19], cs2, acc=True)
    th = T(torch.zeros(33, 50))
    be.tanh_fwd(a, th)
    dth = T(rnd(33, 50, seed=7))
    be.tanh_bwd(b, th, dth, True)
    f64 = T(torch.zeros(3, 3, dtype=torch.float64))
    be.fill_(f64, 1.25)
    return {'a': a, 'c': c, 'f': f, 'b': b, 'd': d, 'e': e, 'dst': dst, 'cs': cs, 'cs2': cs2,
            'th': th, 'dth': dth, 'f64': f64}
  both(backends, case, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize('kind', [0, 1, 2])
def test_head_loss_balance(backends, kind):
  M, dim = 1000, (7 if kind == 0 else 1)
  def case(be, T):
    pred = T(rnd(M, dim + 2, seed=1))[:, :dim]
    tgt = T(rnd(M, dim, seed=2)) if kind != 2 else T((rnd(M, 1, seed=2) > 0).float())
    w = T(torch.rand(M, generator=torch.Generator().manual_seed(3)))
    lr, dp = T(torch.zeros(M)), T(torch.zeros(M, dim))
    be.head_loss(kind, pred, tgt, w, 0.01, lr, dp)
    lr2, dp2 = T(torch.zeros(M)), T(torch.zeros(M, dim))
    be.head_loss(kind, pred, tgt, None, 0.01, lr2, dp2)
    out = {'lr': lr, 'dp': dp, 'lr2': lr2, 'dp2': dp2}
    if kind:
      bal = T(torch.zeros(8, dtype=torch.float64))
      be.balance(kind, pred.contiguous(), lr2, tgt, 0.1 if kind == 1 else 0.5, bal)
      out['bal'] = bal
    return out
  both(backends, case, rtol=1e-4, atol=1e-5)


def test_actor_and_lambda(backends):
  H, N, A = 5, 300, 4
  T1 = H + 1
  def case(be, T):
    mraw, sraw, eps = T(rnd(T1 * N, A, seed=1)), T(rnd(T1 * N, A, seed=2)), T(rnd(T1 * N, A, seed=3))
    act, mean, std = (T(torch.zeros(T1 * N, A)) for _ in range(3))
    be.actor_sample(mraw, sraw, eps, 0.1, 1.0, act, mean, std)
    M = H * N
    ent = T(torch.zeros(M, A))
    be.actor_ent(std[:M], 0.1, 1.0, ent)
    rho, vraw, kraw = T(rnd(T1, N, seed=4, scale=0.5)), T(rnd(T1, N, seed=5, scale=0.5)), T(rnd(T1, N, seed=6) + 2)
    cont0 = T((rnd(N, seed=7) > -1).float())
    reward, value, cont, weight, ret = T(torch.zeros(H, N)), T(torch.zeros(T1, N)), T(torch.zeros(T1, N)), T(torch.zeros(T1, N)), T(torch.zeros(H, N))
    be.lambda_fwd(rho, vraw, kraw, cont0, 0.995, 0.95, reward, value, cont, weight, ret)
    adv = T(torch.tensor([0.05, 1.7]))
    scale = T(torch.rand(A, generator=torch.Generator().manual_seed(8)))
    al = T(torch.zeros(M))
    be.actor_loss(ret, value, adv, ent, scale, weight, al)
    drho, dv, dk = (T(torch.zeros(T1, N)) for _ in range(3))
    be.lambda_bwd(rho, vraw, value, cont, weight, ret, adv, 1.0 / M, 0.995, 0.95, drho, dv, dk)
    dA = T(rnd(M, A, seed=9))
    dm, dsr = T(torch.zeros(M, A)), T(torch.zeros(M, A))
    be.actor_bwd(dA, eps[:M], mean[:M], std[:M], weight.view(-1)[:M], scale, 1.0 / M, 0.1, 1.0, dm, dsr)
    return {'act': act, 'mean': mean, 'std': std, 'ent': ent, 'reward': reward, 'value': value,
            'cont': cont, 'weight': weight, 'ret': ret, 'al': al, 'drho': drho, 'dv': dv, 'dk': dk,
            'dm': dm, 'dsr': dsr}
  both(backends, case, rtol=2e-4, atol=1e-6)


@pytest.mark.parametrize('H,N,A', [(3, 50, 6), (15, 1024, 6), (2, 9, 40)])
def test_actor_discrete(backends, H, N, A):
  def case(be, T):
    M = H * N
    raw = T(rnd(M, A, seed=1))
    G = T(rnd(M, A, seed=2))
    logit, action = T(torch.zeros(M, A)), T(torch.zeros(M, A))
    be.stats_fwd(raw, G, logit, action, None, 1, A, 0.1, 0)
    ent = T(torch.zeros(M, 1))
    be.actor_disc_ent(logit, ent)
    ret, value = T(rnd(H, N, seed=3)), T(rnd(H + 1, N, seed=4))
    weight = T(torch.rand(H + 1, N, generator=torch.Generator().manual_seed(5)))
    adv, scale = T(torch.tensor([0.1, 1.7])), T(torch.tensor([0.3]))
    loss, dlogit = T(torch.zeros(M)), T(torch.zeros(M, A))
    be.actor_disc_loss(logit, action, ret, value, adv, ent, scale, weight, 1.0 / M, loss, dlogit)
    draw = T(torch.zeros(M, A))
    be.stats_bwd(None, dlogit, logit, draw, 1, A, 0.1)
    return {'logit': logit, 'action': action, 'ent': ent, 'loss': loss, 'dlogit': dlogit,
            'draw': draw}
  both(backends, case, rtol=1e-4, atol=1e-6, exact=('action',))


@pytest.mark.parametrize('N,H,W,C,k', [(3, 18, 18, 4, 4), (2, 31, 31, 64, 4), (5, 10, 10, 3, 4),
                                       (2, 13, 13, 8, 5), (3, 30, 30, 8, 6), (2, 64, 64, 3, 6),
                                       (4, 5, 5, 16, 5), (7, 3, 3, 8, 3)])
def test_im2col_col2im(backends, N, H, W, C, k):
  def case(be, T):
    Ho = (H - k) // 2 + 1
    X = T(rnd(N, H, W, C, seed=1))
    col = T(torch.zeros(N * Ho * Ho, k * k * C))
    be.im2col(X, k, 2, col)
    col2 = T(rnd(N * Ho * Ho, k * k * C, seed=2))
    bias = T(rnd(C, seed=3))
    Y, Y0 = T(rnd(N, H, W, C, seed=4)), T(rnd(N, H, W, C, seed=5))
    be.col2im(col2, k, 2, bias, Y)
    be.col2im(col2, k, 2, None, Y0)
    return {'col': col, 'Y': Y, 'Y0': Y0}
  both(backends, case, rtol=1e-5, atol=1e-5, exact=('col',))


def test_image_stage_and_loss(backends):
  def case(be, T):
    B, L, H, W = 3, 4, 10, 10
    g = torch.Generator().manual_seed(0)
    rgb = torch.randint(0, 256, (B, L, H, W, 3), generator=g, dtype=torch.uint8)
    dep = torch.randint(0, 256, (B, L, H, W, 1), generator=g, dtype=torch.uint8)
    img = T(torch.zeros(L * B, H, W, 4))
    be.image_stage(T(rgb), img, 0, 1.0 / 255.0)
    be.image_stage(T(dep), img, 3, 1.0 / 255.0)
    Z = T(rnd(L * B, H, W, 4, seed=1))
    l0, l1 = T(torch.zeros(L * B)), T(torch.zeros(L * B))
    dZ = T(torch.zeros(L * B, H, W, 4))
    be.image_loss(Z, img, 0, 3, 0.01, l0, dZ)
    be.image_loss(Z, img, 3, 4, 0.01, l1, dZ)
    sg = T(torch.zeros(L * B, H, W, 4))
    be.sigmoid(Z, sg)
    return {'img': img, 'l0': l0, 'l1': l1, 'dZ': dZ, 'sg': sg}
  both(backends, case, rtol=2e-5, atol=1e-6)


def test_controllers_and_adam(backends):
  def case(be, T):
    rows = T(torch.zeros(3, 8, dtype=torch.float64))
    be.stats_init(rows)
    X = T(rnd(700, 90, seed=1))
    be.stats(X[:, 5:77], rows[0])
    be.stats(X[:3, :1], rows[1])
    scale, sums = T(torch.tensor([1.0, 0.5, 1e-5, 1.0])), T(torch.tensor([10.0, 50.0, 90.0, 52.0]))
    be.autoadapt(scale, sums, 100.0, 0.5, 1e-5, 1.0, 0.1, 0.1, True)
    scale2 = T(torch.tensor([1.0]))
    be.autoadapt(scale2, T(torch.tensor([800.0])), 100.0, 3.5, 1e-3, 1.0, 0.1, 0.1, False)
    state, adv = T(torch.zeros(3, dtype=torch.float64)), T(torch.zeros(2))
    be.normalize_update(state, rows[0], 0.99, 1e8, adv)
    be.normalize_update(state, rows[0], 0.99, 1e8, adv)
    n = 100003
    p, g = T(rnd(n, seed=2)), T(rnd(n, seed=3, scale=30.0))
    m, v = T(torch.zeros(n)), T(torch.zeros(n))
    ss = T(torch.zeros(1, dtype=torch.float64))
    cnt = T(torch.tensor([0, 0, 0, -1, 0, 0, 0, 0], dtype=torch.int64))
    for _ in range(3):
      be.sumsq(g, ss)
      be.adam(p, g, m, v, 60000, 1e-4, 0.9, 0.999, 1e-6, 1e-2, 100.0, 0.5, ss, cnt, 2)
    src, dst = T(rnd(1000, seed=4)), T(torch.zeros(1000))
    be.slow_sync(src, dst, cnt, 3, 2)
    d1 = dst.clone()
    src += 1.0
    be.slow_sync(src, dst, cnt, 3, 2)
    d2 = dst.clone()
    be.slow_sync(src, dst, cnt, 3, 2)
    be.counter_inc(cnt, 4)
    return {'rows': rows[:2], 'scale': scale, 'scale2': scale2, 'state': state, 'adv': adv,
            'p': p, 'm': m, 'v': v, 'ss': ss, 'cnt': cnt, 'd1': d1, 'd2': d2, 'd3': dst}
  both(backends, case, rtol=1e-5, atol=1e-6)


def test_noise_matches_numpy_philox(backends):
  from oracle import philox
  hs, cu = backends
  cnt = cu.upload(np.array([0, 0, 0, 0, 7, 0, 0, 0]), dtype=torch.int64)
  for n in (10, 4096 + 3):
    g, z = cu.zeros(n), cu.zeros(n)
    cu.fill_gumbel(g, 1234567890123, 2, cnt, 4)
    cu.fill_normal(z, 1234567890123, 3, cnt, 4)
    cu.sync()
    np.testing.assert_allclose(g.cpu().numpy(), philox.gumbel(n, 1234567890123, 2, 7), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(z.cpu().numpy(), philox.normal(n, 1234567890123, 3, 7), rtol=2e-4, atol=2e-5)


TC_SHAPES = [(1024, 512, 512), (1024, 1536, 1024), (16384, 512, 1536), (200, 96, 64),
             (129, 36, 40), (1024, 512, 1036), (4096, 256, 12 * 4)]


@pytest.mark.parametrize('M,N,K', TC_SHAPES)
@pytest.mark.parametrize('ta,tb', [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_tcgen05_3xtf32(backends, M, N, K, ta, tb):
  """The tensor-core kernel must stay at fp32-class accuracy (3xTF32 split)."""
  hs, cu = backends
  A = rnd(*((K, M) if ta else (M, K)), seed=1)
  B = rnd(*((N, K) if tb else (K, N)), seed=2)
  C0 = rnd(M, N, seed=3)
  bias = rnd(N, seed=4)
  ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double()
  scale = float(ref.abs().max())
  dA, dB, dbias = A.cuda(), B.cuda(), bias.cuda()
  outs = {}
  for on in (1, 0):
    cu.set_gemm_tc(on)
    before = cu.launch_count()
    c = C0.cuda()
    cu.gemm(dA, dB, c, ta=ta, tb=tb, beta=0.0, bias=dbias)
    c0 = c.clone()
    cu.gemm(dA, dB, c, ta=ta, tb=tb, beta=1.0)
    cu.sync()
    outs[on] = (c0.cpu().double(), c.cpu().double())
  cu.set_gemm_tc(1)
  # fp32 SIMT rounds per FMA; the tensor core truncates once per K=8 accumulate,
  # so its error grows ~linearly in K/8 ulps (measured 4e-6 at K=512)
  tol = {0: 4e-6 * max(1.0, K / 4096), 1: 1.2e-5 * max(1.0, K / 1024)}
  for on, (c0, c1) in outs.items():
    e0 = float((c0 - (ref + bias.double())).abs().max()) / scale
    e1 = float((c1 - (2 * ref + bias.double())).abs().max()) / scale
    assert e0 < tol[on] and e1 < 2 * tol[on], (on, e0, e1)
  # tensor-core result vs SIMT result, elementwise
  assert float((outs[1][0] - outs[0][0]).abs().max()) / scale < tol[1]


def test_gemm_tcgen05_splitk_wgrad(backends):
  hs, cu = backends
  X, dY = rnd(16384, 512, seed=1), rnd(16384, 512, seed=2, scale=0.1)
  ref = X.double().t() @ dY.double()
  G = torch.zeros(512, 512).cuda()
  cu.set_gemm_tc(1)
  cu.gemm(X.cuda(), dY.cuda(), G, ta=True, beta=1.0)
  cu.sync()
  assert float((G.cpu().double() - ref).abs().max()) / float(ref.abs().max()) < 5e-5


@pytest.mark.parametrize('M,N,K', [(16, 512, 1036), (16, 1536, 512), (16, 1024, 512), (32, 512, 1536),
                                   (9, 100, 300), (16, 40, 64), (24, 2304, 2048)])
@pytest.mark.parametrize('tb', [False, True])
def test_gemm_skinny(backends, M, N, K, tb):
  """The observe-scan (M = sequences per GPU) GEMM path, strided views included."""
  def case(be, T):
    A = T(rnd(M, K + 8, seed=1))[:, 4:4 + K]
    B = T(rnd(*((N, K) if tb else (K, N)), seed=2))
    C = T(rnd(M, N + 4, seed=3))
    bias = T(rnd(N, seed=4))
    c = C[:, :N]
    be.gemm(A, B, c, tb=tb, beta=0.0, bias=bias)
    c1 = c.clone()
    be.gemm(A, B, c, tb=tb, beta=1.0)
    return {'c0': c1, 'c1': c.clone(), 'pad': C[:, N:].clone()}
  both(backends, case, rtol=1e-4, atol=2e-5 * np.sqrt(K))


@pytest.mark.parametrize('M,N,K', [(1024, 512, 512), (1024, 512, 1536), (1024, 1024, 512), (130, 72, 100),
                                   (1000, 64, 36), (512, 512, 1024)])
@pytest.mark.parametrize('ta,tb', [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_tc_staged_loader(backends, M, N, K, ta, tb):
  """Round-2 work in progress: the cp.async-staged operand loader of the 64-wide tcgen05 tile
  (dd_set_gemm_tc_staged) against fp64, same bound as the register-staged loader."""
  hs, cu = backends
  A = rnd(*((K, M) if ta else (M, K)), seed=1).to(cu.device)
  B = rnd(*((N, K) if tb else (K, N)), seed=2).to(cu.device)
  C = torch.zeros(M, N, device=cu.device)
  prev = cu.lib.dd_set_gemm_tc_staged(1)
  try:
    cu.gemm(A, B, C, ta=ta, tb=tb)
    cu.sync()
  finally:
    cu.lib.dd_set_gemm_tc_staged(prev)
  ref = (A.t() if ta else A).double() @ (B.t() if tb else B).double()
  err = float((C.double() - ref).abs().max() / ref.abs().max())
  assert err < 2e-5 * max(1.0, K / 512), err
