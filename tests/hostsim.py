"""Torch-CPU simulation of the kernel primitives  --  TEST INFRASTRUCTURE ONLY.

Every method here states, in plain torch ops on CPU tensors, the exact
semantics of one CUDA primitive exported by the C-ABI library
(include/dd_kernels.h).  It exists for two reasons:
  1. `-m "not gpu"` tests run the product's orchestration (daydreamer_b200/
     engine.py: hand-derived forward + backward of the whole train step) on
     this simulator and compare it with the oracle's autograd, so the host
     logic is validated where there is no GPU;
  2. `-m gpu` tests compare each CUDA primitive with the method of the same
     name on random inputs.
The product never imports this file: daydreamer_b200.backend.CudaBackend is the
only backend the Agent constructs, and it raises if the .so or the GPU is
missing.
"""
import contextlib
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import philox

LN_EPS = 1e-3


class HostSim:
  name = 'hostsim'
  device = torch.device('cpu')

  # -- allocation -----------------------------------------------------------
  def zeros(self, *shape, dtype=torch.float32):
    return torch.zeros(*shape, dtype=dtype)

  def upload(self, array, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(array)).to(dtype).clone()

  @contextlib.contextmanager
  def branch(self):
    """CudaBackend.branch: a second stream.  Here: the same launch order, issued serially."""
    yield

  def join(self):
    pass

  def sync(self):
    pass

  # -- dense ----------------------------------------------------------------
  def gemm(self, A, B, C, ta=False, tb=False, beta=0.0, bias=None):
    a = A.t() if ta else A
    b = B.t() if tb else B
    r = a @ b
    if bias is not None:
      r = r + bias
    C.copy_(r + beta * C if beta else r)

  def ln_fwd(self, Z, scale, bias, H, stats, act):
    mean = Z.mean(-1, keepdim=True)
    var = ((Z - mean) ** 2).mean(-1, keepdim=True)
    rstd = torch.rsqrt(var + LN_EPS)
    y = (Z - mean) * rstd * scale + bias
    H.copy_(F.elu(y) if act else y)
    stats[:, 0] = mean[:, 0]
    stats[:, 1] = rstd[:, 0]

  def ln_bwd(self, dH, Z, stats, scale, bias, dZ, dscale, dbias, act):
    mean, rstd = stats[:, 0:1], stats[:, 1:2]
    xhat = (Z - mean) * rstd
    y = xhat * scale + bias
    dy = dH * torch.where(y > 0, torch.ones_like(y), torch.exp(y)) if act else dH
    if dscale is not None:
      dscale += (dy * xhat).sum(0)
      dbias += dy.sum(0)
    dx = dy * scale
    dZ.copy_(rstd * (dx - dx.mean(-1, keepdim=True)
                     - xhat * (dx * xhat).mean(-1, keepdim=True)))

  def gru_fwd(self, Y, prev, out):
    D = prev.shape[1]
    r = torch.sigmoid(Y[:, :D])
    c = torch.tanh(r * Y[:, D:2 * D])
    u = torch.sigmoid(Y[:, 2 * D:] - 1.0)
    out.copy_(u * c + (1.0 - u) * prev)

  def gru_bwd(self, dOut, Y, prev, dY, dPrev, acc):
    D = prev.shape[1]
    yr, yc, yu = Y[:, :D], Y[:, D:2 * D], Y[:, 2 * D:]
    r = torch.sigmoid(yr)
    c = torch.tanh(r * yc)
    u = torch.sigmoid(yu - 1.0)
    du = dOut * (c - prev)
    dc = dOut * u
    dpre = dc * (1.0 - c * c)
    dY[:, :D] = dpre * yc * r * (1.0 - r)
    dY[:, D:2 * D] = dpre * r
    dY[:, 2 * D:] = du * u * (1.0 - u)
    dp = dOut * (1.0 - u)
    dPrev.copy_(dPrev + dp if acc else dp)

  # -- layer machine (csrc/dd_machine.cu) --------------------------------------
  def program_build(self, layers, rows):
    return (list(layers), int(rows))

  def program_run(self, h):
    """Each layer: acc = sum_seg A_seg @ B[wn0:wn0+N, wk0:wk0+k].T with every operand
    reconstructed from its (x, x_lo) pair exactly as the tensor core sees it, so a stale or
    missing x_lo shows up as a numerical error in the CPU tests."""
    from daydreamer_b200 import program as P
    layers, rows = h
    for L in layers:
      acc = torch.zeros(rows, L.N, dtype=torch.float64)
      for sg in L.segs:
        k = sg.a.shape[1]
        a = P.trunc19(sg.a).double() + (0.0 if sg.a_lo is None else sg.a_lo.double())
        if sg.a_lo is None:
          assert torch.equal(P.trunc19(sg.a), sg.a.contiguous()), 'operand declared exact is not'
        wsl = (slice(L.wn0, L.wn0 + L.N), slice(sg.wk0, sg.wk0 + k))
        b = P.trunc19(L.w[wsl]).double() + L.w_lo[wsl].double()
        acc += a @ b.t()
      acc = acc.float()
      if L.add is not None:
        acc = acc + L.add
      if L.bias is not None:
        acc = acc + L.bias
      if L.epi == P.EPI_PLAIN:
        out = acc
      elif L.epi == P.EPI_LN:
        if L.z is not None:
          L.z.copy_(acc)
        st = L.st if L.st is not None else torch.zeros(rows, 2)
        out = torch.zeros(rows, L.N)
        self.ln_fwd(acc, L.gamma, L.beta, out, st, L.act)
      elif L.epi == P.EPI_LNBWD:
        out = torch.zeros(rows, L.N)
        self.ln_bwd(acc, L.z, L.st, L.gamma, L.beta, out, L.dgamma, L.dbeta, L.act)
      elif L.epi == P.EPI_GRU:
        L.z.copy_(acc)
        st = L.st if L.st is not None else torch.zeros(rows, 2)
        y = torch.zeros(rows, L.N)
        self.ln_fwd(acc, L.gamma, L.beta, y, st, 0)
        if L.x1 is not None:
          L.x1.copy_(y)
        out = torch.zeros(rows, L.N // 3)
        self.gru_fwd(y, L.x0, out)
      elif L.epi == P.EPI_STATS:
        S = L.N // L.ipar
        out = torch.zeros(rows, L.N)
        oh = torch.zeros(rows, L.N)
        self.stats_fwd(acc, L.x0, out, oh, None, S, L.ipar, L.f0,
                       2 if L.x1 is None else (1 if L.x0 is None else 0))
        if L.x1 is not None:
          L.x1.copy_(oh)
      elif L.epi == P.EPI_STATSBWD:
        out = torch.zeros(rows, L.N)
        self.stats_bwd(acc, None, L.x0, out, L.N // L.ipar, L.ipar, L.f0)
      elif L.epi == P.EPI_GRUBWD:
        D = L.N
        dy, dprev = torch.zeros(rows, 3 * D), torch.zeros(rows, D)
        self.gru_bwd(acc, L.x1, L.x0, dy, dprev, False)
        if L.x2 is not None:
          L.x2.copy_(L.x2 + dprev)
        out = torch.zeros(rows, 3 * D)
        self.ln_bwd(dy, L.z, L.st, L.gamma, L.beta if L.beta is not None else torch.zeros(3 * D), out,
                    None, None, 0)
      else:
        A = L.ipar
        out, mean, std = torch.zeros(rows, A), torch.zeros(rows, A), torch.zeros(rows, A)
        self.actor_sample(acc[:, :A].contiguous(), acc[:, A:].contiguous(), L.x0.contiguous(),
                          L.f0, L.f1, out, mean, std)
        L.x1.copy_(mean)
        L.x2.copy_(std)
      L.out.copy_(out)
      if L.out_lo is not None:
        L.out_lo.copy_(P.lo_of(out))

  def split_lo(self, x, lo):
    from daydreamer_b200 import program as P
    lo.copy_(P.lo_of(x))

  def transpose_lo(self, p, pt, pt_lo, jobs):
    from daydreamer_b200 import program as P
    for src, rows, cols, dst, ldd, mode in jobs:
      m = p[src:src + rows * cols].view(rows, cols)
      t = (m if mode == 1 else m.t()).contiguous()
      r, c = t.shape
      for buf, val in ((pt, t), (pt_lo, P.lo_of(t))):
        buf[dst:dst + (r - 1) * ldd + c].as_strided((r, c), (ldd, 1)).copy_(val)

  # -- categorical latent ---------------------------------------------------
  def stats_fwd(self, Z, G, logit, stoch, idx, S, C, mix, mode):
    """mode 0: idx = argmax(logit + G); 1: idx = argmax(logit); 2: no draw."""
    M = Z.shape[0]
    z = Z.reshape(M, S, C)
    p = torch.softmax(z, -1)
    lg = torch.log((1.0 - mix) * p + mix / C)
    logit.copy_(lg.reshape(M, S * C))
    if mode == 2:
      return
    key = lg + G.reshape(M, S, C) if mode == 0 else lg
    i = torch.argmax(key, -1)
    stoch.copy_(F.one_hot(i, C).float().reshape(M, S * C))
    if idx is not None:
      idx.copy_(i.to(idx.dtype))

  def stats_bwd(self, dStoch, dLogit, logit, dZ, S, C, mix):
    """Straight-through (grad wrt probs = softmax(logit)) plus direct grad wrt
    the unimixed logit, pulled back to the raw linear output Z."""
    M = logit.shape[0]
    lg = logit.reshape(M, S, C)
    q = torch.exp(lg)
    dl = torch.zeros_like(lg)
    if dStoch is not None:
      ds = dStoch.reshape(M, S, C)
      dl = dl + q * (ds - (q * ds).sum(-1, keepdim=True))
    if dLogit is not None:
      dl = dl + dLogit.reshape(M, S, C)
    dq = dl / q
    p = (q - mix / C) / (1.0 - mix)
    dp = (1.0 - mix) * dq
    dz = p * (dp - (p * dp).sum(-1, keepdim=True))
    dZ.copy_(dz.reshape(M, S * C))

  def kl_fwd(self, post, prior, S, C, bal, kl_row, ent_post, ent_prior):
    M = post.shape[0]
    a = torch.log_softmax(post.reshape(M, S, C), -1)
    b = torch.log_softmax(prior.reshape(M, S, C), -1)
    pa, pb = torch.exp(a), torch.exp(b)
    kl_row.copy_((pa * (a - b)).sum((-1, -2)))        # bal*kl + (1-bal)*kl
    ent_post.copy_(-(pa * a).sum((-1, -2)))
    ent_prior.copy_(-(pb * b).sum((-1, -2)))

  def kl_bwd(self, post, prior, S, C, bal, coef, scale, dpost, dprior):
    """d/dlogits of coef*scale*(bal*KL(sg(post)||prior)+(1-bal)*KL(post||sg(prior)))."""
    M = post.shape[0]
    a = torch.log_softmax(post.reshape(M, S, C), -1)
    b = torch.log_softmax(prior.reshape(M, S, C), -1)
    pa, pb = torch.exp(a), torch.exp(b)
    k = coef * scale.reshape(())
    kl = (pa * (a - b)).sum(-1, keepdim=True)
    dpost.copy_((k * (1.0 - bal) * pa * ((a - b) - kl)).reshape(M, S * C))
    dprior.copy_((k * bal * (pb - pa)).reshape(M, S * C))

  # -- observe masking --------------------------------------------------------
  def mask_fwd(self, pd, ps, act, first, init_d, init_s, od, os_, oa):
    f = first[:, None]
    od.copy_(pd * (1.0 - f) + init_d[None] * f)
    os_.copy_(ps * (1.0 - f) + init_s[None] * f)
    oa.copy_(act * (1.0 - f))

  def mask_bwd(self, dd, ds, first, dpd, dps, dinit):
    f = first[:, None]
    dinit += (dd * f).sum(0)
    if dpd is not None:
      dpd += dd * (1.0 - f)
      dps += ds * (1.0 - f)

  # -- elementwise utils ------------------------------------------------------
  def add_(self, dst, src):
    dst += src

  def copy_(self, dst, src):
    dst.copy_(src)

  def fill_(self, dst, value):
    dst.fill_(value)

  def transpose01(self, src, dst):
    A, B = src.shape[0], src.shape[1]
    dst.view(B, A, -1).copy_(src.reshape(A, B, -1).transpose(0, 1))

  def colsum(self, X, out, acc=False):
    s = X.sum(0)
    out.copy_(out + s if acc else s)

  def tanh_fwd(self, x, y):
    y.copy_(torch.tanh(x))

  def tanh_bwd(self, dy, y, dx, acc=True):
    g = dy * (1.0 - y * y)
    dx.copy_(dx + g if acc else g)

  # -- image path --------------------------------------------------------------
  def im2col(self, X, k, stride, col):
    N, H, W, C = X.shape
    Ho, Wo = (H - k) // stride + 1, (W - k) // stride + 1
    c = col.view(N, Ho, Wo, k, k, C)
    for kh in range(k):
      for kw in range(k):
        c[:, :, :, kh, kw, :] = X[:, kh:kh + stride * (Ho - 1) + 1:stride,
                                  kw:kw + stride * (Wo - 1) + 1:stride, :]

  def col2im(self, col, k, stride, bias, X):
    N, H, W, C = X.shape
    Ho, Wo = (H - k) // stride + 1, (W - k) // stride + 1
    c = col.view(N, Ho, Wo, k, k, C)
    X.zero_()
    if bias is not None:
      X += bias
    for kh in range(k):
      for kw in range(k):
        X[:, kh:kh + stride * (Ho - 1) + 1:stride,
          kw:kw + stride * (Wo - 1) + 1:stride, :] += c[:, :, :, kh, kw, :]

  def image_stage(self, src, dst, c0, scale):
    B, L = src.shape[:2]
    Ck = src.shape[-1]
    v = src.reshape(B, L, -1, Ck).transpose(0, 1).float() * scale
    dst.view(L * B, -1, dst.shape[-1])[:, :, c0:c0 + Ck] = v.reshape(L * B, -1, Ck)

  def image_loss(self, Z, target, c0, c1, coef, loss_row, dZ):
    M = Z.shape[0]
    z = Z.reshape(M, -1, Z.shape[-1])[:, :, c0:c1]
    t = target.reshape(M, -1, Z.shape[-1])[:, :, c0:c1]
    sg = torch.sigmoid(z)
    loss_row.copy_(((sg - t) ** 2).sum((-1, -2)))
    dZ.view(M, -1, Z.shape[-1])[:, :, c0:c1] = coef * 2.0 * (sg - t) * sg * (1.0 - sg)

  def sigmoid(self, x, y):
    y.copy_(torch.sigmoid(x))

  # -- losses -----------------------------------------------------------------
  def head_loss(self, kind, pred, target, weight, coef, loss_row, dpred):
    """kind 0: sum_d (p-t)^2; 1: (p-symlog(t))^2 summed; 2: softplus(p)-t*p.
    loss_row[m] = weight[m]*loss; dpred = coef*weight*dloss/dpred."""
    w = weight[:, None] if weight is not None else 1.0
    if kind == 0:
      diff = pred - target
      loss, g = diff ** 2, 2.0 * diff
    elif kind == 1:
      diff = pred - torch.sign(target) * torch.log1p(torch.abs(target))
      loss, g = diff ** 2, 2.0 * diff
    else:
      loss = F.softplus(pred) - target * pred
      g = torch.sigmoid(pred) - target
    loss_row.copy_((loss * w).sum(-1))
    dpred.copy_(coef * g * w)

  def balance(self, kind, pred, loss_row, target, thres, out):
    """A/tfutils.py:395-411 sums: out f64[8] = [n_pos, n_neg, loss_pos, loss_neg,
    acc_pos, acc_neg, sum_target, sum_pred]."""
    mean = (torch.sign(pred) * torch.expm1(torch.abs(pred)) if kind == 1
            else torch.sigmoid(pred))[:, 0]
    t = target[:, 0]
    pos = (t > thres).double()
    neg = 1.0 - pos
    pr = (mean > thres).double()
    l = loss_row.double()
    out.copy_(torch.stack([
        pos.sum(), neg.sum(), (l * pos).sum(), (l * neg).sum(),
        (pr * pos).sum(), ((1 - pr) * neg).sum(), t.double().sum(),
        mean.double().sum()]))

  # -- imagination ------------------------------------------------------------
  def actor_sample(self, mraw, sraw, eps, lo, hi, action, mean, std):
    mu = torch.tanh(mraw)
    sd = (hi - lo) * torch.sigmoid(sraw) + lo
    mean.copy_(mu)
    std.copy_(sd)
    action.copy_(mu + sd * eps)

  def actor_ent(self, std, lo, hi, ent):
    c = 0.5 * math.log(2 * math.pi * math.e)
    elo = c + math.log(lo)
    ehi = c + math.log(hi)
    ent.copy_((c + torch.log(std) - elo) / (ehi - elo))

  def actor_bwd(self, dA, eps, mean, std, weight, scale, coef, lo, hi, dmraw, dsraw):
    """dA: grad wrt sampled action (through dynamics).  Entropy term:
    loss -= coef * sum_a scale_a * ent_norm * weight."""
    ehi_elo = math.log(hi) - math.log(lo)
    dstd = dA * eps - coef * weight[:, None] * scale[None] / (ehi_elo * std)
    dmraw.copy_(dA * (1.0 - mean * mean))
    sg = (std - lo) / (hi - lo)
    dsraw.copy_(dstd * (hi - lo) * sg * (1.0 - sg))

  def actor_disc_ent(self, logit, ent):
    a = torch.log_softmax(logit, -1)
    ent.reshape(-1).copy_(-(torch.exp(a) * a).sum(-1) / math.log(logit.shape[-1]))

  def actor_disc_loss(self, logit, action, ret, value, adv, ent, scale, weight, coef,
                      loss_row, dlogit):
    """'reinforce' actor loss rows and their gradient wrt the (unimixed) logits."""
    M, A = logit.shape
    a = torch.log_softmax(logit, -1)
    q = torch.exp(a)
    score = ((ret.reshape(-1)[:M] - value.reshape(-1)[:M] - adv[0]) * adv[1])[:, None]
    w = weight.reshape(-1)[:M, None]
    e = ent.reshape(-1)[:M, None]
    logp = (action * a).sum(-1, keepdim=True)
    h = -(q * a).sum(-1, keepdim=True)
    loss_row.copy_(((-logp * score - scale.reshape(-1)[0] * e) * w)[:, 0])
    dlogit.copy_(coef * w * (score * (q * action.sum(-1, keepdim=True) - action)
                             + scale.reshape(-1)[0] / math.log(A) * q * (a + h)))

  def lambda_fwd(self, rho, vraw, kraw, cont0, gamma, lam,
                 reward, value, cont, weight, ret):
    """rho/vraw/kraw [H+1,N] raw head outputs (row 0 of rho/kraw unused)."""
    H = rho.shape[0] - 1
    symexp = lambda x: torch.sign(x) * torch.expm1(torch.abs(x))
    reward.copy_(symexp(rho[1:]))
    value.copy_(symexp(vraw))
    cont[0] = cont0
    cont[1:] = torch.sigmoid(kraw[1:])
    weight.copy_(torch.cumprod(gamma * cont, 0) / gamma)
    disc = gamma * cont[1:]
    last = value[H]
    for t in reversed(range(H)):
      last = reward[t] + disc[t] * ((1.0 - lam) * value[t + 1] + lam * last)
      ret[t] = last

  def lambda_bwd(self, rho, vraw, value, cont, weight, ret, adv, coef, gamma,
                 lam, drho, dvraw, dkraw):
    """Actor loss L = coef * sum_{t<H,n} -(ret - value[t] - mean)*scale*w;
    adv = f32[2] (mean, scale).  Writes d/d raw head outputs."""
    H = rho.shape[0] - 1
    scale = adv[1]
    dret = -coef * scale * weight[:H]
    dV = torch.zeros_like(value)
    dV[:H] = -dret
    dr = torch.zeros_like(rho)
    dd = torch.zeros_like(rho)
    g = torch.zeros_like(ret[0])
    for t in range(1, H + 1):
      # R_{t-1} = r_t + d_t ((1-lam) V_t + lam R_t),  R_H = V_H
      g = dret[t - 1] + (g * gamma * cont[t - 1] * lam if t > 1 else 0.0)
      d_t = gamma * cont[t]
      nxt = ret[t] if t < H else value[H]
      dr[t] = g
      dd[t] = g * ((1.0 - lam) * value[t] + lam * nxt)
      dV[t] = dV[t] + g * d_t * ((1.0 - lam) + (lam if t == H else 0.0))
    drho.copy_(dr * torch.exp(torch.abs(rho)))
    drho[0] = 0.0
    dvraw.copy_(dV * torch.exp(torch.abs(vraw)))
    dkraw.copy_(dd * gamma * cont * (1.0 - cont))
    dkraw[0] = 0.0

  # -- reductions / controllers --------------------------------------------
  def stats_init(self, rows):
    rows.copy_(torch.tensor([0.0, 0.0, float('inf'), float('-inf'), 0.0, 0.0, 0.0, 0.0],
                            dtype=torch.float64).expand_as(rows))

  def stats(self, X, out):
    """Accumulates into out f64[8] = [sum, sumsq, min, max, abssum, absmax, n, 0]
    (rows start from stats_init)."""
    x = X.double().reshape(-1)
    out[0] += x.sum()
    out[1] += (x * x).sum()
    out[2] = torch.minimum(out[2], x.min())
    out[3] = torch.maximum(out[3], x.max())
    out[4] += x.abs().sum()
    out[5] = torch.maximum(out[5], x.abs().max())
    out[6] += float(x.numel())

  def autoadapt(self, scale, sums, denom, target, lo, hi, vel, thres, inverse):
    avg = sums / denom
    below = avg < target / (1.0 + thres)
    above = avg > target * (1.0 + thres)
    if inverse:
      below, above = above, below
    adj = torch.where(above, scale * (1.0 + vel),
                      torch.where(below, scale / (1.0 + vel), scale))
    scale.copy_(torch.clamp(adj, lo, hi))

  def normalize_update(self, state, st, decay, maxv, out):
    """state f64[3] = (mean, sqrs, step); st = stats() of the batch;
    out f32[2] = (mean, scale) of the bias-corrected transform."""
    n = st[6]
    state[2] += 1.0
    state[0] = decay * state[0] + (1.0 - decay) * st[0] / n
    state[1] = decay * state[1] + (1.0 - decay) * st[1] / n
    corr = 1.0 - decay ** state[2]
    mean = state[0] / corr
    var = state[1] / corr - mean ** 2
    out[0] = mean.float()
    out[1] = torch.rsqrt(torch.clamp(var, min=1.0 / maxv ** 2)).float()

  def sumsq(self, g, out):
    out[0] = (g.double() ** 2).sum()

  def adam(self, p, g, m, v, n_decay, lr, b1, b2, eps, wd, clip, gscale,
           sumsq, counters, ci):
    """counters[ci] += 1 happens here (single launch); norm = sqrt(sumsq)*gscale."""
    counters[ci] += 1
    t = float(counters[ci])
    norm = math.sqrt(float(sumsq[0])) * gscale
    s = gscale * (min(1.0, clip / norm) if clip and norm > 0 else 1.0)
    gg = g * s
    p[:n_decay] *= (1.0 - wd * lr)
    m.mul_(b1).add_(gg, alpha=1.0 - b1)
    v.mul_(b2).addcmul_(gg, gg, value=1.0 - b2)
    mhat = m / (1.0 - b1 ** t)
    vhat = v / (1.0 - b2 ** t)
    p -= lr * mhat / (torch.sqrt(vhat) + eps)

  def slow_sync(self, src, dst, counters, ci, every):
    u = int(counters[ci])
    if u == -1 or u >= every:
      dst.copy_(src)
      u = 0
    counters[ci] = u + 1

  # -- noise ------------------------------------------------------------------
  def fill_gumbel(self, buf, seed, stream, counters, ci):
    n = buf.numel()
    buf.reshape(-1).copy_(torch.from_numpy(
        philox.gumbel(n, seed, stream, int(counters[ci]))))

  def fill_normal(self, buf, seed, stream, counters, ci):
    n = buf.numel()
    buf.reshape(-1).copy_(torch.from_numpy(
        philox.normal(n, seed, stream, int(counters[ci]))))

  def counter_inc(self, counters, ci):
    counters[ci] += 1

  def axpy_(self, dst, src, scale):
    dst += scale.reshape(-1)[0] * src

  def sub(self, dst, a, b):
    dst.copy_(a - b)

  def symexp(self, src, dst):
    dst.copy_(torch.sign(src) * torch.expm1(torch.abs(src)))

  def actor_loss(self, ret, value, adv, ent, scale, weight, out):
    """Per-row actor loss (A/agent.py:351-381, continuous 'backprop')."""
    H = ret.shape[0]
    score = (ret - value[:H] - adv[0]) * adv[1]
    e = (ent * scale[None]).sum(-1).reshape(H, -1)
    out.copy_(((-score - e) * weight[:H]).reshape(-1))
