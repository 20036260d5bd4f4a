"""ctypes binding of the sm_100a kernel library (include/dd_kernels.h).

This is the ONLY backend the product constructs.  It fails loudly when the
shared library or the GPU is missing: there is no CPU or PyTorch fallback for
any primitive.  torch is used for device memory and streams only; every method
enqueues hand-written CUDA kernels through the C-ABI on torch's current stream
(so a torch.cuda.graph capture records them).
"""
import contextlib
import ctypes
import os
import pathlib

import numpy as np
import torch

LIB_PATH = pathlib.Path(__file__).resolve().parent / 'lib' / 'libdd_kernels.so'

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_F = ctypes.c_float
_D = ctypes.c_double
_U64 = ctypes.c_uint64
_U32 = ctypes.c_uint32

# name -> argtypes (all return int); mirrors include/dd_kernels.h one to one
SIGNATURES = {
    'dd_gemm': [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _F, _P, _P],
    'dd_set_gemm_tc': [_I],
    'dd_set_gemm_tc_debug': [_I],
    'dd_set_gemm_tc_persistent': [_I],
    'dd_set_gemm_tc_staged': [_I],
    'dd_set_pdl': [_I],
    'dd_set_gemm_thin': [_I],
    'dd_set_ln_warp_max': [_I],
    'dd_set_ln_vec': [_I],
    'dd_ln_fwd': [_P, _I, _P, _P, _P, _I, _P, _I, _I, _I, _P],
    'dd_ln_bwd': [_P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _P],
    'dd_gru_fwd': [_P, _I, _P, _I, _P, _I, _I, _I, _P],
    'dd_gru_bwd': [_P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _P],
    'dd_stats_fwd': [_P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _F, _I, _P],
    'dd_stats_bwd': [_P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _F, _P],
    'dd_kl_fwd': [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P],
    'dd_kl_bwd': [_P, _I, _P, _I, _I, _I, _I, _F, _F, _P, _P, _I, _P, _I, _P],
    'dd_mask_fwd': [_P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    'dd_mask_bwd': [_P, _I, _P, _I, _P, _P, _I, _P, _I, _P, _I, _I, _I, _P],
    'dd_add': [_P, _I, _P, _I, _I, _I, _P],
    'dd_copy': [_P, _I, _P, _I, _I, _I, _P],
    'dd_fill': [_P, _I, _I, _I, _F, _P],
    'dd_fill_f64': [_P, _L, _D, _P],
    'dd_axpy': [_P, _I, _P, _I, _I, _I, _P, _P],
    'dd_sub': [_P, _P, _P, _L, _P],
    'dd_symexp': [_P, _P, _L, _P],
    'dd_transpose01': [_P, _P, _I, _I, _I, _P],
    'dd_colsum': [_P, _I, _I, _I, _P, _I, _P],
    'dd_tanh_fwd': [_P, _P, _L, _P],
    'dd_tanh_bwd': [_P, _P, _P, _L, _I, _P],
    'dd_im2col': [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    'dd_col2im': [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    'dd_image_stage': [_P, _I, _I, _I, _I, _P, _I, _I, _F, _P],
    'dd_image_loss': [_P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P],
    'dd_sigmoid': [_P, _P, _L, _P],
    'dd_head_loss': [_I, _P, _I, _P, _I, _P, _F, _P, _P, _I, _I, _I, _P],
    'dd_balance': [_I, _P, _P, _P, _F, _I, _P, _P],
    'dd_actor_sample': [_P, _P, _P, _F, _F, _P, _P, _P, _L, _P],
    'dd_actor_ent': [_P, _F, _F, _P, _L, _P],
    'dd_actor_loss': [_P, _P, _P, _P, _P, _P, _P, _L, _I, _P],
    'dd_actor_bwd': [_P, _P, _P, _P, _P, _P, _F, _F, _F, _P, _P, _L, _I, _P],
    'dd_actor_disc_ent': [_P, _P, _L, _I, _P],
    'dd_actor_disc_loss': [_P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _L, _I, _P],
    'dd_lambda_fwd': [_P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _I, _I, _P],
    'dd_lambda_bwd': [_P, _P, _P, _P, _P, _P, _P, _F, _F, _F, _P, _P, _P, _I, _I, _P],
    'dd_stats_init': [_P, _I, _P],
    'dd_stats': [_P, _I, _I, _I, _P, _P],
    'dd_autoadapt': [_P, _P, _I, _F, _F, _F, _F, _F, _F, _I, _P],
    'dd_normalize_update': [_P, _P, _D, _D, _P, _P],
    'dd_sumsq': [_P, _L, _P, _P],
    'dd_adam': [_P, _P, _P, _P, _L, _L, _F, _D, _D, _F, _F, _F, _F, _P, _P, _I, _P],
    'dd_slow_sync': [_P, _P, _L, _P, _I, _I, _P],
    'dd_counter_inc': [_P, _I, _P],
    'dd_fill_gumbel': [_P, _L, _U64, _U32, _P, _I, _P],
    'dd_fill_normal': [_P, _L, _U64, _U32, _P, _I, _P],
    'dd_program_run': [_P, _I, _I, _P],
    'dd_set_machine_debug': [_I],
    'dd_set_machine_ks': [_I],
    'dd_machine_prof': [_I, _P, _I],
    'dd_split_lo': [_P, _I, _P, _I, _I, _I, _P],
    'dd_transpose_lo': [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P],
}


class DdObserveFwd(ctypes.Structure):
  """Mirror of `struct DdObserveFwd` in include/dd_kernels.h."""
  _PTRS = ('action', 'first', 'first0', 'prev_deter', 'prev_stoch', 'init_deter', 'init_stoch',
           'g_obs', 'Win', 'ln_in_s', 'ln_in_b', 'Wgru', 'ln_gru_s', 'ln_gru_b', 'Wobs',
           'ln_obs_s', 'ln_obs_b', 'Wst', 'bst', 'm_deter', 'm_stoch', 'm_act', 'z_in', 'x_in',
           'st_in', 'z_gru', 'y_gru', 'st_gru', 'feat', 'z_obs', 'h_obs', 'st_obs', 'post_logit',
           'post_idx', 'partial', 'barrier', 'timing')
  _fields_ = ([(n, ctypes.c_int) for n in ('B', 'L', 'D', 'U', 'S', 'C', 'A')] +
              [('mix', ctypes.c_float)] + [(n, ctypes.c_void_p) for n in _PTRS])


class DdObserveBwd(ctypes.Structure):
  """Mirror of `struct DdObserveBwd` in include/dd_kernels.h."""
  _PTRS = ('first', 'first0', 'post_logit', 'dpost_logit', 'z_obs', 'st_obs', 'ln_obs_s',
           'ln_obs_b', 'y_gru', 'z_gru', 'st_gru', 'ln_gru_s', 'ln_gru_b', 'm_deter', 'z_in',
           'st_in', 'ln_in_s', 'ln_in_b', 'Wst', 'Wobs', 'Wgru', 'Win', 'dfeat', 'dzs', 'dz_obs',
           'dz_gru', 'dz_in', 'dh', 'dy_gru', 'dx', 'd_initdeter', 'partial', 'barrier', 'timing')
  _fields_ = ([(n, ctypes.c_int) for n in ('B', 'L', 'D', 'U', 'S', 'C', 'nsc', 'xk')] +
              [('mix', ctypes.c_float)] + [(n, ctypes.c_void_p) for n in _PTRS])


class DdSeg(ctypes.Structure):
  """Mirror of `struct DdSeg` in include/dd_kernels.h."""
  _fields_ = [('a', _P), ('a_lo', _P), ('lda', _I), ('k', _I), ('wk0', _I), ('lda_lo', _I)]


class DdLayer(ctypes.Structure):
  """Mirror of `struct DdLayer` in include/dd_kernels.h."""
  _fields_ = [('nseg', _I), ('seg', DdSeg * 2), ('w', _P), ('w_lo', _P), ('ldw', _I),
              ('wrows', _I), ('wcols', _I), ('wn0', _I), ('N', _I), ('epi', _I), ('act', _I),
              ('add', _P), ('ld_add', _I), ('bias', _P), ('gamma', _P), ('beta', _P), ('z', _P),
              ('ld_z', _I), ('st', _P), ('out', _P), ('out_lo', _P), ('ld_out', _I), ('ld_lo', _I),
              ('dgamma', _P), ('dbeta', _P), ('x', _P * 3), ('ldx', _I * 3), ('f', _F * 2),
              ('ipar', _I)]


class ProgramHandle:
  """A compiled layer program: the device table with its TMA descriptors plus references that
  keep every tensor the descriptors point at alive."""

  def __init__(self, table, nlayers, rows, layers):
    self.table, self.nlayers, self.rows, self.layers = table, nlayers, rows, layers


def load_library(path=LIB_PATH):
  """dlopen the kernel library and type every entry point.  Raises if absent."""
  path = pathlib.Path(path)
  if not path.exists():
    raise RuntimeError(
        f'{path} is missing: build it with `python -m daydreamer_b200.build` '
        '(nvcc, sm_100a).  There is no CPU fallback.')
  lib = ctypes.CDLL(str(path))
  for name, argtypes in SIGNATURES.items():
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = ctypes.c_int
  lib.dd_observe_fwd_grid.argtypes = [ctypes.POINTER(DdObserveFwd)]
  lib.dd_observe_fwd_grid.restype = ctypes.c_int
  lib.dd_observe_fwd.argtypes = [ctypes.POINTER(DdObserveFwd), _P]
  lib.dd_observe_fwd.restype = ctypes.c_int
  lib.dd_observe_bwd_grid.argtypes = [ctypes.POINTER(DdObserveBwd)]
  lib.dd_observe_bwd_grid.restype = ctypes.c_int
  lib.dd_observe_bwd.argtypes = [ctypes.POINTER(DdObserveBwd), _P]
  lib.dd_observe_bwd.restype = ctypes.c_int
  lib.dd_program_bytes.argtypes = [_I]
  lib.dd_program_bytes.restype = ctypes.c_size_t
  lib.dd_program_build.argtypes = [ctypes.POINTER(DdLayer), _I, _I, _P]
  lib.dd_program_build.restype = ctypes.c_int
  lib.dd_program_deps.argtypes = [ctypes.POINTER(DdLayer), _I, _I, ctypes.POINTER(ctypes.c_int)]
  lib.dd_program_deps.restype = ctypes.c_int
  lib.dd_program_debug.argtypes = [ctypes.POINTER(ctypes.c_uint)]
  lib.dd_program_debug.restype = ctypes.c_int
  lib.dd_version.restype = ctypes.c_int
  lib.dd_last_error.restype = ctypes.c_char_p
  lib.dd_launch_count.restype = ctypes.c_int64
  return lib


def _mat(t):
  """(ptr, rows, cols, ld) of a 1-D / 2-D (row-strided) / contiguous N-D tensor."""
  if t.dim() == 1:
    assert t.stride(0) == 1 or t.shape[0] <= 1
    return t.data_ptr(), 1, t.shape[0], t.shape[0]
  if t.dim() == 2:
    assert t.stride(1) == 1 or t.shape[1] == 1, t.stride()
    return t.data_ptr(), t.shape[0], t.shape[1], t.stride(0)
  assert t.is_contiguous(), 'N-D views must be contiguous'
  c = t.shape[-1]
  return t.data_ptr(), t.numel() // c, c, c


def _ptr(t):
  return None if t is None else t.data_ptr()


def _flat(t):
  assert t.is_contiguous()
  return t.data_ptr(), t.numel()


def encode_layers(layers, rows):
  """program.Layer list -> the DdLayer array of the C-ABI (pointers and strides of torch views)."""
  n = len(layers)
  arr = (DdLayer * n)()
  for i, L in enumerate(layers):
    d = arr[i]
    d.nseg = len(L.segs)
    assert 0 <= d.nseg <= 2
    for g, sg in enumerate(L.segs):
      pa, r, k, lda = _mat(sg.a)
      assert r == rows, (sg.a.shape, rows)
      d.seg[g].a, d.seg[g].lda, d.seg[g].k, d.seg[g].wk0 = pa, lda, k, int(sg.wk0)
      if sg.a_lo is not None:
        assert sg.a_lo.shape == sg.a.shape
        d.seg[g].a_lo, _, _, d.seg[g].lda_lo = _mat(sg.a_lo)
    if L.segs:
      pw, wrows, wcols, ldw = _mat(L.w)
      pwl, _, _, ldwl = _mat(L.w_lo)
      assert ldwl == ldw and L.w_lo.shape == L.w.shape
      d.w, d.w_lo, d.ldw, d.wrows, d.wcols, d.wn0 = pw, pwl, ldw, wrows, wcols, int(L.wn0)
    d.N, d.epi, d.act = int(L.N), int(L.epi), int(L.act)
    if L.add is not None:
      d.add, _, _, d.ld_add = _mat(L.add)
    d.bias, d.gamma, d.beta = _ptr(L.bias), _ptr(L.gamma), _ptr(L.beta)
    if L.z is not None:
      d.z, _, _, d.ld_z = _mat(L.z)
    d.st = _ptr(L.st)
    d.out, _, _, d.ld_out = _mat(L.out)
    if L.out_lo is not None:
      d.out_lo, _, _, d.ld_lo = _mat(L.out_lo)
    d.dgamma, d.dbeta = _ptr(L.dgamma), _ptr(L.dbeta)
    for q, t in enumerate((L.x0, L.x1, L.x2)):
      if t is not None:
        d.x[q], _, _, d.ldx[q] = _mat(t)
    d.f[0], d.f[1], d.ipar = float(L.f0), float(L.f1), int(L.ipar)
  return arr


def program_deps(lib, layers, rows):
  """[(dep of segment 0, dep of segment 1)] per layer: how many leading layers of the program write
  the segment (dd_program_deps; host only, works on CPU tensors)."""
  n = len(layers)
  out = (ctypes.c_int * (2 * n))()
  rc = lib.dd_program_deps(encode_layers(layers, rows), n, int(rows), out)
  if rc != 0:
    raise RuntimeError(lib.dd_last_error().decode())
  return [(out[2 * i], out[2 * i + 1]) for i in range(n)]


class CudaBackend:
  name = 'cuda-sm100a'

  def __init__(self, device=None):
    if not torch.cuda.is_available():
      raise RuntimeError(
          'daydreamer_b200 needs a CUDA device (B200, sm_100a); none is visible '
          'and there is no CPU fallback.')
    self.lib = load_library()
    self.device = torch.device(device if device is not None else
                               f'cuda:{torch.cuda.current_device()}')
    torch.cuda.set_device(self.device)
    if os.environ.get('DD_PDL', '') != '':
      self.lib.dd_set_pdl(int(os.environ['DD_PDL']))
    if os.environ.get('DD_LN_VEC', '') != '':
      self.lib.dd_set_ln_vec(int(os.environ['DD_LN_VEC']))
    if os.environ.get('DD_LN_WARP_MAX', '') != '':
      self.lib.dd_set_ln_warp_max(int(os.environ['DD_LN_WARP_MAX']))
    if os.environ.get('DD_GEMM_THIN', '') != '':
      self.lib.dd_set_gemm_thin(int(os.environ['DD_GEMM_THIN']))
    if os.environ.get('DD_GEMM_TC_STAGED', '') != '':
      self.lib.dd_set_gemm_tc_staged(int(os.environ['DD_GEMM_TC_STAGED']))
    if os.environ.get('DD_GEMM_TC_PERSISTENT', '') != '':
      self.lib.dd_set_gemm_tc_persistent(int(os.environ['DD_GEMM_TC_PERSISTENT']))
    if os.environ.get('DD_MACHINE_DEBUG', '') != '':        # dd_machine.cu F_* switches (A/B runs)
      self.lib.dd_set_machine_debug(int(os.environ['DD_MACHINE_DEBUG']))

  def set_pdl(self, on):
    return self.lib.dd_set_pdl(int(on))

  # -- plumbing ---------------------------------------------------------------
  @property
  def stream(self):
    return torch.cuda.current_stream(self.device).cuda_stream

  def _ck(self, rc):
    if rc != 0:
      raise RuntimeError(self.lib.dd_last_error().decode())

  @contextlib.contextmanager
  def branch(self):
    """Launches issued inside run on a second stream, ordered after everything issued so far on
    the current one; `join()` orders the current stream after them.  Under stream capture both
    become edges of the graph, so the two branches replay concurrently."""
    cur = torch.cuda.current_stream(self.device)
    if getattr(self, '_side', None) is None:
      self._side = torch.cuda.Stream(device=self.device)      # default (lowest) priority
    self._side.wait_stream(cur)
    with torch.cuda.stream(self._side):
      yield

  def join(self):
    torch.cuda.current_stream(self.device).wait_stream(self._side)

  def set_gemm_tc(self, on):
    return self.lib.dd_set_gemm_tc(int(on))

  def launch_count(self):
    return int(self.lib.dd_launch_count())

  def zeros(self, *shape, dtype=torch.float32):
    return torch.zeros(*shape, dtype=dtype, device=self.device)

  def upload(self, array, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(array)).to(self.device, dtype)

  def sync(self):
    torch.cuda.synchronize(self.device)

  # -- dense ----------------------------------------------------------------
  def gemm(self, A, B, C, ta=False, tb=False, beta=0.0, bias=None):
    pa, ra, ca, lda = _mat(A)
    pb, rb, cb, ldb = _mat(B)
    pc, M, N, ldc = _mat(C)
    K = ra if ta else ca
    assert (ca if ta else ra) == M and (cb if tb else rb) == K and (rb if tb else cb) == N, (
        A.shape, B.shape, C.shape, ta, tb)
    self._ck(self.lib.dd_gemm(pa, lda, int(ta), pb, ldb, int(tb), pc, ldc, M, N, K,
                              float(beta), _ptr(bias), self.stream))

  def ln_fwd(self, Z, scale, bias, H, stats, act):
    pz, M, N, ldz = _mat(Z)
    ph, _, _, ldh = _mat(H)
    self._ck(self.lib.dd_ln_fwd(pz, ldz, scale.data_ptr(), bias.data_ptr(), ph, ldh,
                                stats.data_ptr(), M, N, int(act), self.stream))

  def ln_bwd(self, dH, Z, stats, scale, bias, dZ, dscale, dbias, act):
    pd, M, N, ldd = _mat(dH)
    pz, _, _, ldz = _mat(Z)
    po, _, _, ldo = _mat(dZ)
    self._ck(self.lib.dd_ln_bwd(pd, ldd, pz, ldz, stats.data_ptr(), scale.data_ptr(),
                                bias.data_ptr(), po, ldo, _ptr(dscale), _ptr(dbias), M, N,
                                int(act), self.stream))

  def gru_fwd(self, Y, prev, out):
    py, M, _, ldy = _mat(Y)
    pp, _, D, ldp = _mat(prev)
    po, _, _, ldo = _mat(out)
    self._ck(self.lib.dd_gru_fwd(py, ldy, pp, ldp, po, ldo, M, D, self.stream))

  def gru_bwd(self, dOut, Y, prev, dY, dPrev, acc):
    pg, M, D, ldg = _mat(dOut)
    py, _, _, ldy = _mat(Y)
    pp, _, _, ldp = _mat(prev)
    pdy, _, _, lddy = _mat(dY)
    pdp, _, _, lddp = _mat(dPrev)
    self._ck(self.lib.dd_gru_bwd(pg, ldg, py, ldy, pp, ldp, pdy, lddy, pdp, lddp, M, D,
                                 int(acc), self.stream))

  # -- categorical latent ---------------------------------------------------
  def stats_fwd(self, Z, G, logit, stoch, idx, S, C, mix, mode):
    pz, M, _, ldz = _mat(Z)
    pg, ldg = (None, 0) if G is None else (_mat(G)[0], _mat(G)[3])
    pl, _, _, ldl = _mat(logit)
    ps, lds = (None, 0) if stoch is None else (_mat(stoch)[0], _mat(stoch)[3])
    self._ck(self.lib.dd_stats_fwd(pz, ldz, pg, ldg, pl, ldl, ps, lds, _ptr(idx), M, S, C,
                                   float(mix), int(mode), self.stream))

  def stats_bwd(self, dStoch, dLogit, logit, dZ, S, C, mix):
    pds, ldds = (None, 0) if dStoch is None else (_mat(dStoch)[0], _mat(dStoch)[3])
    pdl, lddl = (None, 0) if dLogit is None else (_mat(dLogit)[0], _mat(dLogit)[3])
    pl, M, _, ldl = _mat(logit)
    pz, _, _, ldz = _mat(dZ)
    self._ck(self.lib.dd_stats_bwd(pds, ldds, pdl, lddl, pl, ldl, pz, ldz, M, S, C,
                                   float(mix), self.stream))

  def kl_fwd(self, post, prior, S, C, bal, kl_row, ent_post, ent_prior):
    pp, M, _, ldp = _mat(post)
    pq, _, _, ldq = _mat(prior)
    self._ck(self.lib.dd_kl_fwd(pp, ldp, pq, ldq, M, S, C, kl_row.data_ptr(),
                                ent_post.data_ptr(), ent_prior.data_ptr(), self.stream))

  def kl_bwd(self, post, prior, S, C, bal, coef, scale, dpost, dprior):
    pp, M, _, ldp = _mat(post)
    pq, _, _, ldq = _mat(prior)
    pdp, _, _, lddp = _mat(dpost)
    pdq, _, _, lddq = _mat(dprior)
    self._ck(self.lib.dd_kl_bwd(pp, ldp, pq, ldq, M, S, C, float(bal), float(coef),
                                scale.data_ptr(), pdp, lddp, pdq, lddq, self.stream))

  # -- fused observe scan -------------------------------------------------------
  def _observe_params(self, dims, mix, tensors):
    p = DdObserveFwd()
    for k, v in dims.items():
      setattr(p, k, int(v))
    p.mix = float(mix)
    for name in DdObserveFwd._PTRS:
      t = tensors.get(name)
      if t is None:
        continue                      # optional pointer (timing)
      assert t.is_contiguous(), name
      setattr(p, name, t.data_ptr())
    return p

  def observe_fwd_grid(self, dims, mix, tensors):
    return int(self.lib.dd_observe_fwd_grid(ctypes.byref(self._observe_params(dims, mix, tensors))))

  def observe_fwd(self, dims, mix, tensors):
    p = self._observe_params(dims, mix, tensors)
    self._ck(self.lib.dd_observe_fwd(ctypes.byref(p), self.stream))

  def _observe_bwd_params(self, dims, mix, tensors):
    p = DdObserveBwd()
    for k, v in dims.items():
      setattr(p, k, int(v))
    p.mix = float(mix)
    for name in DdObserveBwd._PTRS:
      t = tensors.get(name)
      if t is None:
        continue                      # optional pointer (timing)
      assert t.is_contiguous(), name
      setattr(p, name, t.data_ptr())
    return p

  def observe_bwd_grid(self, dims, mix, tensors):
    return int(self.lib.dd_observe_bwd_grid(ctypes.byref(self._observe_bwd_params(dims, mix, tensors))))

  def observe_bwd(self, dims, mix, tensors):
    p = self._observe_bwd_params(dims, mix, tensors)
    self._ck(self.lib.dd_observe_bwd(ctypes.byref(p), self.stream))

  # -- observe masking --------------------------------------------------------
  def mask_fwd(self, pd, ps, act, first, init_d, init_s, od, os_, oa):
    ppd, M, D, ldpd = _mat(pd)
    pps, _, SC, ldps = _mat(ps)
    pa, _, A, lda = _mat(act)
    pod, _, _, ldod = _mat(od)
    pos, _, _, ldos = _mat(os_)
    poa, _, _, ldoa = _mat(oa)
    self._ck(self.lib.dd_mask_fwd(ppd, ldpd, pps, ldps, pa, lda, first.data_ptr(),
                                  init_d.data_ptr(), init_s.data_ptr(), pod, ldod, pos, ldos,
                                  poa, ldoa, M, D, SC, A, self.stream))

  def mask_bwd(self, dd, ds, first, dpd, dps, dinit):
    pdd, M, D, lddd = _mat(dd)
    pds, _, SC, ldds = _mat(ds)
    pdpd, lddpd = (None, 0) if dpd is None else (_mat(dpd)[0], _mat(dpd)[3])
    pdps, lddps = (None, 0) if dps is None else (_mat(dps)[0], _mat(dps)[3])
    self._ck(self.lib.dd_mask_bwd(pdd, lddd, pds, ldds, first.data_ptr(), pdpd, lddpd, pdps,
                                  lddps, dinit.data_ptr(), M, D, SC, self.stream))

  # -- elementwise utils ------------------------------------------------------
  def add_(self, dst, src):
    pd, r, c, ldd = _mat(dst)
    ps, _, _, lds = _mat(src)
    self._ck(self.lib.dd_add(pd, ldd, ps, lds, r, c, self.stream))

  def copy_(self, dst, src):
    pd, r, c, ldd = _mat(dst)
    ps, r2, c2, lds = _mat(src)
    assert r * c == r2 * c2, (dst.shape, src.shape)
    if (r, c) != (r2, c2):          # contiguous reshape on one side
      assert ldd == c and lds == c2
      r, c, ldd, lds = 1, r * c, r * c, r * c
    self._ck(self.lib.dd_copy(pd, ldd, ps, lds, r, c, self.stream))

  def fill_(self, dst, value):
    if dst.dtype == torch.float64:
      p, n = _flat(dst)
      self._ck(self.lib.dd_fill_f64(p, n, float(value), self.stream))
      return
    pd, r, c, ldd = _mat(dst)
    self._ck(self.lib.dd_fill(pd, ldd, r, c, float(value), self.stream))

  def axpy_(self, dst, src, scale):
    pd, r, c, ldd = _mat(dst)
    ps, _, _, lds = _mat(src)
    self._ck(self.lib.dd_axpy(pd, ldd, ps, lds, r, c, scale.data_ptr(), self.stream))

  def sub(self, dst, a, b):
    p, n = _flat(dst)
    self._ck(self.lib.dd_sub(p, _flat(a)[0], _flat(b)[0], n, self.stream))

  def symexp(self, src, dst):
    p, n = _flat(src)
    self._ck(self.lib.dd_symexp(p, _flat(dst)[0], n, self.stream))

  def transpose01(self, src, dst):
    A, B = src.shape[0], src.shape[1]
    F = src.numel() // (A * B)
    self._ck(self.lib.dd_transpose01(_flat(src)[0], _flat(dst)[0], A, B, F, self.stream))

  def colsum(self, X, out, acc=False):
    px, M, N, ldx = _mat(X)
    self._ck(self.lib.dd_colsum(px, ldx, M, N, out.data_ptr(), int(acc), self.stream))

  def tanh_fwd(self, x, y):
    p, n = _flat(x)
    self._ck(self.lib.dd_tanh_fwd(p, _flat(y)[0], n, self.stream))

  def tanh_bwd(self, dy, y, dx, acc=True):
    p, n = _flat(dy)
    self._ck(self.lib.dd_tanh_bwd(p, _flat(y)[0], _flat(dx)[0], n, int(acc), self.stream))

  # -- image path --------------------------------------------------------------
  def im2col(self, X, k, stride, col):
    N, H, W, C = X.shape
    assert X.is_contiguous() and col.is_contiguous()
    Ho, Wo = (H - k) // stride + 1, (W - k) // stride + 1
    assert col.numel() == N * Ho * Wo * k * k * C, (X.shape, col.shape)
    self._ck(self.lib.dd_im2col(X.data_ptr(), N, H, W, C, k, stride, col.data_ptr(), self.stream))

  def col2im(self, col, k, stride, bias, X):
    N, H, W, C = X.shape
    assert X.is_contiguous() and col.is_contiguous()
    Ho, Wo = (H - k) // stride + 1, (W - k) // stride + 1
    assert col.numel() == N * Ho * Wo * k * k * C, (X.shape, col.shape)
    self._ck(self.lib.dd_col2im(col.data_ptr(), N, H, W, C, k, stride, _ptr(bias), X.data_ptr(),
                                self.stream))

  def image_stage(self, src, dst, c0, scale):
    B, L = src.shape[:2]
    Ck = src.shape[-1]
    P = src.numel() // (B * L * Ck)
    assert src.dtype == torch.uint8 and src.is_contiguous() and dst.is_contiguous()
    assert dst.numel() == B * L * P * dst.shape[-1]
    self._ck(self.lib.dd_image_stage(src.data_ptr(), B, L, P, Ck, dst.data_ptr(), dst.shape[-1],
                                     int(c0), float(scale), self.stream))

  def image_loss(self, Z, target, c0, c1, coef, loss_row, dZ):
    M, Ctot = Z.shape[0], Z.shape[-1]
    P = Z.numel() // (M * Ctot)
    assert Z.is_contiguous() and target.is_contiguous() and dZ.is_contiguous()
    self._ck(self.lib.dd_image_loss(Z.data_ptr(), target.data_ptr(), M, P, Ctot, int(c0), int(c1),
                                    float(coef), loss_row.data_ptr(), dZ.data_ptr(), self.stream))

  def sigmoid(self, x, y):
    p, n = _flat(x)
    self._ck(self.lib.dd_sigmoid(p, _flat(y)[0], n, self.stream))

  # -- losses -----------------------------------------------------------------
  def head_loss(self, kind, pred, target, weight, coef, loss_row, dpred):
    pp, M, dim, ldp = _mat(pred)
    pt, _, _, ldt = _mat(target)
    pdp, _, _, lddp = _mat(dpred)
    self._ck(self.lib.dd_head_loss(int(kind), pp, ldp, pt, ldt, _ptr(weight), float(coef),
                                   loss_row.data_ptr(), pdp, lddp, M, dim, self.stream))

  def balance(self, kind, pred, loss_row, target, thres, out):
    self._ck(self.lib.dd_balance(int(kind), _flat(pred)[0], loss_row.data_ptr(),
                                 _flat(target)[0], float(thres), pred.shape[0],
                                 out.data_ptr(), self.stream))

  # -- imagination ------------------------------------------------------------
  def actor_sample(self, mraw, sraw, eps, lo, hi, action, mean, std):
    p, n = _flat(mraw)
    self._ck(self.lib.dd_actor_sample(p, _flat(sraw)[0], _flat(eps)[0], float(lo), float(hi),
                                      _flat(action)[0], _flat(mean)[0], _flat(std)[0], n,
                                      self.stream))

  def actor_ent(self, std, lo, hi, ent):
    p, n = _flat(std)
    self._ck(self.lib.dd_actor_ent(p, float(lo), float(hi), _flat(ent)[0], n, self.stream))

  def actor_loss(self, ret, value, adv, ent, scale, weight, out):
    M, A = ent.shape
    self._ck(self.lib.dd_actor_loss(_flat(ret)[0], _flat(value)[0], adv.data_ptr(),
                                    _flat(ent)[0], scale.data_ptr(), _flat(weight)[0],
                                    out.data_ptr(), M, A, self.stream))

  def actor_bwd(self, dA, eps, mean, std, weight, scale, coef, lo, hi, dmraw, dsraw):
    M, A = dA.shape
    self._ck(self.lib.dd_actor_bwd(_flat(dA)[0], _flat(eps)[0], _flat(mean)[0], _flat(std)[0],
                                   _flat(weight)[0], scale.data_ptr(), float(coef), float(lo),
                                   float(hi), _flat(dmraw)[0], _flat(dsraw)[0], M, A,
                                   self.stream))

  def actor_disc_ent(self, logit, ent):
    M, A = logit.shape
    self._ck(self.lib.dd_actor_disc_ent(_flat(logit)[0], _flat(ent)[0], M, A, self.stream))

  def actor_disc_loss(self, logit, action, ret, value, adv, ent, scale, weight, coef,
                      loss_row, dlogit):
    M, A = logit.shape
    self._ck(self.lib.dd_actor_disc_loss(
        _flat(logit)[0], _flat(action)[0], _flat(ret)[0], _flat(value)[0], adv.data_ptr(),
        _flat(ent)[0], scale.data_ptr(), _flat(weight)[0], float(coef), loss_row.data_ptr(),
        _flat(dlogit)[0], M, A, self.stream))

  def lambda_fwd(self, rho, vraw, kraw, cont0, gamma, lam, reward, value, cont, weight, ret):
    H, N = rho.shape[0] - 1, rho.shape[1]
    self._ck(self.lib.dd_lambda_fwd(
        _flat(rho)[0], _flat(vraw)[0], _flat(kraw)[0], cont0.data_ptr(), float(gamma),
        float(lam), _flat(reward)[0], _flat(value)[0], _flat(cont)[0], _flat(weight)[0],
        _flat(ret)[0], H, N, self.stream))

  def lambda_bwd(self, rho, vraw, value, cont, weight, ret, adv, coef, gamma, lam,
                 drho, dvraw, dkraw):
    H, N = rho.shape[0] - 1, rho.shape[1]
    self._ck(self.lib.dd_lambda_bwd(
        _flat(rho)[0], _flat(vraw)[0], _flat(value)[0], _flat(cont)[0], _flat(weight)[0],
        _flat(ret)[0], adv.data_ptr(), float(coef), float(gamma), float(lam), _flat(drho)[0],
        _flat(dvraw)[0], _flat(dkraw)[0], H, N, self.stream))

  # -- reductions / controllers --------------------------------------------
  def stats_init(self, rows):
    self._ck(self.lib.dd_stats_init(rows.data_ptr(), rows.shape[0], self.stream))

  def stats(self, X, out):
    px, r, c, ldx = _mat(X)
    self._ck(self.lib.dd_stats(px, ldx, r, c, out.data_ptr(), self.stream))

  def autoadapt(self, scale, sums, denom, target, lo, hi, vel, thres, inverse):
    self._ck(self.lib.dd_autoadapt(scale.data_ptr(), sums.data_ptr(), scale.numel(),
                                   float(denom), float(target), float(lo), float(hi),
                                   float(vel), float(thres), int(inverse), self.stream))

  def normalize_update(self, state, st, decay, maxv, out):
    self._ck(self.lib.dd_normalize_update(state.data_ptr(), st.data_ptr(), float(decay),
                                          float(maxv), out.data_ptr(), self.stream))

  def sumsq(self, g, out):
    p, n = _flat(g)
    self._ck(self.lib.dd_sumsq(p, n, out.data_ptr(), self.stream))

  def adam(self, p, g, m, v, n_decay, lr, b1, b2, eps, wd, clip, gscale, sumsq, counters, ci):
    self._ck(self.lib.dd_adam(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                              p.numel(), int(n_decay), float(lr), float(b1), float(b2),
                              float(eps), float(wd), float(clip), float(gscale),
                              sumsq.data_ptr(), counters.data_ptr(), int(ci), self.stream))

  def slow_sync(self, src, dst, counters, ci, every):
    assert src.numel() == dst.numel()
    self._ck(self.lib.dd_slow_sync(src.data_ptr(), dst.data_ptr(), src.numel(),
                                   counters.data_ptr(), int(ci), int(every), self.stream))

  def counter_inc(self, counters, ci):
    self._ck(self.lib.dd_counter_inc(counters.data_ptr(), int(ci), self.stream))

  # -- layer machine -------------------------------------------------------------
  def program_build(self, layers, rows):
    """Compile a list of program.Layer into a device table (TMA descriptors are encoded here:
    synchronous, call outside stream capture)."""
    n = len(layers)
    arr = encode_layers(layers, rows)
    nbytes = int(self.lib.dd_program_bytes(n))
    table = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
    torch.cuda.synchronize(self.device)
    self._ck(self.lib.dd_program_build(arr, n, int(rows), table.data_ptr()))
    return ProgramHandle(table, n, int(rows), list(layers))

  def program_run(self, h):
    self._ck(self.lib.dd_program_run(h.table.data_ptr(), h.nlayers, h.rows, self.stream))

  def program_debug(self):
    """Diagnostic words of a stuck wait inside the layer machine (after a failed sync)."""
    out = (ctypes.c_uint * 8)()
    stuck = self.lib.dd_program_debug(out)
    return bool(stuck), [int(x) for x in out]

  def split_lo(self, x, lo):
    px, r, c, ldx = _mat(x)
    pl, _, _, ldl = _mat(lo)
    self._ck(self.lib.dd_split_lo(px, ldx, pl, ldl, r, c, self.stream))

  def transpose_lo(self, p, pt, pt_lo, jobs):
    """jobs: list of (src offset, rows, cols, dst offset, dst row stride, mode); mode 0 writes the
    transpose of the [rows, cols] matrix at p[src:], mode 1 a re-strided copy; both with x_lo."""
    for i in range(0, len(jobs), 64):
      part = jobs[i:i + 64]
      n = len(part)
      cols = [(ctypes.c_int * n)(*[int(j[q]) for j in part]) for q in range(6)]
      self._ck(self.lib.dd_transpose_lo(p.data_ptr(), pt.data_ptr(), pt_lo.data_ptr(), n,
                                        *cols, self.stream))

  # -- noise ------------------------------------------------------------------
  def fill_gumbel(self, buf, seed, stream_id, counters, ci):
    p, n = _flat(buf)
    self._ck(self.lib.dd_fill_gumbel(p, n, int(seed), int(stream_id), counters.data_ptr(),
                                     int(ci), self.stream))

  def fill_normal(self, buf, seed, stream_id, counters, ci):
    p, n = _flat(buf)
    self._ck(self.lib.dd_fill_normal(p, n, int(seed), int(stream_id), counters.data_ptr(),
                                     int(ci), self.stream))
