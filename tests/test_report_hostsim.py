"""Host logic check (no GPU): the image summaries of report() -- context reconstruction,
open-loop prediction along recorded actions, imagined rollout with the policy -- on the
torch-CPU primitive simulator against the oracle restatement of
WorldModel.report (A/agent.py:266-282) and Greedy.report (A/behaviors.py:32-46)."""
import numpy as np
import pytest
import torch

from oracle import dreamer_oracle as O
from daydreamer_b200 import engine as E
from tests import util
from tests.hostsim import HostSim


def run_report(be, tol):
  ospec = O.tiny_spec(discrete=True, act_dim=5, cnn_keys={'image': (18, 18, 3), 'depth': (18, 18, 1)},
                      cnn_depth=4, enc_kernels=(4, 4), dec_kernels=(3, 4, 4))
  spec = util.to_engine_spec(ospec)
  B, L, C5 = 3, 8, 5
  T2, H, A = L - C5, ospec.horizon, ospec.act_dim
  params = O.init_params(ospec, seed=1)
  main = E.Engine(spec, be, B, L)
  main.load_params(params)
  rep = E.VideoReport(main, B, L)
  batch = util.make_batch(ospec, B, L, seed=3)
  rng = np.random.default_rng(7)
  gum = lambda *s: (-np.log(-np.log(rng.uniform(1e-6, 1 - 1e-6, s)))).astype(np.float32)
  g_ctx = gum(C5, B, ospec.stoch, ospec.classes)
  g_ol = gum(T2, B, ospec.stoch, ospec.classes)
  g_im = gum(H, B, ospec.stoch, ospec.classes)
  g_act = gum(H + 1, B, A)

  # ---- oracle -------------------------------------------------------------
  learner = O.Learner(ospec, params)
  p = learner.params
  data = learner.preprocess(batch)
  with torch.no_grad():
    embed = O.encoder(data, p, ospec)
    post, _ = O.observe(embed[:, :C5], data['action'][:, :C5], data['is_first'][:, :C5], None, p,
                        ospec, torch.tensor(g_ctx))
    recon = O.image_decoder(O.feat_of(post, ospec).reshape(-1, ospec.feat), p, ospec)
    recon = recon.reshape((B, C5) + recon.shape[1:])
    state = {k: post[k][:, -1] for k in ('deter', 'stoch', 'logit')}
    frames = []
    for t in range(T2):
      nxt = O.img_step(state, data['action'][:, C5 + t], p, ospec, torch.tensor(g_ol[t]))
      state = {k: nxt[k] for k in ('deter', 'stoch', 'logit')}
      frames.append(O.image_decoder(O.feat_of(state, ospec), p, ospec))
    openl = torch.stack(frames, 1)
    model = torch.cat([recon, openl], 1)
    start = {k: post[k][:, -1] for k in ('deter', 'stoch', 'logit')}
    traj = learner.imagine(p, start, torch.ones(B), torch.tensor(g_im), torch.tensor(g_act))
    imag = O.image_decoder(traj['feat'].reshape(-1, ospec.feat), p, ospec)
    imag = imag.reshape((H + 1, B) + imag.shape[1:]).transpose(0, 1)

  # ---- product logic on the simulator ----------------------------------------
  ctx = rep.ctx
  sub = {k: v[:, :C5] for k, v in batch.items()}
  util.load_batch(ctx, spec, sub, None, None)
  dev = ctx.g_obs.device
  ctx.g_obs.copy_(torch.tensor(g_ctx).reshape(C5, B, -1).to(dev))
  rep.ol.g_img.copy_(torch.tensor(g_ol).reshape(T2, B, -1).to(dev))
  rep.im.g_img.copy_(torch.tensor(g_im).reshape(H, B, -1).to(dev))
  rep.im.eps_act.copy_(torch.tensor(g_act).to(dev))
  rep.actions.copy_(torch.tensor(batch['action'][:, C5:]).transpose(0, 1).to(dev))
  rep.launch(draw=False)
  be.sync()
  truth = {k: np.asarray(batch[k], np.float32) / np.float32(255.0) for k in ospec.cnn_keys}
  vids = rep.videos(truth)

  grid = lambda v: v.transpose(1, 2, 0, 3, 4).reshape(v.shape[1], v.shape[2], -1, v.shape[4])
  c0 = 0
  for k, shp in ospec.cnn_keys.items():
    m = model[..., c0:c0 + shp[2]].numpy()
    want = grid(np.concatenate([truth[k], m, (m - truth[k] + 1) / 2], 2))
    got = vids[f'openl_{k}']
    assert got.shape == want.shape == (L, 3 * 18, B * 18, shp[2])
    assert np.abs(got - want).max() < tol, (k, np.abs(got - want).max())
    want = grid(imag[..., c0:c0 + shp[2]].numpy())
    got = vids[f'task_imag_{k}']
    assert got.shape == want.shape == (H + 1, 18, B * 18, shp[2])
    assert np.abs(got - want).max() < tol, (k, np.abs(got - want).max())
    c0 += shp[2]


def test_report_videos_match_oracle():
  run_report(HostSim(), 2e-5)


@pytest.mark.gpu
def test_report_videos_match_oracle_gpu():
  """The same comparison with every primitive on the CUDA kernels (pixels in [0, 1])."""
  from daydreamer_b200.backend import CudaBackend
  run_report(CudaBackend(), 1e-4)
