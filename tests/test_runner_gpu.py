"""The sync-free side runner on the GPU: a small CapsNet built from the drop-in layers, trained for an
epoch by `runner.train` (staged copies, no per-step synchronisation) and by the reference's loop order with its three
host stalls per step (main.py:55-77): same avg_loss bits, same parameters, same predictions."""
import copy
import types

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


class SmallCapsNet(nn.Module):
    """conv -> primary capsules -> routing -> capsule lengths, the reference CapsuleNet's chain (models.py:86-117) at a
    small size; model(x, y, True) also returns a reconstruction like models.py:118-124."""

    def __init__(self, pkg):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 32, 5)                                              # 16x16 -> 12x12
        self.primary = pkg.CapsuleLayer(None, n_caps=8, n_nodes=-1, in_C=32, out_C=4, kernel=4, stride=2)   # 5x5 -> 100 caps
        self.digits = pkg.CapsuleLayer(None, n_caps=10, n_nodes=4 * 5 * 5, in_C=8, out_C=16)
        self.decoder = nn.Linear(16, 3 * 16 * 16)

    def forward(self, x, y=None, is_recon=False):
        v = self.digits(self.primary(F.relu(self.conv1(x)))).squeeze()
        scores = (v ** 2).sum(dim=-1) ** 0.5
        if is_recon:
            picked = v[torch.arange(v.size(0), device=v.device), y]
            return scores, self.decoder(picked)
        return scores


def loss_fn(scores, y, params, x=None, recon=None):
    """loss_fns.py:11-23"""
    t = F.one_hot(y, scores.size(1)).float()
    loss = (t * F.relu(0.9 - scores) ** 2 + 0.5 * (1 - t) * F.relu(scores - 0.1) ** 2).sum()
    if recon is not None:
        loss = loss + params.recon_coef * F.mse_loss(recon, x.reshape(x.size(0), -1), reduction='sum')
    return loss / scores.size(0)


def stalled_epoch(x, y, model, optimizer, params):
    """main.py:38-77 (shuffle off)"""
    model.train()
    n_batch = (len(y) + params.batch_size - 1) // params.batch_size
    avg_loss, y_hat = 0, []
    for x_bch, y_bch in zip(np.array_split(x, n_batch), np.array_split(y, n_batch)):
        x_bch = torch.from_numpy(x_bch).float().permute(0, 3, 1, 2).contiguous().to(device=params.device)
        y_bch = torch.from_numpy(y_bch).to(device=params.device)
        y_hat_bch, recon = model(x_bch, y_bch, True)
        loss = loss_fn(y_hat_bch, y_bch, params, x_bch, recon)
        y_hat.append(y_hat_bch.data.cpu().numpy())
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        avg_loss += loss.item() / n_batch
    return avg_loss, np.concatenate(y_hat, axis=0)


def test_runner_epoch_matches_stalled_loop_on_gpu():
    import cs231_capsule_yolo_traffic_sign_detection_b200 as pkg
    pkg._cabi.lib()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    rng = np.random.RandomState(0)
    x = rng.uniform(-1, 1, (70, 16, 16, 3)).astype(np.float32)
    y = rng.randint(0, 10, 70).astype(np.int64)
    params = types.SimpleNamespace(device='cuda', batch_size=16, model='capsule', recon=True, recon_coef=5e-4)
    torch.manual_seed(0)
    m_ref = SmallCapsNet(pkg).cuda()
    m_run = copy.deepcopy(m_ref)
    o_ref, o_run = torch.optim.Adam(m_ref.parameters(), lr=1e-3), torch.optim.Adam(m_run.parameters(), lr=1e-3)
    seen = []
    metric = lambda yy, yh, p: seen.append(yh.copy()) or float((yh.argmax(1) == yy).mean())
    for _ in range(2):
        want_loss, want_pred = stalled_epoch(x, y, m_ref, o_ref, params)
        got_loss, got_metric = pkg.runner.train(x, y, m_run, o_run, loss_fn, metric, params, shuffle=False)
        assert got_loss == want_loss
        assert np.array_equal(seen[-1], want_pred)
        assert got_metric == float((want_pred.argmax(1) == y).mean())
    for a, b in zip(m_ref.parameters(), m_run.parameters()):
        assert torch.equal(a, b)
    # evaluate: no_grad path of the drop-in layer, ragged last batches
    ev = pkg.runner.evaluate(x[:37], y[:37], m_run, lambda s, yy, p: loss_fn(s, yy, p), metric,
                             types.SimpleNamespace(device='cuda', batch_size=16, model='cnn', recon=False))
    assert np.isfinite(ev[0]) and seen[-1].shape == (37, 10)


def test_graphed_runner_matches_eager_runner():
    """graph=True: every step replayed from one CUDA graph per batch size (first two batches of a size eager), against
    the same runner without graphs; both with a capturable Adam.  np.array_split(99, 7) gives one batch of 15 and six
    of 14: the graph of size 14 is captured at its third batch and replayed three times per epoch, size 15 stays eager."""
    import cs231_capsule_yolo_traffic_sign_detection_b200 as pkg
    pkg._cabi.lib()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    rng = np.random.RandomState(1)
    x = rng.uniform(-1, 1, (99, 16, 16, 3)).astype(np.float32)
    y = rng.randint(0, 10, 99).astype(np.int64)
    params = types.SimpleNamespace(device='cuda', batch_size=16, model='capsule', recon=True, recon_coef=5e-4)
    torch.manual_seed(0)
    m_e = SmallCapsNet(pkg).cuda()
    m_g = copy.deepcopy(m_e)
    o_e = torch.optim.Adam(m_e.parameters(), lr=1e-3, capturable=True)
    o_g = torch.optim.Adam(m_g.parameters(), lr=1e-3, capturable=True)
    preds = []
    metric = lambda yy, yh, p: preds.append(yh.copy()) or 0.0
    for epoch in range(2):
        le = pkg.runner.train(x, y, m_e, o_e, loss_fn, metric, params, shuffle=False)[0]
        lg = pkg.runner.train(x, y, m_g, o_g, loss_fn, metric, params, shuffle=False, graph=True)[0]
        assert abs(lg - le) <= 1e-6 * abs(le), (epoch, lg, le)
        assert np.allclose(preds[-1], preds[-2], rtol=1e-5, atol=1e-7)
    assert any(isinstance(v, pkg.runner._GraphedStep) for v in m_g._caps_runner_graphs['steps'].values())
    for a, b in zip(m_e.parameters(), m_g.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)
    with pytest.raises(ValueError):
        pkg.runner.train(x, y, m_g, torch.optim.Adam(m_g.parameters()), loss_fn, metric, params, graph=True)
    # a different optimizer object starts over: its graphs are captured against its own state tensors
    o2 = torch.optim.Adam(m_g.parameters(), lr=1e-3, capturable=True)
    pkg.runner.train(x[:30], y[:30], m_g, o2, loss_fn, metric, params, shuffle=False, graph=True)
    assert m_g._caps_runner_graphs['optimizer']() is o2
    assert not any(isinstance(v, pkg.runner._GraphedStep) for v in m_g._caps_runner_graphs['steps'].values())   # 2 batches of 15: still eager
    pkg.runner.reset_graphs(m_g)
    assert '_caps_runner_graphs' not in m_g.__dict__
