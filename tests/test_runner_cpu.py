"""The sync-free side runner against a transcription of the reference's epoch loops
(main.py:38-95 train, :98-140 evaluate) on the CPU: same batches, same step order, same avg_loss / metric /
parameter bits."""
import copy
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from cs231_capsule_yolo_traffic_sign_detection_b200 import runner


class TinyNet(nn.Module):
    """Stands in for CapsuleNet's call shapes: model(x) -> scores, model(x, y, True) -> (scores, recon)."""

    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 4, 3, padding=1)
        self.fc = nn.Linear(4 * 8 * 8, 5)
        self.dec = nn.Linear(5, 3 * 8 * 8)

    def forward(self, x, y=None, is_recon=False):
        s = self.fc(F.relu(self.conv(x)).reshape(x.size(0), -1))
        if is_recon:
            return s, self.dec(s * F.one_hot(y, 5).float())
        return s


def loss_fn(y_hat, y, params, x=None, recon=None):
    loss = F.cross_entropy(y_hat, y)
    if recon is not None:
        loss = loss + params.recon_coef * F.mse_loss(recon, x.reshape(x.size(0), -1), reduction='sum') / x.size(0)
    return loss


def metric(y, y_hat, params):
    return float((y_hat.argmax(1) == y).mean())


def reference_train(x, y, model, optimizer, loss_fn, metric, params):
    """main.py:38-95 with shuffle off, tqdm off (test infrastructure: the semantics the runner must keep)."""
    model.train()
    total = len(y)
    n_batch = (total + params.batch_size - 1) // params.batch_size
    x_split, y_split = np.array_split(x, n_batch), np.array_split(y, n_batch)
    avg_loss, y_hat = 0, []
    for x_bch, y_bch in zip(x_split, y_split):
        x_bch = torch.from_numpy(x_bch).float().permute(0, 3, 1, 2).to(device=params.device)
        y_bch = torch.from_numpy(y_bch).to(device=params.device)
        if params.model == 'capsule' and params.recon:
            y_hat_bch, recon = model(x_bch, y_bch, True)
            loss = loss_fn(y_hat_bch, y_bch, params, x_bch, recon)
        else:
            y_hat_bch = model(x_bch)
            loss = loss_fn(y_hat_bch, y_bch, params)
        y_hat.append(y_hat_bch.data.cpu().numpy())
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        avg_loss += loss.item() / n_batch
    y_hat = np.concatenate(y_hat, axis=0)
    return avg_loss, metric(y, y_hat, params)


def reference_evaluate(x, y, model, loss_fn, metric, params):
    model.eval()
    total = len(y)
    n_batch = (total + params.batch_size - 1) // params.batch_size
    x_split, y_split = np.array_split(x, n_batch), np.array_split(y, n_batch)
    avg_loss, y_hat = 0, []
    with torch.no_grad():
        for x_bch, y_bch in zip(x_split, y_split):
            x_bch = torch.from_numpy(x_bch).float().permute(0, 3, 1, 2).to(device=params.device)
            y_bch = torch.from_numpy(y_bch).to(device=params.device)
            y_hat_bch = model(x_bch)
            loss = loss_fn(y_hat_bch, y_bch, params)
            y_hat.append(y_hat_bch.data.cpu().numpy())
            avg_loss += loss.item() / n_batch
    return avg_loss, metric(y, np.concatenate(y_hat, axis=0), params)


def _data(n=23):
    rng = np.random.RandomState(0)
    return rng.uniform(-1, 1, (n, 8, 8, 3)).astype(np.float32), rng.randint(0, 5, n).astype(np.int64)


def _params(model, recon):
    return types.SimpleNamespace(device='cpu', batch_size=5, model=model, recon=recon, recon_coef=5e-4)


def test_train_epoch_matches_reference_loop():
    x, y = _data()
    for model_name, recon in (('capsule', True), ('cnn', False)):
        params = _params(model_name, recon)
        torch.manual_seed(0)
        m_ref = TinyNet()
        m_run = copy.deepcopy(m_ref)
        o_ref, o_run = torch.optim.Adam(m_ref.parameters(), lr=1e-2), torch.optim.Adam(m_run.parameters(), lr=1e-2)
        for _ in range(2):          # two epochs: optimizer state carries over
            want = reference_train(x, y, m_ref, o_ref, loss_fn, metric, params)
            got = runner.train(x, y, m_run, o_run, loss_fn, metric, params, shuffle=False)
            assert got == want              # avg_loss accumulated in the same order and precision: same bits
        for a, b in zip(m_ref.parameters(), m_run.parameters()):
            assert torch.equal(a, b)


def test_evaluate_matches_reference_loop_and_ragged_batches():
    x, y = _data(17)                        # 17 = 5 + 4 + 4 + 4 (np.array_split), like main.py:43-44
    params = _params('cnn', False)
    torch.manual_seed(1)
    m = TinyNet()
    assert runner.evaluate(x, y, m, loss_fn, metric, params) == reference_evaluate(x, y, m, loss_fn, metric, params)
    assert runner.evaluate(x, y, m, loss_fn, metric, params, no_metric=True)[1] == -1       # main.py:85


def test_shuffle_follows_numpy_global_rng():
    x, y = _data()
    params = _params('cnn', False)
    torch.manual_seed(0)
    m1 = TinyNet()
    m2 = copy.deepcopy(m1)
    np.random.seed(3)
    i = np.random.permutation(len(y))       # utils.shuffle (utils.py:146-148)
    want = reference_train(x[i], y[i], m1, torch.optim.SGD(m1.parameters(), lr=0.1), loss_fn, metric, params)
    np.random.seed(3)
    got = runner.train(x, y, m2, torch.optim.SGD(m2.parameters(), lr=0.1), loss_fn, metric, params)
    assert got == want


def test_other_input_dtypes_and_float_targets():
    """uint8 / float64 images go through main.py:57's `.float()`; float targets (the darkcapsule label grid) pass through
    unchanged; params.device may be a torch.device."""
    rng = np.random.RandomState(2)
    params = types.SimpleNamespace(device=torch.device('cpu'), batch_size=4, model='darkcapsule', recon=False, recon_coef=0.0)
    y = rng.uniform(0, 1, (10, 5)).astype(np.float32)
    mse = lambda y_hat, yy, p: ((y_hat - yy) ** 2).mean()
    for x in (rng.randint(0, 255, (10, 8, 8, 3)).astype(np.uint8), rng.uniform(-1, 1, (10, 8, 8, 3))):
        torch.manual_seed(0)
        m = TinyNet()
        got = runner.evaluate(x, y, m, mse, lambda a, b, p: float(np.abs(a - b).mean()), params)
        want = reference_evaluate(x, y, m, mse, lambda a, b, p: float(np.abs(a - b).mean()), params)
        assert got == want


def test_graph_flag_is_ignored_on_cpu():
    x, y = _data(9)
    params = _params('cnn', False)
    torch.manual_seed(0)
    m1 = TinyNet()
    m2 = copy.deepcopy(m1)
    a = runner.train(x, y, m1, torch.optim.SGD(m1.parameters(), lr=0.1), loss_fn, metric, params, shuffle=False)
    b = runner.train(x, y, m2, torch.optim.SGD(m2.parameters(), lr=0.1), loss_fn, metric, params, shuffle=False, graph=True)
    assert a == b and '_caps_runner_graphs' not in m2.__dict__
