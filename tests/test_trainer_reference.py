"""The trainer drop-in (vault_b200/trainer.py) pinned against the REFERENCE trainer itself: ``Twitter201XTrainer`` of
ref:vault/tmsc_utils/trainer.py (train loop :282-420, loss bookkeeping :369-375, evaluate :422-484, metrics :513-549, optimizer / schedule
:244-280) and ``EarlyStopping`` of ref:vault/train_utils.py:13-160, loaded by path through oracle/ref_loader.py (from /root/reference, or
from the byte-identical copies oracle/build_ref.py leaves under oracle/_ref/).  CPU only: a toy classifier and a toy dataset go through
both loops from the same seed; the fused GPU step is replaced by a host step that applies the same rule the reference's optimizer +
scheduler apply, with the learning rate taken from ``VaultTrainStep.lr_at`` (the schedule code the GPU step really uses).

Compared: every logged metric dict (sample-weighted train_loss per evaluation window, eval_loss, eval_accuracy, macro_f1_score), the
early-stopping decision and its best metrics, the test metrics, the per-step learning rates and the final weights."""
import copy
import os
import sys
from types import SimpleNamespace

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.trainer_available(), reason="reference trainer files not present (python -m oracle.build_ref)")


class ToyData(torch.utils.data.Dataset):
    """(id, x, label): separable-ish 3-class points, a fixed fraction mislabelled so that accuracy / F1 are not trivially 1."""
    name = "toy"

    def __init__(self, n, seed):
        g = torch.Generator().manual_seed(seed)
        self.y = torch.randint(0, 3, (n,), generator=g)
        self.x = torch.nn.functional.one_hot(self.y, 3).float() * 1.5 + torch.randn(n, 3, generator=g)
        flip = torch.rand(n, generator=g) < 0.15
        self.y = torch.where(flip, (self.y + 1) % 3, self.y)

    def __len__(self):
        return len(self.y)

    def __getitem__(self, i):
        return (i, self.x[i], self.y[i])

    @staticmethod
    def collate_fn(items):
        ids, xs, ys = zip(*items)
        return [torch.tensor(ids), torch.stack(xs), torch.stack(ys)]


class ToyModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(5)
        self.lin = torch.nn.Linear(3, 3)
        with torch.no_grad():
            self.lin.weight.copy_(torch.randn(3, 3, generator=g) * 0.3)
            self.lin.bias.zero_()

    def forward(self, x=None):
        return self.lin(x)


def handler(**over):
    log = SimpleNamespace(metrics=[], test=[], best=[], logged=0)
    h = SimpleNamespace(device="cpu", learning_rate=5e-2, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8, weight_decay=0.0, correct_bias=False,
                        train_batch_size=8, eval_batch_size=16, dataloader_num_workers=0, num_train_epochs=4, warmup_ratio=0.1, max_steps=-1, eval_steps=5,
                        disable_tqdm=True, early_stopping_patience=None, model_save=False, model_load_filename=None, model_save_filename=None, _log=log)
    h.set_dict_metrics = lambda r, test=False: (log.test if test else log.metrics).append({k: float(v) for k, v in r.items()})
    h.set_best = lambda *a, **k: log.best.append((a, k))
    h.log = lambda: setattr(log, "logged", log.logged + 1)
    h.aggregate_results = lambda: None
    h.plot = lambda: None
    for k, v in over.items():
        setattr(h, k, v)
    return h


class HostStep:
    """Stand-in for VaultTrainStep on the CPU: forward, CE (mean), zero_grad, backward, the HF-AdamW rule with the learning rate of
    VaultTrainStep.lr_at(step) -- the same sequence the fused GPU step performs (vault_b200/train.py), one torch op at a time."""

    def __init__(self, model, lr, betas, eps, weight_decay, correct_bias, total_steps, warmup_ratio, loss):
        from vault_b200.train import VaultTrainStep

        assert loss == "ce"
        self.model = model
        self.opt = ref_loader.HFAdamW(model.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        self._sched = SimpleNamespace(lr=lr, total_steps=total_steps, warmup_ratio=warmup_ratio)
        self._lr_at = lambda step: VaultTrainStep.lr_at(self._sched, step)
        self.step_idx, self.lrs = 0, []

    def step(self, batch):
        lr = self._lr_at(self.step_idx)
        for g in self.opt.param_groups:
            g["lr"] = lr
        self.lrs.append(lr)
        logits = self.model(x=batch["x"])
        loss = torch.nn.functional.cross_entropy(logits, batch["labels"])
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        self.step_idx += 1
        val = loss.item()
        return SimpleNamespace(loss=lambda: val)

    def synchronize(self):
        pass


def run_reference(h, train, dev, test):
    tr, _ = ref_loader.load_reference_trainer()

    class RefToy(tr.Twitter201XTrainer):
        def input_batch_kwargs(self, batch):
            return dict(x=batch[1])

    lrs = []
    t = RefToy(ToyModel(), train, h, dev_dataset=dev, test_dataset=test)
    orig = t.init_optimizer_scheduler

    def wrapped(n):
        opt, sch = orig(n)
        step0 = opt.step

        def step(*a, **k):  # the learning rate the optimizer really applies at each step
            lrs.append(opt.param_groups[0]["lr"])
            return step0(*a, **k)

        opt.step = step
        return opt, sch

    t.init_optimizer_scheduler = wrapped
    torch.manual_seed(123)
    t.train()
    return t, lrs


def run_mine(h, train, dev, test):
    from vault_b200.trainer import Twitter201XTrainer

    class MyToy(Twitter201XTrainer):
        def input_batch_kwargs(self, batch):
            return dict(x=batch[1])

    steps = []

    def factory(model, **hp):
        steps.append(HostStep(model, **hp))
        return steps[-1]

    t = MyToy(ToyModel(), train, h, dev_dataset=dev, test_dataset=test, step_factory=factory)
    torch.manual_seed(123)
    res = t.train()
    return t, steps[0].lrs, res


@pytest.mark.parametrize("over", [dict(), dict(eval_steps=3, early_stopping_patience=2, num_train_epochs=8), dict(max_steps=7, eval_steps=2),
                                  dict(correct_bias=True, weight_decay=0.01, warmup_ratio=0.3)])
def test_loop_matches_the_reference_trainer(over):
    train, dev, test = ToyData(52, 1), ToyData(33, 2), ToyData(41, 3)  # 52 / 8 -> a ragged last batch: sample-weighted train_loss matters
    h_ref, h_my = handler(**over), handler(**over)
    t_ref, lr_ref = run_reference(h_ref, train, dev, test)
    t_my, lr_my, res = run_mine(h_my, train, dev, test)
    # learning rate applied at every step: linear warm-up from 0 over int(ratio * steps), then linear decay (ref :256-280)
    assert len(lr_ref) == len(lr_my) and len(lr_my) > 0
    assert max(abs(a - b) for a, b in zip(lr_ref, lr_my)) < 1e-12
    # every evaluation window: same metric names and values (loss bookkeeping ref :369-375, metrics ref :513-549)
    assert len(h_ref._log.metrics) == len(h_my._log.metrics) and len(h_my._log.metrics) > 0
    for a, b in zip(h_ref._log.metrics, h_my._log.metrics):
        assert set(a) == set(b) == {"train_loss", "eval_loss", "eval_accuracy", "macro_f1_score"}
        for k in a:
            assert abs(a[k] - b[k]) < 1e-6, (k, a, b)
    assert len(h_ref._log.test) == len(h_my._log.test) == 1
    for k in h_ref._log.test[0]:
        assert abs(h_ref._log.test[0][k] - h_my._log.test[0][k]) < 1e-6 and abs(res[k] - h_my._log.test[0][k]) < 1e-12
    # early stopping: same best metrics (or none without patience), same number of set_best calls, training ended at the same step
    b_ref, b_my = t_ref.early_stopping.get_metrics(), t_my.early_stopping.get_metrics()
    assert (b_ref is None) == (b_my is None) and len(h_ref._log.best) == len(h_my._log.best)
    if b_ref is not None:
        assert set(b_ref) == set(b_my)
        for k in b_ref:
            assert abs(float(b_ref[k]) - float(b_my[k])) < 1e-6
    for (n1, p1), (n2, p2) in zip(t_ref.model.named_parameters(), t_my.model.named_parameters()):
        assert n1 == n2 and torch.allclose(p1, p2, rtol=0, atol=1e-6)
    assert h_ref._log.logged == h_my._log.logged == 1


def test_early_stopping_checkpoint_round_trip_matches_the_reference():
    """ref:vault/train_utils.py:127-140 -- with model_save the reference writes model.state_dict() to a temp file at every new best and
    train_end() loads it back; the drop-in keeps a host copy.  Both must hand back the weights of the BEST evaluation, not the last."""
    over = dict(eval_steps=3, early_stopping_patience=2, num_train_epochs=8, model_save=True, learning_rate=0.3)
    train, dev, test = ToyData(52, 1), ToyData(33, 2), ToyData(41, 3)
    h_ref, h_my = handler(**over), handler(**over)
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        h_ref.model_save_filename = os.path.join(tmp, "ref.pt")
        h_my.model_save_filename = os.path.join(tmp, "my.pt")
        t_ref, _ = run_reference(h_ref, train, dev, test)
        t_my, _, _ = run_mine(h_my, train, dev, test)
        sd_ref, sd_my = torch.load(h_ref.model_save_filename), torch.load(h_my.model_save_filename)
    best = t_ref.early_stopping.get_metrics()
    assert best is not None and int(best["best_step"]) < 3 * len(h_ref._log.metrics)  # training went on past the best evaluation
    assert set(sd_ref) == set(sd_my)
    for k in sd_ref:
        assert torch.allclose(sd_ref[k], sd_my[k], rtol=0, atol=1e-6)
        assert torch.allclose(dict(t_my.model.state_dict())[k].cpu(), sd_my[k])  # the live model IS the restored best one
    # ... and it differs from where training stopped: patience ran out AFTER the best evaluation, so later steps moved the weights
    assert len(h_my._log.metrics) * 3 > int(best["best_step"])


def test_state_dict_round_trip_through_the_reference_early_stopping_on_vault_models():
    """The state-dict contract the reference's EarlyStopping relies on (torch.save(model.state_dict()) -> load_state_dict), exercised on the
    drop-in VaultForTMSC itself (CPU: construction / state_dict only): same keys as the reference class, values survive the round trip."""
    from oracle import synth
    from oracle.ref_loader import hf_configs
    from vault_b200 import VaultForTMSC

    _, tu = ref_loader.load_reference_trainer()
    d = synth.Dims.tiny()
    vc, lc = hf_configs(d)
    m = VaultForTMSC(vc, n_classes=3, vilt_dropout_prob=0.1, bert_config=lc)
    ref_m, _ = ref_loader.build_reference_tmsc(d, synth.make_state_dict(d, seed=0))
    keys = {k for k in m.state_dict() if not k.endswith("position_ids") and not k.endswith("token_type_ids")}
    ref_keys = {k for k in ref_m.state_dict() if not k.endswith("position_ids") and not k.endswith("token_type_ids")}
    assert keys == ref_keys
    es = tu.EarlyStopping(m, patience=1, save_model=True, higher_better=True)
    assert es.step(0.5, eval_accuracy=0.5) is False  # new best: saved to the temp file
    before = copy.deepcopy({k: v.clone() for k, v in m.state_dict().items()})
    with torch.no_grad():
        for p in m.parameters():
            p.add_(1.0)
    assert es.step(0.4, eval_accuracy=0.4) is True  # no improvement, patience 1 -> stop
    best = es.best_model()
    assert best is m
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert es.get_metrics()["best_eval_accuracy"] == 0.5
