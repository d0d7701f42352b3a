"""Host logic of the data-parallel trainer drop-in (vault_b200/trainer.py) with a stub step and a toy classifier: CPU, gloo."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vault_b200.trainer import BestTracker, ShardSampler, VaultTrainerForBloombergTwitterCorpus, VaultTrainerForMVSA, VaultTrainerForTMSC


class ToyData(torch.utils.data.Dataset):
    """Items shaped like the reference's Twitter201X examples: (id, input_ids, text_mask, type_ids, image, image_mask, label)."""
    name = "toy"

    def __init__(self, n, wrong_every=5):
        self.n, self.wrong_every = n, wrong_every

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        ids = torch.full((4,), i, dtype=torch.long)
        label = i % 3 if i % self.wrong_every else (i + 1) % 3  # the toy model predicts i % 3: every 5th item is "wrong"
        return (i, ids, torch.ones(4, dtype=torch.long), torch.zeros(4, dtype=torch.long), torch.zeros(3, 2, 2), torch.ones(2, 2), label)

    @staticmethod
    def collate_fn(items):
        cols = list(zip(*items))
        return [list(cols[0])] + [torch.stack(c) for c in cols[1:6]] + [torch.tensor(cols[6])]


class ToyModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))

    def forward(self, input_ids=None, **kw):
        return torch.nn.functional.one_hot(input_ids[:, 0] % 3, 3).float() * 4.0 + self.w


class StubStep:
    seen = None

    def __init__(self, model, **hp):
        self.hp, self.calls, self.ids = hp, 0, []

    def step(self, batch):
        self.calls += 1
        self.ids.extend(batch["input_ids"][:, 0].tolist())
        loss = float(batch["input_ids"][:, 0].float().mean())  # any batch-dependent number
        return SimpleNamespace(loss=lambda: loss)

    def synchronize(self):
        pass


def handler(**over):
    log = SimpleNamespace(metrics=[], test=[], best=[], logged=0)
    h = SimpleNamespace(device="cpu", learning_rate=2e-5, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8, weight_decay=0.0, correct_bias=False,
                        train_batch_size=4, eval_batch_size=4, dataloader_num_workers=0, num_train_epochs=2, warmup_ratio=0.1, max_steps=-1, eval_steps=None,
                        disable_tqdm=True, early_stopping_patience=None, model_save=False, model_load_filename=None, _log=log)
    h.set_dict_metrics = lambda r, test=False: (log.test if test else log.metrics).append(dict(r))
    h.set_best = lambda *a, **k: log.best.append((a, k))
    h.log = lambda: setattr(log, "logged", log.logged + 1)
    for k, v in over.items():
        setattr(h, k, v)
    return h


def make(ds_n=22, **over):
    steps = []

    def factory(model, **hp):
        s = StubStep(model, **hp)
        steps.append(s)
        return s

    h = handler(**over)
    t = VaultTrainerForTMSC(ToyModel(), ToyData(ds_n), h, dev_dataset=ToyData(10), test_dataset=ToyData(15), step_factory=factory)
    return t, h, steps


def test_single_process_loop_matches_reference_bookkeeping():
    t, h, steps = make()
    res = t.train()
    s = steps[0]
    n_batches = 6  # ceil(22 / 4)
    assert s.calls == 2 * n_batches and sorted(s.ids) == sorted(list(range(22)) * 2)
    # schedule: len(loader) * epochs, reference hyper-parameters, CE loss
    assert s.hp["total_steps"] == 12 and s.hp["warmup_ratio"] == 0.1 and s.hp["lr"] == 2e-5 and s.hp["correct_bias"] is False and s.hp["loss"] == "ce"
    # one evaluation window per epoch (eval_steps None -> epoch): train_loss is the sample-weighted mean of the step losses
    assert len(h._log.metrics) == 2
    assert abs(h._log.metrics[0]["eval_accuracy"] - 0.8) < 1e-9          # items 0 and 5 of the 10 dev items are wrong
    assert abs(res["eval_accuracy"] - 12 / 15) < 1e-9 and len(h._log.test) == 1 and h._log.logged == 1
    assert 0.0 < h._log.metrics[0]["eval_loss"] < 2.0


def test_max_steps_and_early_stopping():
    t, h, steps = make(max_steps=3)
    t.train()
    assert steps[0].calls == 3
    # patience: the metric never improves after the first evaluation -> stop after `patience` further evaluations
    t, h, steps = make(early_stopping_patience=2, eval_steps=2, num_train_epochs=5)
    t.train()
    assert steps[0].calls == 6 and len(h._log.metrics) == 3 and len(h._log.best) == 1
    bt = BestTracker(ToyModel(), patience=1, higher_better=False)
    assert bt.step(1.0, eval_loss=1.0) is False and bt.step(0.5, eval_loss=0.5) is False and bt.step(0.7, eval_loss=0.7) is True
    assert bt.get_metrics() == {"best_eval_loss": 0.5}


def test_task_variants_losses_and_predictions():
    logits1, y1 = torch.tensor([2.0, -1.0, 0.3]), torch.tensor([1.0, 0.0, 0.0])
    b = VaultTrainerForBloombergTwitterCorpus(ToyModel(), ToyData(4), handler())
    assert b.loss_kind == "bce" and b.early_stopping_metric == "eval_loss" and not b.higher_better
    assert torch.allclose(b.calculate_loss(logits1, y1, False), torch.nn.functional.binary_cross_entropy_with_logits(logits1, y1))
    assert b.get_eval_preds_from_batch(logits1) == [1, 0, 1] and b.input_batch_kwargs(({"input_ids": 1}, 2)) == {"input_ids": 1}
    raw = ToyData(4)
    raw.preprocessed = False
    m = VaultTrainerForMVSA(ToyModel(), raw, handler())
    assert m.loss_kind == "ce2"
    lg = torch.tensor([[3.0, 0, 0, 0, 0, 5.0], [0, 2.0, 0, 1.0, 0, 0]])
    lab = torch.tensor([[0, 2], [1, 1]])
    ce = torch.nn.functional.cross_entropy
    assert torch.allclose(m.calculate_loss(lg, lab, False), 0.5 * (ce(lg[:, :3], lab[:, 0]) + ce(lg[:, 3:], lab[:, 1])))
    assert m.get_eval_preds_from_batch(lg) == [[0, 2], [1, 0]]
    assert m.evaluation_metrics([[0, 2], [1, 1]], [[0, 2], [1, 0]])["eval_accuracy"] == 0.75
    pre = ToyData(4)
    pre.preprocessed = True
    assert VaultTrainerForMVSA(ToyModel(), pre, handler()).loss_kind == "ce"


def test_shard_sampler_covers_every_item_once():
    parts = [list(ShardSampler(11, 3, r)) for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(11)) and [len(p) for p in parts] == [4, 4, 3]


def _worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t, h, steps = make()
        res = t.train()
        out.put((rank, steps[0].calls, sorted(steps[0].ids), steps[0].hp["total_steps"], h._log.metrics, res, len(h._log.test), h._log.logged))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_ranks_shard_training_and_agree_on_metrics():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(30)
    (r0, calls0, ids0, total0, metrics0, res0, ntest0, logged0), (r1, calls1, ids1, total1, metrics1, res1, ntest1, logged1) = got
    # 22 items over 2 ranks: 11 each per epoch -> 3 batches per rank per epoch; schedule counts PER-RANK batches (reference formula)
    assert calls0 == calls1 == 6 and total0 == total1 == 6
    assert len(ids0) == len(ids1) == 22 and sorted(set(ids0 + ids1)) == list(range(22))
    # evaluation: both ranks compute the same, complete metrics; only rank 0 talks to the experiment handler
    assert abs(res0["eval_accuracy"] - 12 / 15) < 1e-9 and res0 == res1
    assert len(metrics0) == 2 and metrics1 == [] and ntest0 == 1 and ntest1 == 0 and logged0 == 1 and logged1 == 0
    assert abs(metrics0[0]["eval_accuracy"] - 0.8) < 1e-9
