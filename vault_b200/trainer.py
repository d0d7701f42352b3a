"""Data-parallel drop-in for the reference's training loop.

Mirrors ``Twitter201XTrainer.train / evaluate`` (ref:vault/tmsc_utils/trainer.py:282-484) and its three task subclasses
(ref:vault/models/vault/trainer.py): same constructor, same hook methods (``input_batch_kwargs``, ``batch_labels``, ``batch_len``,
``calculate_loss``, ``get_eval_preds_from_batch``, ``get_eval_true_from_batch``, ``evaluation_metrics``), the same
``experiment_handler`` duck type (the reference's ``ExperimentHandler`` is used unchanged: attribute reads + ``set_dict_metrics`` /
``set_best`` / ``log`` / ``aggregate_results`` / ``plot``), so ``experiments/clsf_vault.py`` keeps working with the import swapped and
``torchrun --nproc-per-node N`` in front.  What changes is the loop body:

  * one process per GPU; a ``DistributedSampler`` shards the training set (``set_epoch`` per epoch), evaluation sets are split
    ``indices[rank::world]`` (no padding duplicates) and predictions are all-gathered, so every rank holds the same metrics;
  * the step is ``VaultTrainStep`` (CUDA graphs, fused loss / AdamW, gradient all-reduce overlapped with backward); batches come
    out of a pinned-memory loader and are copied on a side stream while the previous step computes;
  * no per-step ``loss.item()`` (ref :369): each step returns a handle, the handles are resolved once per evaluation window;
  * logging, metric bookkeeping, checkpoint saving happen on rank 0 only.

The schedule is the reference's: ``num_steps = len(loader) * epochs`` with ``len(loader)`` the PER-RANK batch count, linear warm-up
over ``int(warmup_ratio * num_steps)`` steps, then linear decay; HF-AdamW rule without bias correction unless asked.
"""
from __future__ import annotations

import copy
import logging
from typing import Any, Callable, Dict, Iterable, List, Optional

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, Sampler


def _dist():
    import torch.distributed as dist

    return dist if (dist.is_available() and dist.is_initialized()) else None


def world_rank():
    d = _dist()
    return (d.get_world_size(), d.get_rank()) if d is not None else (1, 0)


class ShardSampler(Sampler):
    """Evaluation shard of rank r: indices r, r+world, ... -- every sample exactly once across ranks (DistributedSampler would
    pad the tail with repeats and skew the metrics)."""

    def __init__(self, n: int, world: int, rank: int):
        self.idx = list(range(rank, n, world))

    def __iter__(self):
        return iter(self.idx)

    def __len__(self):
        return len(self.idx)


class BestTracker:
    """The part of ref:vault/train_utils.py:13-160 (EarlyStopping) the loop needs: patience counting on one metric, the metrics of
    the best evaluation, and (optionally) the best weights -- kept as a host copy of the state dict instead of a temp file."""

    def __init__(self, model, patience: Optional[int], save_model: bool = False, delta: float = 0.0, higher_better: bool = False):
        self.model, self.patience, self.save_model, self.delta, self.higher_better = model, patience, save_model, delta, higher_better
        self.best, self.cnt, self.best_metrics, self.state = None, 0, None, None

    def new_best(self, metric: float) -> bool:
        if self.best is None:
            return True
        return metric > self.best + self.delta if self.higher_better else metric < self.best - self.delta

    def _save(self):
        if self.save_model:
            self.state = {k: v.detach().to("cpu", copy=True) for k, v in self.model.state_dict().items()}

    def step(self, metric: Optional[float], **metrics) -> bool:
        if self.patience is None or metric is None:
            self._save()
            return False
        if self.new_best(metric):
            self.best, self.cnt = metric, 0
            self.best_metrics = {"best_" + k: v for k, v in metrics.items()}
            self._save()
        else:
            self.cnt += 1
        return self.cnt >= self.patience

    def get_metrics(self) -> Optional[Dict[str, Any]]:
        return self.best_metrics

    def best_model(self):
        if self.state is not None:
            self.model.load_state_dict(self.state)
        return self.model


def _default_step_factory(model, **hp):
    from .train import VaultTrainStep

    return VaultTrainStep(model, **hp)


class Twitter201XTrainer:
    """See the module docstring.  ``step_factory(model, lr=..., betas=..., eps=..., weight_decay=..., correct_bias=..., total_steps=...,
    warmup_ratio=..., loss=...)`` builds the fused step (default: ``VaultTrainStep``); it must offer ``step(batch_dict) -> handle`` with
    ``handle.loss()`` and ``synchronize()``."""

    early_stopping_metric = "eval_accuracy"
    higher_better = True
    loss_kind = "ce"

    def __init__(self, model, dataset: Dataset, experiment_handler, dev_dataset: Optional[Dataset] = None, test_dataset: Optional[Dataset] = None,
                 logging_level=None, step_factory: Optional[Callable] = None):
        self.model, self.dataset, self.dev_dataset, self.test_dataset = model, dataset, dev_dataset, test_dataset
        self.do_eval, self.do_test = dev_dataset is not None, test_dataset is not None
        self.exp_handler = experiment_handler
        self.world, self.rank = world_rank()
        self.early_stopping = BestTracker(model, getattr(experiment_handler, "early_stopping_patience", None),
                                          bool(getattr(experiment_handler, "model_save", False)), higher_better=self.higher_better)
        self.step_factory = step_factory or _default_step_factory
        self.logger = logging.getLogger(__name__)
        self.logger.setLevel(logging_level or logging.WARNING)
        self.train_step = None

    # ---- hooks (same names and meaning as the reference) -----------------------------------------------------------------
    def input_batch_kwargs(self, batch: Iterable[Any]) -> Dict[str, Any]:
        """ref:vault/models/vault/trainer.py:19-37 -- (id, input_ids, text_mask, type_ids, image, image_mask, label)."""
        _, input_ids, text_mask, type_ids, image, image_mask, _ = batch
        return dict(input_ids=input_ids, attention_mask=text_mask, token_type_ids=type_ids, pixel_values=image, pixel_mask=image_mask)

    def batch_labels(self, batch):
        return batch[-1]

    def batch_len(self, batch) -> int:
        return len(self.batch_labels(batch))

    def get_logits_from_model(self, return_vals, *a, **k):
        return return_vals

    def calculate_loss(self, logits: torch.Tensor, labels: torch.Tensor, train: bool) -> torch.Tensor:
        """Evaluation-time loss (the training loss is fused into the step)."""
        return torch.nn.functional.cross_entropy(logits, labels)

    def get_eval_preds_from_batch(self, logits: torch.Tensor) -> List[List[int]]:
        preds = [ex.argmax(dim=-1).tolist() for ex in logits]
        return [[p] if isinstance(p, int) else p for p in preds]

    def get_eval_true_from_batch(self, labels: torch.Tensor) -> List[List[int]]:
        return [[lab.item()] if lab.ndim == 0 else lab.tolist() for lab in labels]

    def evaluation_metrics(self, eval_true, eval_preds, data_loader=None) -> Dict[str, float]:
        """ref:vault/tmsc_utils/trainer.py:502-545 -- accuracy over all targets + macro F1."""
        from sklearn.metrics import precision_recall_fscore_support

        flat_t = [x for row in eval_true for x in row]
        flat_p = [x for row in eval_preds for x in row]
        _, _, macro_f1, _ = precision_recall_fscore_support(flat_t, flat_p, average="macro", zero_division=0)
        return dict(eval_accuracy=float(np.mean([p == t for p, t in zip(flat_p, flat_t)])), macro_f1_score=float(macro_f1))

    # ---- plumbing ---------------------------------------------------------------------------------------------------------
    def _device(self) -> torch.device:
        return torch.device(getattr(self.exp_handler, "device", "cuda"))

    def _loader(self, ds: Dataset, batch_size: int, train: bool, epoch_seed: int = 0) -> DataLoader:
        sampler = None
        if train and self.world > 1:
            from torch.utils.data.distributed import DistributedSampler

            sampler = DistributedSampler(ds, num_replicas=self.world, rank=self.rank, shuffle=True, seed=epoch_seed)
        elif not train and self.world > 1:
            sampler = ShardSampler(len(ds), self.world, self.rank)
        return DataLoader(ds, batch_size=batch_size, shuffle=train and sampler is None, sampler=sampler, collate_fn=getattr(ds, "collate_fn", None),
                          num_workers=int(getattr(self.exp_handler, "dataloader_num_workers", 0)), pin_memory=self._device().type == "cuda")

    def _is_main(self) -> bool:
        return self.rank == 0

    def init_schedule(self, num_batches: int) -> Dict[str, Any]:
        """ref:vault/tmsc_utils/trainer.py:244-280 restated as the fused step's hyper-parameters."""
        eh = self.exp_handler
        return dict(lr=float(eh.learning_rate), betas=(float(eh.adam_beta1), float(eh.adam_beta2)), eps=float(eh.adam_epsilon),
                    weight_decay=float(eh.weight_decay), correct_bias=bool(getattr(eh, "correct_bias", False)),
                    total_steps=int(num_batches * int(eh.num_train_epochs)), warmup_ratio=float(eh.warmup_ratio), loss=self.loss_kind)

    def _flush_losses(self, pending: List) -> (float, int):
        tot, n = 0.0, 0
        for handle, bl in pending:
            tot += handle.loss() * bl
            n += bl
        pending.clear()
        return tot, n

    def _allreduce_sums(self, *vals: float) -> List[float]:
        d = _dist()
        if d is None or self.world == 1:
            return list(vals)
        t = torch.tensor(vals, dtype=torch.float64, device=self._device() if d.get_backend() == "nccl" else "cpu")
        d.all_reduce(t)
        return t.tolist()

    # ---- train / evaluate -------------------------------------------------------------------------------------------------
    def train(self):
        eh = self.exp_handler
        dev = self._device()
        self.model = self.model.to(dev)
        self.model.train()
        if getattr(eh, "model_load_filename", None) is not None:
            self.model.load_state_dict(torch.load(eh.model_load_filename))
        loader = self._loader(self.dataset, int(eh.train_batch_size), train=True)
        dev_loader = self._loader(self.dev_dataset, int(eh.eval_batch_size), train=False) if self.do_eval else None
        test_loader = self._loader(self.test_dataset, int(eh.eval_batch_size), train=False) if self.do_test else None
        self.train_step = self.step_factory(self.model, **self.init_schedule(len(loader)))
        num_epochs = int(eh.num_train_epochs)
        eval_steps = int(getattr(eh, "eval_steps", None) or len(loader))
        max_steps = int(getattr(eh, "max_steps", -1))
        early_stop = False
        pending: List = []
        train_loss, n_samples = 0.0, 0
        for epoch in range(num_epochs):
            if early_stop:
                break
            if hasattr(loader.sampler, "set_epoch"):
                loader.sampler.set_epoch(epoch)
            for step, batch in enumerate(loader):
                step += epoch * len(loader)
                early_stop = max_steps > 0 and step >= max_steps
                if early_stop:
                    break
                if step % eval_steps == 0:
                    self._flush_losses(pending)
                    train_loss, n_samples = 0.0, 0
                kw = dict(self.input_batch_kwargs(batch))
                kw["labels"] = self.batch_labels(batch)
                pending.append((self.train_step.step(kw), self.batch_len(batch)))  # H2D + graph replay are enqueued; nothing waits here
                if len(pending) >= 1024:  # bound the outstanding read-backs (each step owns an entry of the step's pinned loss ring)
                    tl, ns = self._flush_losses(pending)
                    train_loss, n_samples = train_loss + tl, n_samples + ns
                if (step + 1) % eval_steps == 0:
                    tl, ns = self._flush_losses(pending)
                    tl, ns = self._allreduce_sums(train_loss + tl, n_samples + ns)
                    results = dict(train_loss=tl / max(ns, 1.0))
                    if self.do_eval:
                        results.update(self.evaluate(dev_loader))
                    if self._is_main() and hasattr(eh, "set_dict_metrics"):
                        eh.set_dict_metrics(results)
                    self.logger.info("step %d (epoch %d): %s", step + 1, epoch + 1, results)
                    early_stop = self.early_stopping.step(results.get(self.early_stopping_metric), **{**results, "epoch": epoch + 1, "step": step + 1})
                    if early_stop:
                        break
        self._flush_losses(pending)
        if self.train_step is not None and hasattr(self.train_step, "synchronize"):
            self.train_step.synchronize()
        best = self.early_stopping.get_metrics()
        if best is not None and self._is_main() and hasattr(eh, "set_best"):
            eh.set_best("early_stopping", metric=self.early_stopping_metric, higher_better=True)
        test_results = None
        if self.do_test:
            test_results = self.evaluate(test_loader)
            if self._is_main() and hasattr(eh, "set_dict_metrics"):
                eh.set_dict_metrics(test_results, test=True)
        self.train_end()
        return test_results

    def train_end(self):
        """ref:vault/tmsc_utils/trainer.py:162-167 -- rank 0 only."""
        eh = self.exp_handler
        self.model = self.early_stopping.best_model()
        if not self._is_main():
            return
        if hasattr(eh, "log"):
            eh.log()
        if getattr(eh, "model_save", False) and getattr(eh, "model_save_filename", None):
            torch.save({k: v.detach().cpu() for k, v in self.model.state_dict().items()}, eh.model_save_filename)
        for fn in ("aggregate_results", "plot"):
            if hasattr(eh, fn):
                getattr(eh, fn)()

    def batch_to_device(self, batch):
        dev = self._device()
        mv = lambda v: v.to(dev, non_blocking=True) if torch.is_tensor(v) else v
        return [({k: mv(v) for k, v in e.items()} if isinstance(e, dict) else mv(e)) for e in batch]

    def evaluate(self, data_loader: DataLoader, tqdm_message: Optional[str] = None) -> Dict[str, float]:
        """This rank's shard through the inference kernels; predictions / labels / loss sums gathered from all ranks."""
        if self.train_step is not None and hasattr(self.train_step, "synchronize"):
            self.train_step.synchronize()  # the last optimizer update must have landed before weights are read
        self.model.eval()
        preds, true, loss_sum = [], [], 0.0
        for batch in data_loader:
            batch = self.batch_to_device(batch)
            with torch.no_grad():
                logits = self.get_logits_from_model(self.model(**self.input_batch_kwargs(batch)), batch, data_loader)
            labels = self.batch_labels(batch)
            loss = self.calculate_loss(logits, labels, train=False)
            if loss is not None:
                loss_sum += float(loss) * self.batch_len(batch)
            preds.extend(self.get_eval_preds_from_batch(logits))
            true.extend(self.get_eval_true_from_batch(labels))
        d = _dist()
        n_total = len(data_loader.dataset)
        if d is not None and self.world > 1:
            gathered = [None] * self.world
            d.all_gather_object(gathered, (preds, true, loss_sum))
            preds = [p for g in gathered for p in g[0]]
            true = [t for g in gathered for t in g[1]]
            loss_sum = sum(g[2] for g in gathered)
        results = dict(eval_loss=loss_sum / max(n_total, 1))
        results.update(self.evaluation_metrics(true, preds, data_loader=data_loader))
        self.model.train()
        return results


class VaultTrainerForTMSC(Twitter201XTrainer):
    """ref:vault/models/vault/trainer.py:15-37"""


class VaultTrainerForBloombergTwitterCorpus(Twitter201XTrainer):
    """ref:vault/models/vault/trainer.py:40-87 -- one logit, BCE-with-logits, model selection on eval_loss; batches are
    (inputs dict, labels) (ref:vault/vl_utils/trainer.py:13-27)."""

    early_stopping_metric = "eval_loss"
    higher_better = False
    loss_kind = "bce"

    def input_batch_kwargs(self, batch):
        return batch[0]

    def batch_len(self, batch) -> int:
        return len(next(iter(batch[0].values())))

    def calculate_loss(self, logits, labels, train):
        return torch.nn.functional.binary_cross_entropy_with_logits(logits, labels.to(logits.dtype))

    def get_eval_preds_from_batch(self, logits):
        return (logits.sigmoid() >= 0.5).int().tolist()

    def get_eval_true_from_batch(self, labels):
        return labels.int().tolist()

    def evaluation_metrics(self, eval_true, eval_preds, data_loader=None):
        from sklearn.metrics import precision_recall_fscore_support

        acc = float(np.mean([p == t for p, t in zip(eval_preds, eval_true)]))
        _, _, f1, _ = precision_recall_fscore_support(eval_true, eval_preds, average="weighted", zero_division=0)
        return dict(eval_accuracy=acc, f1_score=float(f1))


class VaultTrainerForMVSA(VaultTrainerForBloombergTwitterCorpus):
    """ref:vault/models/vault/trainer.py:90-170 -- pre-processed labels: plain CE over 3 classes; raw annotations: the logits'
    two halves against the (text, image) label pair."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.preprocessed = bool(getattr(self.dataset, "preprocessed", True))
        self.loss_kind = "ce" if self.preprocessed else "ce2"

    def calculate_loss(self, logits, labels, train):
        ce = torch.nn.functional.cross_entropy
        if self.preprocessed:
            return ce(logits, labels)
        n = logits.shape[-1]
        return 0.5 * (ce(logits[..., : n // 2], labels[..., 0]) + ce(logits[..., n // 2:], labels[..., 1]))

    def get_eval_preds_from_batch(self, logits):
        if self.preprocessed:
            return logits.argmax(-1).tolist()
        n = logits.shape[-1]
        return [[a.argmax(-1).item(), b.argmax(-1).item()] for a, b in zip(logits[..., : n // 2], logits[..., n // 2:])]

    def get_eval_true_from_batch(self, labels):
        return labels.tolist()

    def evaluation_metrics(self, eval_true, eval_preds, data_loader=None):
        from sklearn.metrics import precision_recall_fscore_support

        if self.preprocessed:
            return super().evaluation_metrics(eval_true, eval_preds, data_loader)
        flat_t = [x for row in eval_true for x in row]
        flat_p = [x for row in eval_preds for x in row]
        _, _, f1, _ = precision_recall_fscore_support(flat_t, flat_p, average="weighted", zero_division=0)
        return dict(eval_accuracy=float(np.mean([p == t for p, t in zip(flat_p, flat_t)])), f1_score=float(f1))
