"""Helpers of the QC experiment driver with the reference's names (QC/util.py:19-275): ``save_checkpoint``,
``get_metric_by_task_type``, ``restricted_float``, ``count_params``, ``read_dataset``, ``train``, ``validate`` and
``AverageMeter`` (QC/LogMetric.py:24-42, whose running average is kept as it is, quirk included).

``read_dataset`` differs in one point.  The reference parses QM9 / MUTAG files with rdkit and networkx in DataLoader
workers (QC/datasets/, QC/GraphReader/): neither the data nor rdkit exist offline, so those names raise
``NotImplementedError`` with that explanation, and ``dataset="synthetic"`` yields seeded QM9-shaped batches
(``synth.qm9_like_batch``: 13 node features, 5 edge features, 12 normalised targets) already resident on the device, in the
tuple layout of ``collate_g_concat_edge_data`` -- ``(batch_size, g, b, x, e_d, e_src, e_tgt, target)`` with ``e_tgt`` as the
index vector of the one-hot ``Etgt`` (the layers accept dense, sparse or index form; the dense form is O(N * E)) and
``g = None`` (the dense adjacency is never read by the models).

Data-parallel runs (BASELINE config 5): under torchrun every rank draws its own shard of every global batch
(``parallel.shard_molecules``) and ``train`` all-reduces the gradients once per step (``parallel.allreduce_gradients``)
weighted by local / global batch size, so the update is the one of the global-batch mean loss.
"""
from __future__ import annotations

import argparse
import os
import shutil
import time

import torch
from torch import nn


class AverageMeter:
    """QC/LogMetric.py:24-42."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.avg, self.count = 0, 0, 0

    def update(self, val, n=1):
        self.val = val
        if self.count:
            self.avg = self.avg * (self.count - n) / self.count + val * n / self.count
        else:
            self.avg = val
        self.count += n


def save_checkpoint(state, is_best, directory):
    os.makedirs(directory, exist_ok=True)
    checkpoint_file = os.path.join(directory, "checkpoint.pth")
    torch.save(state, checkpoint_file)
    if is_best:
        shutil.copyfile(checkpoint_file, os.path.join(directory, "model_best.pth"))


def get_metric_by_task_type(task_type, target_features):
    """(criterion, evaluation, metric name, is-better, best-of) -- QC/util.py:29-54."""
    if task_type == "regression":
        return (nn.MSELoss(reduction="mean"), lambda out, tgt: torch.mean(torch.abs(out - tgt)), "Error ratio",
                lambda x, y: x < y, min)
    if task_type == "classification":
        if target_features == 1:
            return (nn.BCEWithLogitsLoss(reduction="mean"),
                    lambda out, tgt: torch.mean(torch.sigmoid(out).round().eq(tgt).double()), "Accuracy", lambda x, y: x > y, max)
        return (nn.NLLLoss(reduction="mean"), lambda out, tgt: torch.mean(out.max(1)[1].type_as(tgt).eq(tgt).double()),
                "Accuracy", lambda x, y: x > y, max)
    raise ValueError("Unrecognised task type")


def restricted_float(x, inter):
    x = float(x)
    if x < inter[0] or x > inter[1]:
        raise argparse.ArgumentTypeError("%r not in range [%g, %g]" % (x, inter[0], inter[1]))
    return x


def count_params(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


class SyntheticLoader:
    """Seeded QM9-shaped batches on the device; ``len`` batches per epoch, the same batches every epoch (as a fixed
    dataset without shuffling would give).  With ``world > 1`` every rank holds its shard of each global batch."""

    def __init__(self, n_batches, batch_size, hidden, seed, device, rank=0, world=1):
        from .. import parallel, synth
        self.batches, self.global_batch = [], batch_size
        lo, hi = parallel.shard_molecules(batch_size, rank, world)
        for i in range(n_batches):
            b = synth.qm9_like_batch(batch_size, hidden, seed=seed + i, device=device)
            tgt = torch.randn(batch_size, 12, device=device, generator=torch.Generator(device=device).manual_seed(seed + 7919 * (i + 1)))
            if world > 1:                          # keep the molecules [lo, hi) of the global batch, ids re-based to 0
                keep_n = (b["batch"] >= lo) & (b["batch"] < hi)
                new_id = torch.cumsum(keep_n.to(torch.int64), 0) - 1
                keep_e = keep_n[b["esrc"]]
                b = {"node_features": b["node_features"][keep_n], "edge_features": b["edge_features"][keep_e],
                     "esrc": new_id[b["esrc"][keep_e]], "etgt": new_id[b["etgt"][keep_e]], "batch": b["batch"][keep_n] - lo}
                tgt = tgt[lo:hi]
            self.batches.append((hi - lo, None, b["batch"], b["node_features"], b["edge_features"], b["esrc"], b["etgt"], tgt))

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return iter(self.batches)


def read_dataset(dataset, root, batch_size, num_workers, hidden=73, device="cuda", rank=0, world=1, n_batches=(8, 2, 2)):
    """-> (node_features, edge_features, target_features, task_type, train_loader, valid_loader, test_loader)."""
    if dataset == "synthetic":
        mk = lambda n, seed: SyntheticLoader(n, batch_size, hidden, seed, device, rank, world)
        return 13, 5, 12, "regression", mk(n_batches[0], 0), mk(n_batches[1], 100_000), mk(n_batches[2], 200_000)
    if dataset in ("qm9", "mutag"):
        raise NotImplementedError("the %s reader of the reference (QC/datasets, QC/GraphReader) needs rdkit and the downloaded "
                                  "files under %r: neither exists offline -- use --dataset synthetic" % (dataset, root))
    if dataset == "enzymes":
        raise NotImplementedError("Enzymes not yet implemented")
    raise NotImplementedError("General loading not yet implemented")


def _slice_target(target, target_range):
    if target_range is None:
        return target
    if len(target_range) == 1:
        return target[:, target_range[0]]
    return target[:, slice(*target_range)]


def train(train_loader, model, criterion, optimizer, epoch, evaluation, logger=None, target_range=(0, None), tgt_name="",
          metric_name="metric", cuda=True, log_interval=20, world=1, global_batch=None):
    from .. import parallel
    batch_time, data_time, losses, metric = AverageMeter(), AverageMeter(), AverageMeter(), AverageMeter()
    model.train()
    end = time.time()
    for i, (batch_size, g, b, x, e_d, e_src, e_tgt, target) in enumerate(train_loader):
        target = _slice_target(target, target_range)
        data_time.update(time.time() - end)
        optimizer.zero_grad()
        output = model(node_features=x, edge_features=e_d, Esrc=e_src, Etgt=e_tgt, batch=b)
        train_loss = criterion(output, target)
        losses.update(train_loss.item(), batch_size)
        metric.update(evaluation(output, target).item(), batch_size)
        train_loss.backward()
        if world > 1:       # gradient of the global-batch mean: local means weighted by their share of the batch
            parallel.allreduce_gradients(model.parameters(), local_weight=batch_size / float(global_batch or batch_size * world))
        optimizer.step()
        batch_time.update(time.time() - end)
        end = time.time()
        if i % log_interval == 0 and i > 0:
            print("Epoch: [{0}][{1}/{2}]\t"
                  "Time {batch_time.val:.3f} ({batch_time.avg:.3f})\t"
                  "Data {data_time.val:.3f} ({data_time.avg:.3f})\t"
                  "Loss {loss.val:.4f} ({loss.avg:.4f})\t"
                  "{metric_name} {metric.val:.4f} ({metric.avg:.4f})"
                  .format(epoch, i, len(train_loader), batch_time=batch_time, data_time=data_time, loss=losses,
                          metric_name=metric_name, metric=metric), flush=True)
    if logger is not None:
        logger.log_value("train_epoch_loss", losses.avg)
        logger.log_value("train_epoch_{metric}".format(metric=metric_name), metric.avg)
    print("Epoch: [{0}] {tgt_name} Avg {metric_name} {metric.avg:.3f}; Average Loss {loss.avg:.3f}; Avg Time x Batch {b_time.avg:.3f}"
          .format(epoch, metric_name=metric_name, metric=metric, loss=losses, b_time=batch_time, tgt_name=tgt_name), flush=True)
    return losses.avg, batch_time.avg


def validate(val_loader, model, criterion, evaluation, logger=None, target_range=(0, None), tgt_name="", metric_name="metric",
             cuda=True, log_interval=20, world=1):
    batch_time, losses, metric = AverageMeter(), AverageMeter(), AverageMeter()
    model.eval()
    with torch.no_grad():
        end = time.time()
        for i, (batch_size, g, b, x, e_d, e_src, e_tgt, target) in enumerate(val_loader):
            target = _slice_target(target, target_range)
            output = model(node_features=x, edge_features=e_d, Esrc=e_src, Etgt=e_tgt, batch=b)
            losses.update(criterion(output, target).item(), batch_size)
            metric.update(evaluation(output, target).item(), batch_size)
            batch_time.update(time.time() - end)
            end = time.time()
            if i % log_interval == 0 and i > 0:
                print("Test: [{0}/{1}]\t"
                      "Time {batch_time.val:.3f} ({batch_time.avg:.3f})\t"
                      "Loss {loss.val:.4f} ({loss.avg:.4f})\t"
                      "{metric_name} {metric.val:.4f} ({metric.avg:.4f})"
                      .format(i, len(val_loader), batch_time=batch_time, loss=losses, metric_name=metric_name, metric=metric),
                      flush=True)
    if world > 1:
        # every rank validated its own shard: the metric (and the checkpoint decision taken on it) is the global mean
        from .. import parallel
        dev = next(model.parameters()).device
        metric.avg = parallel.allreduce_mean(metric.avg * metric.count, metric.count, dev)
        losses.avg = parallel.allreduce_mean(losses.avg * losses.count, losses.count, dev)
    print(" * {tgt_name} Average {metric_name} {metric.avg:.3f}; Average Loss {loss.avg:.3f}"
          .format(metric_name=metric_name, metric=metric, loss=losses, tgt_name=tgt_name), flush=True)
    if logger is not None:
        logger.log_value("test_epoch_loss", losses.avg)
        logger.log_value("test_epoch_{metric}".format(metric=metric_name), metric.avg)
    return metric.avg
