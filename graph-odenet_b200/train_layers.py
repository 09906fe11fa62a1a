"""The reference's layer-depth sweep ``train_layers.py`` (GCN/train_layers.py:1-183, GAT/train_layers.py) on libgode.

For every model of the depth-parameterised families (``GCNK`` ... ``ODEK2``), every depth in
``[layers_min, layers_max]`` and ``--runs`` seeds, train for at most ``--epochs`` epochs, record the validation curve,
mark the epoch at which the dataset's hard-coded thresholds are met (GCN/train_layers.py:119-127,166: the curve is
frozen at the previous epoch's value from there on) and the test metrics, and write one **result pickle per model**,
``{dataset}_{model}.pickle`` (GCN/train_layers.py:180-181) -- the file ``report_layers.py`` plots from.

On-disk format (what SURVEY 8f.4 asks to keep): ``pickle.HIGHEST_PROTOCOL`` of a dict with

    layer_val_acc, layer_val_loss   float64 [layers_max + 1, runs, epochs]   (rows below layers_min stay zero)
    layer_convergence               float64 [layers_max + 1, runs]           (initialised to ``epochs``)
    layer_test_acc, layer_test_loss float64 [layers_max + 1, runs]
    min_layers, max_layers          int  (max_layers = --layers_max + 1, the exclusive bound the reference stores;
                                          min_layers is raised to nlayers + 1 when a model refuses a depth)

Same flags as the reference plus ``--data-root`` / ``--npz`` / ``--out-dir`` / ``--models`` (builder extensions;
defaults reproduce the reference, which sweeps all eight models into the working directory).
"""
from __future__ import annotations

import argparse
import os
import pickle

import numpy as np
import torch
import torch.nn.functional as F
import torch.optim as optim

from . import utils

MODEL_NAMES = ["GCNK", "GCNKnorm", "RESK1", "RESK2", "RESK1norm", "RESK2norm", "ODEK1", "ODEK2"]   # GCN/train_layers.py:49-58

# (validation accuracy, validation loss) anchors of the two-layer GCN the thresholds derive from, GCN/train_layers.py:119-127
THRESHOLDS = {"cora": (0.7782, 0.7929), "citeseer": (0.6443, 1.2454), "pubmed": (0.7726, 0.7136)}


def thresholds(dataset):
    """A run has converged at the first epoch with acc_val > 0.9 * anchor accuracy and loss_val < 1.1 * anchor loss."""
    acc, loss = THRESHOLDS[dataset]
    return acc * 0.9, loss * 1.1


def new_result_table(layers_min, layers_max_exclusive, runs, epochs):
    """The per-model record of GCN/train_layers.py:131-141 (``layers_max_exclusive`` = --layers_max + 1)."""
    return {
        "layer_val_acc": np.zeros([layers_max_exclusive, runs, epochs]),
        "layer_val_loss": np.zeros([layers_max_exclusive, runs, epochs]),
        "layer_convergence": epochs * np.ones([layers_max_exclusive, runs]),
        "layer_test_acc": np.zeros([layers_max_exclusive, runs]),
        "layer_test_loss": np.zeros([layers_max_exclusive, runs]),
        "min_layers": layers_min,
        "max_layers": layers_max_exclusive,
    }


def record_epoch(rec, nlayers, run, epoch, val_loss, val_acc, acc_threshold, loss_threshold):
    """Store one epoch's validation metrics; on convergence freeze the rest of the curve at the PREVIOUS epoch's values
    (the reference indexes ``epoch - 1``, which at epoch 0 is the still-zero last column) and say so."""
    rec["layer_val_loss"][nlayers, run, epoch] = val_loss
    rec["layer_val_acc"][nlayers, run, epoch] = val_acc
    if val_acc > acc_threshold and val_loss < loss_threshold:
        rec["layer_convergence"][nlayers, run] = epoch
        rec["layer_val_loss"][nlayers, run, epoch:] = rec["layer_val_loss"][nlayers, run, epoch - 1]
        rec["layer_val_acc"][nlayers, run, epoch:] = rec["layer_val_acc"][nlayers, run, epoch - 1]
        return True
    return False


def save_results(rec, dataset, model, out_dir="."):
    path = os.path.join(out_dir, "{dataset}_{model}.pickle".format(dataset=dataset, model=model))
    with open(path, "wb") as f:
        pickle.dump(rec, f, protocol=pickle.HIGHEST_PROTOCOL)
    return path


def load_results(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--no-cuda", action="store_true", default=False, help="Disables CUDA training (unsupported: no CPU path).")
    p.add_argument("--fastmode", action="store_true", default=False, help="Validate during training pass.")
    p.add_argument("--seed", type=int, default=42, help="Random seed.")
    p.add_argument("--epochs", type=int, default=200, help="Number of epochs to train.")
    p.add_argument("--runs", type=int, default=500, help="Number of times to train and evaluate the model.")
    p.add_argument("--lr", type=float, default=0.01, help="Initial learning rate.")
    p.add_argument("--weight_decay", type=float, default=5e-4, help="Weight decay (L2 loss on parameters).")
    p.add_argument("--hidden", type=int, default=16, help="Number of hidden units.")
    p.add_argument("--layers_min", type=int, default=3, help="Minimum number of layers to test with")
    p.add_argument("--layers_max", type=int, default=5, help="Maximum number of layers to test with")
    p.add_argument("--dropout", type=float, default=0.5, help="Dropout rate (1 - keep probability).")
    p.add_argument("--dataset", choices=["cora", "citeseer", "pubmed"], default="cora", help="Which dataset to use")
    p.add_argument("--early_stopping_epochs", type=int, default=10, help="Number of epochs to evaluate early stopping on.")
    p.add_argument("--early_stopping_threshold", type=float, default=1e-10,
                   help="Minimum decrease in validation loss over last early_stopping_epochs.")
    # builder extensions
    p.add_argument("--models", nargs="*", default=None, choices=MODEL_NAMES, help="subset of the eight models (default: all)")
    p.add_argument("--out-dir", default=".", help="where the result pickles go (the reference writes to the working directory)")
    p.add_argument("--data-root", default=None, help="directory holding ind.<dataset>.* (default $GODE_DATA or ./data)")
    p.add_argument("--npz", default=None, help="loader output saved as .npz (tests/golden/planetoid_<ds>.npz)")
    return p


def main(family="GCN", argv=None, out=print):
    args = build_parser().parse_args(argv)
    if args.no_cuda or not torch.cuda.is_available():
        raise RuntimeError("graph-odenet_b200 has no CPU path: a CUDA device is required")
    assert args.layers_min < args.layers_max
    layers_max = args.layers_max + 1
    if family == "GAT":
        from .GAT import models
    else:
        from .GCN import models
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)

    if args.npz:
        data = utils.load_npz(args.npz, family)
    elif family == "GAT":
        data = utils.load_data_gat(args.dataset, args.data_root)
    else:
        data = utils.load_data_new(args.dataset, args.data_root)
    data = tuple(t.cuda() for t in data)
    graph, (features, labels, idx_train, idx_val, idx_test) = data[:-5], data[-5:]
    acc_threshold, loss_threshold = thresholds(args.dataset)

    def train(model, optimizer):
        model.train()
        optimizer.zero_grad()
        output = model(features, *graph)
        loss_train = F.nll_loss(output[idx_train], labels[idx_train])
        loss_train.backward()
        optimizer.step()
        if not args.fastmode:
            model.eval()
            with torch.no_grad():
                output = model(features, *graph)
        loss_val = F.nll_loss(output[idx_val], labels[idx_val])
        acc_val = utils.accuracy(output[idx_val], labels[idx_val])
        return loss_val.item(), acc_val.item()

    def test(model):
        model.eval()
        with torch.no_grad():
            output = model(features, *graph)
        return (F.nll_loss(output[idx_test], labels[idx_test]).item(),
                utils.accuracy(output[idx_test], labels[idx_test]).item())

    model_data, paths = {}, {}
    for m in (args.models or MODEL_NAMES):
        Model = getattr(models, m)
        rec = model_data[m] = new_result_table(args.layers_min, layers_max, args.runs, args.epochs)
        for nlayers in range(args.layers_min, layers_max):
            for run in range(args.runs):
                try:
                    model = Model(nfeat=features.shape[1], nhid=args.hidden, nclass=labels.max().item() + 1,
                                  dropout=args.dropout, nlayers=nlayers)
                except ValueError:
                    rec["min_layers"] = nlayers + 1      # can't build a residual network with that many blocks
                    continue
                model = model.cuda()
                optimizer = optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay)
                for epoch in range(args.epochs):
                    val_loss, val_acc = train(model, optimizer)
                    if record_epoch(rec, nlayers, run, epoch, val_loss, val_acc, acc_threshold, loss_threshold):
                        break
                rec["layer_test_loss"][nlayers, run], rec["layer_test_acc"][nlayers, run] = test(model)
                out("{nlayers} layers's run #{run} Test -- epochs: {epochs:d} acc: {acc:.2f}%".format(
                    nlayers=nlayers, run=run, epochs=int(rec["layer_convergence"][nlayers, run]),
                    acc=100 * rec["layer_test_acc"][nlayers, run]), flush=True)
        out("Optimization with model \"{model}\" on dataset \"{dataset}\" Finished!".format(model=m, dataset=args.dataset),
            flush=True)
        paths[m] = save_results(rec, args.dataset, m, args.out_dir)
    return {"results": model_data, "paths": paths}
