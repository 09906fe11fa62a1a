"""The reference's QC experiment driver (QC/train_egcn.py:30-206) on libgode: same flags, model table, Adam with the
linear learning-rate decay between ``schedule[0]`` and ``schedule[1]`` of the epochs, per-epoch train / validate, best
checkpoint by validation metric, final test pass.  ``eodesum`` maps to the edge-conditioned ODE model (the reference maps
it to ``UnimplementedModel``); ``--dataset synthetic`` (default here) replaces the rdkit-parsed QM9 files, see
``QC/util.py``.  Under torchrun the molecules of every batch are split over the ranks and the gradients all-reduced.

    python -m graph_odenet_b200.QC.train_egcn --model egcns2s --hidden 64 --batch-size 256 --epochs 2
"""
from __future__ import annotations

import argparse
import os

import torch
import torch.optim as optim

from . import layer_models as models
from .util import (count_params, get_metric_by_task_type, read_dataset, restricted_float, save_checkpoint, train, validate)

MODEL_DICT = {
    "egcnsum": models.EdgeGCN_K_Sum,
    "egcns2s": models.EdgeGCN_K_Set2Set,
    "ennsum": models.MPNN_ENN_K_Sum,
    "enns2s": models.MPNN_ENN_K_Set2Set,
    "eressum": models.UnimplementedModel,
    "eress2s": models.EdgeRES1_K_Set2Set,
    "eodesum": models.EdgeODE_K_Sum,            # builder extension (BASELINE config 5)
    "eodes2s": models.UnimplementedModel,
}
DATASET_PATHS = {"qm9": "./data/qm9/dsgdb9nsd/", "mutag": "./data/mutag/", "enzymes": "./data/enzymes/", "synthetic": ""}


def build_parser():
    p = argparse.ArgumentParser(description="Neural message passing")
    p.add_argument("--dataset", default="synthetic", help='"synthetic" (QM9-shaped, offline) or the reference\'s "qm9", "mutag", "enzymes"')
    p.add_argument("--dataset-type", choices=["classification", "regression"])
    p.add_argument("--dataset-path", help="custom dataset path")
    p.add_argument("--log_path", default="./log/{model}-{layers}/{dataset}/all")
    p.add_argument("--resume", default="./checkpoint/{model}-{layers}/{dataset}/all", help="path to latest checkpoint ('' = none)")
    p.add_argument("--model", choices=sorted(MODEL_DICT), default="egcnsum")
    p.add_argument("--hidden", type=int, default=73, metavar="H")
    p.add_argument("--batch-size", type=int, default=20, metavar="B")
    p.add_argument("--layers", type=int, default=3, metavar="L")
    p.add_argument("--s2s", type=int, default=4, metavar="S")
    p.add_argument("--no-cuda", action="store_true", default=False, help="unsupported: graph-odenet_b200 has no CPU path")
    p.add_argument("--epochs", type=int, default=10, metavar="N")
    p.add_argument("--lr", type=lambda x: restricted_float(x, [1e-5, 1e-2]), default=1e-3, metavar="LR")
    p.add_argument("--lr-decay", type=lambda x: restricted_float(x, [.01, 1]), default=0.6, metavar="LR-DECAY")
    p.add_argument("--weight-decay", type=lambda x: restricted_float(x, [0, 1]), default=5e-4, metavar="WEIGHT-DECAY")
    p.add_argument("--schedule", type=list, default=[0.2, 0.8], metavar="S")
    p.add_argument("--momentum", type=float, default=0.9, metavar="M")
    p.add_argument("--log-interval", type=int, default=20, metavar="N")
    p.add_argument("--prefetch", type=int, default=8, help="pre-fetching threads (file-based datasets)")
    p.add_argument("--synthetic-batches", type=int, default=8, help="training batches per epoch of the synthetic dataset")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.no_cuda or not torch.cuda.is_available():
        raise RuntimeError("graph-odenet_b200 has no CPU path: a CUDA device is required")
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        from .. import parallel
        parallel.init_process_group(dev)
    root = args.dataset_path if args.dataset_path else DATASET_PATHS.get(args.dataset, "")
    resume_dir = args.resume.format(dataset=args.dataset, model=args.model, layers=args.layers) if args.resume else None
    Model_Class = MODEL_DICT[args.model]

    print("Preparing dataset")
    nb = args.synthetic_batches
    node_features, edge_features, target_features, task_type, train_loader, valid_loader, test_loader = read_dataset(
        args.dataset, root, args.batch_size, args.prefetch, hidden=args.hidden, device=dev, rank=rank, world=world,
        n_batches=(nb, max(nb // 4, 1), max(nb // 4, 1)))
    task_type = args.dataset_type or task_type

    print("\tCreate model")
    torch.manual_seed(0)                       # identical replicas on every rank
    model = Model_Class(node_features=node_features, edge_features=edge_features, target_features=target_features,
                        hidden_features=args.hidden, num_layers=args.layers, dropout=0.5, type=task_type,
                        s2s_processing_steps=args.s2s)
    print("#Parameters: {param_count}".format(param_count=count_params(model)))
    print("Optimizer")
    optimizer = optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay)
    criterion, evaluation, metric_name, metric_compare, metric_best = get_metric_by_task_type(task_type, target_features)
    lr_step = (args.lr - args.lr * args.lr_decay) / (args.epochs * args.schedule[1] - args.epochs * args.schedule[0])
    best_er1 = 0
    model = model.to(dev)
    criterion = criterion.to(dev)
    history = []
    for epoch in range(args.epochs):
        if args.epochs * args.schedule[0] < epoch < args.epochs * args.schedule[1]:
            args.lr -= lr_step
            for group in optimizer.param_groups:
                group["lr"] = args.lr
        loss, t_batch = train(train_loader, model, criterion, optimizer, epoch, evaluation, None, metric_name=metric_name,
                              log_interval=args.log_interval, world=world, global_batch=args.batch_size)
        er1 = validate(valid_loader, model, criterion, evaluation, None, metric_name=metric_name, log_interval=args.log_interval,
                       world=world)
        is_best = metric_compare(er1, best_er1)       # the reference starts best_er1 at 0 for both metric directions
        best_er1 = metric_best(er1, best_er1)
        history.append((loss, er1, t_batch))
        if resume_dir and rank == 0:
            save_checkpoint({"epoch": epoch + 1, "state_dict": model.state_dict(), "best_er1": best_er1,
                             "optimizer": optimizer.state_dict()}, is_best=is_best, directory=resume_dir)
    test_metric = validate(test_loader, model, criterion, evaluation, metric_name=metric_name, log_interval=args.log_interval,
                           world=world)
    if world > 1:
        torch.distributed.destroy_process_group()
    return {"history": history, "test": test_metric, "params": count_params(model)}


if __name__ == "__main__":
    main()
