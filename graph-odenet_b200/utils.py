"""Data loading and small helpers with the reference's names and return tuples (GCN/utils.py:134-229,
GAT/utils.py:134-216).  Host-side, one-off work (pickle -> networkx -> scipy -> torch); not on the hot path.

``load_data_new(dataset, data_root=...)`` returns ``(adj, features, labels, idx_train, idx_val, idx_test)`` exactly as
GCN/utils.py:202 -- ``adj`` the row-normalised D^-1 (A + I) as an uncoalesced fp32 sparse COO tensor with int64
indices; ``load_data_gat`` returns ``(src, tgt, Mtgt, features, labels, idx_train, idx_val, idx_test)`` as
GAT/utils.py:216.  The reference resolves ``data/`` relative to the working directory; here the directory is an
argument (default: ``$GODE_DATA`` or ``./data``).  ``*.npz`` snapshots of the loader output (tests/golden/) load through
``load_npz`` -- that is what the GPU tests use, since the pickles do not travel to the GPU box.
"""
from __future__ import annotations

import os
import pickle
import sys

import numpy as np
import torch


def _data_root(data_root):
    return data_root or os.environ.get("GODE_DATA") or "data"


def parse_index_file(filename):
    return [int(line.strip()) for line in open(filename)]


def normalize(mx):
    """Row-normalise a scipy sparse matrix in float64: D^-1 mx, inf -> 0 (GCN/utils.py:205-212)."""
    import scipy.sparse as sp
    rowsum = np.array(mx.sum(1))
    with np.errstate(divide="ignore"):
        r_inv = np.power(rowsum, -1.0).flatten()
    r_inv[np.isinf(r_inv)] = 0.0
    return sp.diags(r_inv).dot(mx)


def sparse_mx_to_torch_sparse_tensor(sparse_mx):
    """scipy sparse -> torch sparse COO, fp32 values / int64 indices, uncoalesced (GCN/utils.py:222-229)."""
    m = sparse_mx.tocoo().astype(np.float32)
    idx = torch.from_numpy(np.vstack((m.row, m.col)).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(m.data), torch.Size(m.shape))


def accuracy(output, labels):
    preds = output.max(1)[1].type_as(labels)
    return preds.eq(labels).double().sum() / len(labels)


def count_params(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def _planetoid(dataset_str, data_root):
    """The Planetoid pickles, merged and re-ordered as GCN/utils.py:154-183 does."""
    import networkx as nx
    import scipy.sparse as sp
    root = _data_root(data_root)
    objs = []
    for name in ("x", "y", "tx", "ty", "allx", "ally", "graph"):
        with open(os.path.join(root, "ind.%s.%s" % (dataset_str, name)), "rb") as f:
            objs.append(pickle.load(f, encoding="latin1") if sys.version_info > (3, 0) else pickle.load(f))
    x, y, tx, ty, allx, ally, graph = objs
    reorder = parse_index_file(os.path.join(root, "ind.%s.test.index" % dataset_str))
    rng = np.sort(reorder)
    if dataset_str == "citeseer":      # isolated test nodes: zero rows at their positions
        full = range(min(reorder), max(reorder) + 1)
        tx_ext = sp.lil_matrix((len(full), x.shape[1]))
        tx_ext[rng - min(rng), :] = tx
        ty_ext = np.zeros((len(full), y.shape[1]))
        ty_ext[rng - min(rng), :] = ty
        tx, ty = tx_ext, ty_ext
    features = sp.vstack((allx, tx)).tolil()
    features[reorder, :] = features[rng, :]
    labels = np.vstack((ally, ty))
    labels[reorder, :] = labels[rng, :]
    G = nx.from_dict_of_lists(graph)
    return G, features, labels, len(y), rng


def _tensors(features, labels, n_train, test_range):
    feats = torch.FloatTensor(np.array(normalize(features).todense()))
    lab = torch.LongTensor(np.argmax(labels, axis=1))      # argmax: stable where a row is all zero (citeseer)
    return (feats, lab, torch.LongTensor(list(range(n_train))), torch.LongTensor(list(range(n_train, n_train + 500))),
            torch.LongTensor(test_range.tolist()))


def load_data_new(dataset_str="cora", data_root=None):
    import networkx as nx
    import scipy.sparse as sp
    G, features, labels, n_train, rng = _planetoid(dataset_str, data_root)
    adj = nx.adjacency_matrix(G)
    adj = normalize(adj + sp.eye(adj.shape[0]))
    feats, lab, itr, iva, ite = _tensors(features, labels, n_train, rng)
    return sparse_mx_to_torch_sparse_tensor(adj), feats, lab, itr, iva, ite


def load_data_gat(dataset_str="cora", data_root=None):
    """GAT/utils.py:134-216: every undirected edge once (networkx edge order), no self-loops; Mtgt[tgt_e, e] = 1."""
    import scipy.sparse as sp
    G, features, labels, n_train, rng = _planetoid(dataset_str, data_root)
    edges = np.array(G.edges, dtype=np.int64).reshape(-1, 2)
    E = edges.shape[0]
    Mtgt = sp.coo_matrix((np.ones(E), (edges[:, 1], np.arange(E))), shape=(labels.shape[0], E), dtype=np.float32)
    feats, lab, itr, iva, ite = _tensors(features, labels, n_train, rng)
    return (torch.from_numpy(edges[:, 0].copy()), torch.from_numpy(edges[:, 1].copy()), sparse_mx_to_torch_sparse_tensor(Mtgt),
            feats, lab, itr, iva, ite)


def load_npz(path, family="GCN"):
    """Loader output saved by tests/golden/make_golden.py (planetoid_<ds>.npz) -> the same tuples as above."""
    import scipy.sparse as sp
    c = np.load(path)
    n = int(c["n"])
    feats = sp.csr_matrix((c["feat_data"], c["feat_indices"], c["feat_indptr"]), shape=(n, int(c["nfeat"])))
    feats = torch.from_numpy(np.asarray(feats.todense(), dtype=np.float32))
    lab = torch.from_numpy(c["labels"].astype(np.int64))
    itr, iva, ite = (torch.from_numpy(c[k].astype(np.int64)) for k in ("idx_train", "idx_val", "idx_test"))
    if family == "GAT":
        src, tgt = torch.from_numpy(c["gat_src"].astype(np.int64)), torch.from_numpy(c["gat_tgt"].astype(np.int64))
        E = src.numel()
        Mtgt = torch.sparse_coo_tensor(torch.stack([tgt, torch.arange(E)]), torch.ones(E), (n, E))
        return src, tgt, Mtgt, feats, lab, itr, iva, ite
    idx = torch.from_numpy(np.vstack([c["coo_row"], c["coo_col"]]).astype(np.int64))
    adj = torch.sparse_coo_tensor(idx, torch.from_numpy(c["coo_val"].astype(np.float32)), (n, n))
    return adj, feats, lab, itr, iva, ite
