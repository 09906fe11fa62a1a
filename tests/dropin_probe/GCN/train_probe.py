"""A stand-in for the reference's ``GCN/train_res.py`` on the GPU box (where /root/reference does not exist): it imports its
siblings by BARE name exactly as the reference's scripts do (GCN/train_res.py:13-14, GCN/models.py:4-5) and trains two
epochs on the Cora fixture.  Run through dropin/run_reference.py by tests/test_dropin.py."""
import sys

import numpy as np
import torch
import torch.nn.functional as F
import torch.optim as optim

import models                                   # bare name, as GCN/train_res.py:14
from layers import GraphConvolution             # bare name, as GCN/models.py:4
from utils import accuracy, count_params        # bare name, as GCN/train_res.py:13

fix = np.load(sys.argv[1])
n = int(fix["n"])
import scipy.sparse as sp
feats = torch.from_numpy(np.asarray(sp.csr_matrix((fix["feat_data"], fix["feat_indices"], fix["feat_indptr"]),
                                                  shape=(n, int(fix["nfeat"]))).todense(), dtype=np.float32)).cuda()
adj = torch.sparse_coo_tensor(torch.from_numpy(np.vstack([fix["coo_row"], fix["coo_col"]]).astype(np.int64)),
                              torch.from_numpy(fix["coo_val"]), (n, n)).cuda()
labels = torch.from_numpy(fix["labels"].astype(np.int64)).cuda()
idx_train = torch.from_numpy(fix["idx_train"].astype(np.int64)).cuda()
torch.manual_seed(42)
model = models.ODEGCN3(nfeat=feats.shape[1], nhid=16, nclass=int(labels.max()) + 1, dropout=0.5).cuda()
assert isinstance(model.gc1, GraphConvolution)
opt = optim.Adam(model.parameters(), lr=0.01, weight_decay=5e-4)
losses = []
for epoch in range(2):
    model.train()
    opt.zero_grad()
    model.nfe = 0
    out = model(feats, adj)
    loss = F.nll_loss(out[idx_train], labels[idx_train])
    nfe_f = model.nfe
    model.nfe = 0
    loss.backward()
    opt.step()
    losses.append(float(loss))
    print("epoch %d loss %.4f acc %.4f nfe %d/%d params %d" % (epoch, losses[-1], float(accuracy(out[idx_train], labels[idx_train])),
                                                              nfe_f, model.nfe, count_params(model)))
assert all(np.isfinite(losses))
print("PROBE OK layers=%s models=%s" % (sys.modules["layers"].__file__, sys.modules["models"].__file__))
