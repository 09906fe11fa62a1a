"""CPU, world_size 2 and 3 over gloo: the index logic of the row partition (bit-exact) and the halo exchange.

The reference has no distributed code, so the bar is "P ranks == 1 rank on the same graph":
 * each rank's renumbered block, un-renumbered, is exactly the global COO restricted to its rows (and, for the
   transpose block, to its columns);
 * after ``HaloPlan.exchange`` the halo tail of the operand holds exactly the peers' rows;
 * block SpMM on the exchanged operand equals the global product's rows (fp32 sums, different order: 1e-6).
No libgode compute call is made here (no GPU): the pack step is passed in as an index_select.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from graph_odenet_b200 import parallel, synth
        row, col, val = synth.powerlaw_graph(n, avg_degree=8, locality=0.7, window=64, seed=3, device="cpu")
        bounds = parallel.partition_bounds(n, world)
        assert bounds[0] == 0 and bounds[-1] == n and all(b1 >= b0 for b0, b1 in zip(bounds, bounds[1:]))
        lo, hi = bounds[rank], bounds[rank + 1]
        n_own = hi - lo
        g = torch.Generator().manual_seed(11)
        X = torch.randn(n, d, generator=g)
        A = torch.sparse_coo_tensor(torch.stack([row, col]), val, (n, n)).coalesce()
        pack = lambda buf, idx: buf[idx.long()].contiguous()
        for transpose in (False, True):
            r, c, v, halo = parallel.local_block(row, col, val, bounds, rank, transpose=transpose)
            # (1) bit-exact partition: the un-renumbered block is the global COO restricted to the owned rows
            gr, gc = (col, row) if transpose else (row, col)
            m = (gr >= lo) & (gr < hi)
            assert torch.equal(r + lo, gr[m])
            assert torch.equal(parallel.global_cols(c, halo, lo, n_own), gc[m])
            assert torch.equal(v.view(torch.int32), val[m].view(torch.int32))
            assert torch.equal(halo, torch.unique(halo)) and not ((halo >= lo) & (halo < hi)).any()
            assert int(c.max()) < n_own + halo.numel()
            # (2) halo exchange
            hp = parallel.HaloPlan(halo, bounds, rank)
            assert sum(hp.recv_counts) == halo.numel()
            buf = torch.full((n_own + hp.n_halo, d), float("nan"))
            buf[:n_own] = X[lo:hi]
            hp.exchange(buf, pack=pack)
            assert torch.equal(buf[n_own:], X[halo])
            # (3) block product == rows of the global product
            blk = torch.sparse_coo_tensor(torch.stack([r, c]), v, (n_own, n_own + hp.n_halo)).coalesce()
            got = torch.sparse.mm(blk, buf)
            want = torch.sparse.mm(A.t().coalesce() if transpose else A, X)[lo:hi]
            assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
        # what every rank sends must be what its peers expect: totals agree across the group
        tot = torch.tensor([sum(hp.send_counts), sum(hp.recv_counts)], dtype=torch.int64)
        dist.all_reduce(tot)
        assert int(tot[0]) == int(tot[1])
        out_q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out_q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_halo_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1501, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(m == "ok" for _, m in res), res


def test_partition_bounds_cover_everything():
    from graph_odenet_b200 import parallel
    for n in (0, 1, 7, 1000, 10_000_001):
        for w in (1, 2, 3, 8):
            b = parallel.partition_bounds(n, w)
            sizes = np.diff(b)
            assert b[0] == 0 and b[-1] == n and sizes.min() >= 0 and sizes.max() - sizes.min() <= 1


def test_single_rank_plan_has_no_halo():
    from graph_odenet_b200 import parallel, synth
    row, col, val = synth.powerlaw_graph(300, avg_degree=6, seed=1, device="cpu")
    r, c, v, halo = parallel.local_block(row, col, val, [0, 300], 0)
    assert halo.numel() == 0 and torch.equal(r, row) and torch.equal(c, col)
    hp = parallel.HaloPlan(halo, [0, 300], 0)
    x = torch.ones(300, 4)
    assert hp.exchange(x) is x


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from graph_odenet_b200 import parallel
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        x, y = torch.randn(11, 6), torch.randn(11, 3)          # 11 "molecules": uneven shards
        full = torch.nn.functional.mse_loss(model(x), y)
        want = torch.autograd.grad(full, list(model.parameters()))
        lo, hi = parallel.shard_molecules(11, rank, world)
        for p in model.parameters():
            p.grad = None
        torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi]).backward()
        parallel.allreduce_gradients(model.parameters(), local_weight=(hi - lo) / 11)
        ok = all(torch.allclose(p.grad, w, rtol=1e-5, atol=1e-7) for p, w in zip(model.parameters(), want))
        q.put((rank, "ok" if ok else "gradient mismatch"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: %s\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_gloo():
    """Config 5 (molecule batches, data-parallel): sharded batch-mean gradients, weighted and summed, equal the
    full-batch gradient."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(m == "ok" for _, m in res), res


@pytest.mark.parametrize("method,sweeps", [("degree", 0), ("rcm", 0), ("rcm", 4)])
def test_locality_relabelling_is_an_exact_bijection(method, sweeps):
    """the reordering pass is a pure relabelling -- perm is a permutation, relabel / unrelabel are exact inverses,
    the relabelled CSR un-relabels to the original entry set bit for bit, and it shrinks the halo of a graph whose locality
    is hidden behind shuffled ids."""
    import graph_odenet_b200  # noqa: F401
    from graph_odenet_b200 import parallel, synth
    n = 20000
    row, col, val = synth.powerlaw_graph(n, avg_degree=12, locality=0.9, window=256, seed=3, device="cpu")
    g = torch.Generator().manual_seed(1)
    shuf = torch.randperm(n, generator=g)
    r0, c0 = shuf[row], shuf[col]
    perm = parallel.locality_order(r0, c0, n, method, sweeps)
    assert torch.equal(torch.sort(perm).values, torch.arange(n))
    r1, c1 = parallel.relabel(r0, c0, perm)
    rb, cb = parallel.unrelabel(r1, c1, perm)
    assert torch.equal(rb, r0) and torch.equal(cb, c0)
    # same multiset of (row, col, value) triples after un-relabelling, whatever order the relabelled matrix is stored in
    o1 = torch.argsort(r1 * n + c1)
    rr, cc = parallel.unrelabel(r1[o1], c1[o1], perm)
    key_a = torch.sort(rr * n + cc)
    key_b = torch.sort(r0 * n + c0)
    assert torch.equal(key_a.values, key_b.values)
    assert torch.equal(val[o1][key_a.indices].view(torch.int32), val[key_b.indices].view(torch.int32))
    before, after = parallel.halo_fraction(r0, c0, n, 8), parallel.halo_fraction(r1, c1, n, 8)
    if method == "rcm" and sweeps:
        assert after < 0.75 * before, (before, after)
    else:
        assert after <= before * 1.05, (before, after)
