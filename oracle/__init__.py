"""CPU oracle for the graph-odenet hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker / the reported CPU
baseline -- never as the thing that is shipped or measured as "ours".

Contents
--------
``odeint.py``     restatement of the third-party ``torchdiffeq`` solver the reference
                  calls (GCN/models.py:5,192).  torchdiffeq is NOT vendored, NOT pinned and
                  NOT installed here -> **parity unpinned** for the solver: it is checked
                  against analytic ODE solutions and against the reference's own modules
                  driven through it, not against torchdiffeq itself.
``gcn_ref.py``    functional (torch-CPU) restatement of the reference GCN layer/ODE-function
                  arithmetic (GCN/layers.py, GCN/models.py).  Pinned against the reference
                  modules imported from /root/reference by ``tests/golden/make_golden.py``.
``gat_ref.py``    same for GAT/layers.py.
``qc_ref.py``     same for QC/layers.py.
``graph_ops.py``  numpy restatement of adjacency normalisation, COO->CSR canonicalisation,
                  transpose and row partitioning (GCN/utils.py:205-229).  Integer work: bit-exact.
``synth.py``      seeded synthetic graph generators shared by tests and bench.
"""
