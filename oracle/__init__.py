"""TEST INFRASTRUCTURE ONLY -- CPU oracles for the MSDeformAttn hot path.

Nothing under ``snipper_b200/`` may import this package.  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / --impl reference).

* ``oracle.c_oracle``  -- ctypes front-end of ``msda_oracle.c`` (plain C, fwd + analytic bwd)
* ``oracle.torch_ref`` -- torch/CPU restatement of the reference's grid_sample formulation
  and of Snipper's per-frame ``MSDeformAttn`` wrapper (differentiable through autograd)

Parity is pinned against vectors produced by the real reference: see ``oracle/make_golden.py``
and ``tests/golden/``.
"""
