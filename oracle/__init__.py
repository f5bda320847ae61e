"""CPU oracle for the GNODE hot path of dkssud715/swarm-ode.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as a fallback of the CUDA path.

PARITY UNPINNED.  The reference (``/root/reference``) holds no tests, golden
vectors, checkpoints or stored outputs for this path, and all of its
arithmetic lives in two third-party libraries that are neither vendored in the
reference nor installable here (no network, no wheel):

* ``torchdiffeq`` (no version pinned by the reference; 0.2.x assumed) --
  ``odeint`` fixed-grid ``euler`` / ``midpoint`` / ``rk4`` (3/8 rule) and the
  adaptive ``dopri5``;
* ``torch_geometric`` (no version pinned; 2.5/2.6 assumed) --
  ``SAGEConv(aggr='mean')``, ``Batch.from_data_list``.

The modules below restate the published algorithms of those libraries and
anchor them on the reference's own call sites
(``scripts/train_gde.py:20-106``, ``:108-271``, ``:363-375``;
``scripts/gnode.py:160-174``).  What *does* pin the restatement:
  - the Dormand-Prince tableau is checked against SciPy's ``RK45`` constants,
  - dopri5 values are checked against ``scipy.integrate.solve_ivp`` and
    analytic solutions,
  - fixed-step solvers are checked against closed-form matrix polynomials,
  - SAGEConv is checked on hand-computed graphs,
  - GraphConverter is integer/exact and is checked on hand-built windows.
See ``tests/test_oracle_*.py``.
"""
