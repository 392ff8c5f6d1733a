"""CPU oracle for the pyMOTO hot path (numpy/scipy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing under ``pymoto_b200/`` may import this package: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and
there only as the checker / the CPU arm that is timed beside the GPU path -- never as the product path.

Every function cites the reference file:line it restates (paths relative to ``/root/reference``, pyMOTO
v2.0.1).  The arithmetic that the reference delegates to numpy / scipy (``np.add.at``, ``csr_matvec``,
``csc_matvec``, ``csr_matmat``, ``splu``, ``einsum``) is delegated to the same numpy / scipy here, so this is
the reference's algorithm on the reference's own numeric kernels without the Module/Signal runtime.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference`` (matplotlib stubbed) and stores its outputs as fixtures under ``tests/golden/``;
``tests/test_oracle.py`` checks this oracle against those fixtures (and against the live reference whenever
``/root/reference`` exists), including the reference's own known-answer tests for this path
(tests/test_assembly.py, tests/test_solvers_multigrid.py, tests/test_domain.py).
"""
from .grid import Grid  # noqa: F401
from . import assembly, filter, solvers, chain, nextrows  # noqa: F401
