"""CPU oracle for the triflow implicit method-of-lines hot path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (numpy + scipy) of
the reference's numpy-compiler + SciPy-SuperLU path:

* ``oracle.numpy_compiler``  <- reference ``triflow/core/compilers.py:181-332``
* ``oracle.schemes``         <- reference ``triflow/core/schemes.py:29-300,502-559``

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker / the timed CPU
baseline.  Nothing under ``triflow_b200/`` imports it; the product path raises
when its CUDA library is missing instead of falling back to this code.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the reference
itself (``/root/reference``, unmodified sources under the import shims of
SURVEY.md Appendix A) in the build container and dumps F, J (CSC triplets) and
trajectories for the BASELINE.json configs at CPU-sized N into
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` requires this oracle to
reproduce every one of them bit-for-bit (trajectories: <= 1e-13 relative, they
go through SuperLU whose pivot order may differ between SciPy builds).
"""
