"""CPU oracle for the ESDG right-hand side + LSRK45 hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy (FP64) restatement of the
reference Julia algorithm (yiminllin/ESDG-CNS); each function cites the
reference file:line it follows.  It is the *checker*:

  * only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
    ``--impl reference`` legs of ``bench.py`` may import it;
  * the product (``esdg_cns_b200``) never imports it and has no CPU fallback.

PARITY UNPINNED: the reference is 100 % Julia, there is no Julia toolchain in
this image, and the reference ships no golden vectors for the rhs path.  The
oracle is pinned instead by (i) a port of every property of
``examples/EntropyStableEuler.jl/test/runtests.jl``, (ii) the discrete
entropy-conservation residual at round-off, (iii) free-stream preservation,
(iv) vortex convergence, and (v) an independent scalar C restatement
(``oracle/c/esdg_oracle.c``) that must agree with the numpy one to round-off.
"""
