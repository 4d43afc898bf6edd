"""CPU oracle for the EqVIO vision-update hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy fp64 restatement of the
reference algorithm (pvangoor/eqvio, files cited per function as
``path:line`` relative to the reference checkout).  It is the *checker* for
the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``eqvio_b200/`` imports it, and the product path has no CPU
fallback.

Pinning status: the reference cannot be compiled in this image (Eigen3,
OpenCV C++ and yaml-cpp are absent, no network) and its own test-suite holds
NO numeric golden vectors -- every reference test is a self-consistency
property.  The restatement is therefore pinned by
(1) a port of every one of those property tests (``tests/test_oracle_*``),
(2) agreement of two independently written evaluation orders (the dense
"follow-the-reference" update here vs. the structured Cholesky form used on
the GPU), and (3) committed golden vectors generated *by this oracle*
(``tests/golden``), which guard against regressions but do not originate from
the reference binary.  Bit-level parity with the Eigen build is unpinned.
"""
