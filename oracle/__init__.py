"""CPU oracle for the x-vector extraction + back-end scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as the CPU
arm that is timed *beside* the GPU number), never as the thing shipped.

Each function restates one piece of the reference algorithm and cites the
``/root/reference`` file:line it follows.  Parity pinning: every function here is
checked against the *imported reference itself* (``tests/golden/make_golden.py`` runs the
reference's own modules in the build container and commits their outputs as fixtures
under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them).  The only
unpinned piece is the Kaldi ``compute-eer`` rule (Kaldi is not vendored in the
reference and is absent here) -- see ``scoring.eer_kaldi``.
"""
