"""CPU oracle for the DiffUS B-mode renderer hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and the CPU-baseline / ``--impl reference`` legs of ``bench.py`` may import it, and there
only as the checker (or as the thing timed *as the CPU baseline*), never as the path that
is shipped.  The product path (``diffus_b200``) never imports this package and fails
loudly if its CUDA library is missing.

Contents
--------
``port.py``
    A CPU restatement (torch, fp32 or fp64, autograd-capable) of the reference algorithm
    for the path, each function citing the reference file:line it follows.  It contains
    BOTH the reference's literal algorithm (one dense linear solve per truncation depth,
    ``echo_dense_solve``) and the algebraically identical 2x2 prefix-product closed form
    (``echo_closed_form``) used for sizes the literal algorithm cannot reach.
``reference_loader.py``
    Imports the UNMODIFIED reference from ``/root/reference`` (present only in the build
    container) with its absent plotting/IO dependencies stubbed.  Used by
    ``make_golden.py`` and by ``tests/test_oracle_vs_reference.py`` (skipped when the
    reference tree is absent, e.g. on the GPU box).
``make_golden.py``
    Generates ``tests/golden/*.npz`` by running the real reference; the fixtures are what
    pins the oracle (and therefore the CUDA path) on machines without the reference.

Parity status: the reference has no tests or golden vectors of its own for this path
(SURVEY.md section 4), so parity is pinned BY EXECUTION of the reference in the build
container: the committed fixtures under ``tests/golden/`` are outputs of the reference's
own code (torch 2.11.0 CPU), and ``port.py`` is checked against them and against the live
reference when it is importable.
"""
