"""CPU oracle for the polars_ols least-squares hot path.

TEST INFRASTRUCTURE ONLY: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  The product
(``polars_ols_b200``) never does.

``oracle/ols_oracle.c`` restates ``/root/reference/src/least_squares.rs`` (solvers) in plain C;
``oracle/semantics.py`` restates the marshalling / null-policy / pre- and post-processing semantics
of ``src/expressions.rs`` and ``polars_ols/least_squares.py`` with numpy on top of it.

Parity pin: the reference holds no stored fixtures; it is pinned here against
  * the README's known-answer frame (``tests/golden/readme_frame.json``, transcribed from
    ``README.md:50-138`` by ``tests/golden/make_golden.py``),
  * the Rust unit tests ``src/lib.rs:47-171`` re-derived in ``tests/test_oracle.py``,
  * numpy ``lstsq``/``solve`` and scikit-learn ``Ridge``/``ElasticNet`` — the third-party oracles
    ``tests/test_ols.py`` itself uses, at its tolerances or tighter.
The reference binary itself cannot be built here (no Rust toolchain), so there is no ``oracle/_ref``.
"""
from .loader import lib, build_oracle  # noqa: F401
