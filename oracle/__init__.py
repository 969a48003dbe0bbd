"""CPU oracle for the IntrinsicTimescales.jl aABC / ACW hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain fp64 NumPy restatement of the
reference's algorithm for the hot path (every function cites the reference
file:line it follows).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``intrinsictimescales.jl_b200``) never imports it and has no CPU fallback.

PARITY PINNING STATUS (see DESIGN.md "Oracle"):
  * The reference is pure Julia; ``julia`` is not installed in this image and the
    GPU box runs the same image, so the reference itself cannot be executed and no
    ``oracle/_ref`` build exists.
  * The reference ships no golden vectors.  The oracle is pinned against every
    known-answer / property test the reference holds for this path
    (tests/test_oracle_reference_kats.py lists them with file:line).
  * Path-wise agreement with the SOSRA solver streams (StochasticDiffEq), the exact
    Levenberg-Marquardt iterates (NonlinearSolve), Romberg.jl beyond its two KATs and
    StatsBase.sample streams is **parity unpinned**: those dependencies are not
    vendored under /root/reference and their published algorithms are restated here.
"""

from . import philox, ou, summary, distances, acwred, acw, models, abc  # noqa: F401
