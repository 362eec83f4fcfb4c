"""
TEST INFRASTRUCTURE ONLY -- minimal stand-in for the un-vendored third-party
package ``tad-mctc==0.7.0`` (``/root/reference/setup.cfg:36``).

It restates, from the published upstream semantics, only
the handful of helpers the reference's hot path calls, so that the UNMODIFIED
reference sources under ``/root/reference/src`` can be imported in the build
container to (i) validate ``oracle/d4_oracle.py`` and (ii) generate the golden
vectors under ``tests/golden`` (``oracle/make_golden.py``).  It is never
imported by the product package.

parity unpinned for these third-party pieces except through the reference's
in-tree known-answer vectors (see ``tests/test_oracle_kat.py``).
"""
from . import batch, convert, data, math, ncoord, storch, typing  # noqa: F401

__version__ = "0.7.0+shim"
