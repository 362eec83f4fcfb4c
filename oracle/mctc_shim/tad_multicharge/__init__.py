"""
TEST INFRASTRUCTURE ONLY -- stand-in for ``tad-multicharge==0.5.0``
(``/root/reference/setup.cfg:37``), absent from ``/root/reference`` and from
this image, so that the UNMODIFIED reference can run its default ``q=None``
path in the build container.  The arithmetic is the restatement in
``oracle/eeq_oracle.py`` (header there: algorithm, pinning).
"""
from __future__ import annotations

import sys
from pathlib import Path

_ORACLE = str(Path(__file__).resolve().parents[2])
if _ORACLE not in sys.path:
    sys.path.insert(0, _ORACLE)

from eeq_oracle import get_eeq_charges  # noqa: E402

__all__ = ["get_eeq_charges"]
