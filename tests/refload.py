"""Test-side alias of oracle/refload.py (loader of the unmodified reference)."""
from oracle.refload import *  # noqa: F401,F403
from oracle.refload import _stub_matplotlib, available, in_scratch, load, reference_root  # noqa: F401
