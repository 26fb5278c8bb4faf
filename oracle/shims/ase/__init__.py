"""Minimal stand-in for the parts of ASE the reference uses. Test infrastructure only."""
from .atoms import Atom, Atoms  # noqa: F401
