"""Minimal stand-in for the openbabel Python bindings. Test infrastructure only."""
from . import openbabel  # noqa: F401
