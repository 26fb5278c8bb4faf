"""Module entry point (reference __main__.py:1-6)."""

from .datasetbuilder import _commandline

if __name__ == "__main__":
    _commandline()
