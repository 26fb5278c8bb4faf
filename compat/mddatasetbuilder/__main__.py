"""Module entry point (reference ``__main__.py``)."""

from mddatasetbuilder_b200.datasetbuilder import _commandline

if __name__ == "__main__":
    _commandline()
