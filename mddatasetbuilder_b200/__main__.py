"""``python -m mddatasetbuilder_b200``: runs the ``datasetbuilder`` command line on the B200 path."""

import sys

from . import datasetbuilder


def main() -> int:
    datasetbuilder._commandline()
    return 0


if __name__ == "__main__":
    sys.exit(main())
