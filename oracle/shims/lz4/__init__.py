"""Stand-in for python-lz4 (reference utils.py:8). Test infrastructure only."""
