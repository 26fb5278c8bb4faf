"""Identity framing standing in for lz4.frame (reference utils.py:119,139).

The reference only round-trips private temp files through compress/decompress, so an
identity codec preserves behaviour. Test infrastructure only."""


def compress(data, compression_level=0, **kw):
    return bytes(data)


def decompress(data, **kw):
    return bytes(data)
