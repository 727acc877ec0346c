"""Chunking helper (directdemod/chunker.py:21-84): chunk boundaries of a source plus the
dictionary of named variables that operators carry from chunk to chunk.  Host logic only;
the same object is also the seam along which a long stream is time-sharded across GPUs
(directdemod_b200/shard.py)."""

from . import constants


class chunker:
    def __init__(self, sigsrc, chunkSize=constants.PROC_CHUNKSIZE):
        """sigsrc: anything with a ``length``; chunkSize: samples per chunk.

        Full chunks are produced while ``start + chunkSize < length`` (strict), then the
        remainder -- so a length that is an exact multiple still ends with a full-size chunk
        and an empty source yields the single chunk [0, 0] (chunker.py:32-45)."""
        length = sigsrc.length
        size = chunkSize
        bounds = []
        start = 0
        while start + size < length:
            bounds.append([start, start + size])
            start += size
        if not bounds:
            bounds.append([0, length])
        elif bounds[-1][1] != length:
            bounds.append([bounds[-1][1], length])
        self._chunks = bounds
        self._vars = {}

    @property
    def getChunks(self):
        return self._chunks

    def set(self, name, value):
        self._vars[name] = value

    def get(self, name, init=None):
        """Value of a carried variable; with ``init`` given the variable is created on first
        use, without it a missing name raises KeyError (chunker.py:77-84)."""
        if init is None:
            return self._vars[name]
        if name not in self._vars:
            self._vars[name] = init
        return self._vars[name]
