"""IQ recording sources (directdemod/source.py:18-324): same classes and ``read`` semantics as the
reference (complex64 samples minus 127.5+127.5j), plus ``readRaw`` which hands the interleaved
unsigned 8-bit bytes through untouched.  A ``commSignal`` built from such a ``RawIQ8`` block runs
the fused chain with the conversion inside the kernel (DDM_IN_CU8): a quarter of the PCIe and
HBM bytes, same results.  File access itself stays host-side numpy memmaps, like the reference's.
"""

from __future__ import annotations

from abc import ABCMeta, abstractmethod

import numpy as np
import scipy.io.wavfile

from . import constants


class RawIQ8:
    """A block of raw unsigned 8-bit I/Q: ``data`` is a uint8 array of shape (n, 2)."""

    def __init__(self, data):
        data = np.asarray(data)
        if data.dtype != np.uint8 or data.ndim != 2 or data.shape[1] != 2:
            raise TypeError("RawIQ8 wants a uint8 array of shape (n, 2)")
        self.data = data

    def __len__(self):
        return self.data.shape[0]

    def to_complex64(self):
        """What the reference's read() returns for these bytes (source.py:117-118)."""
        s = self.data[:, 0] + 1j * self.data[:, 1]
        return np.array(s).astype("complex64") - (127.5 + 1j * 127.5)


class source(metaclass=ABCMeta):
    @property
    @abstractmethod
    def sourceType(self):
        pass

    @property
    @abstractmethod
    def sampFreq(self):
        pass

    @property
    @abstractmethod
    def length(self):
        pass

    @abstractmethod
    def read(self, fromIndex, toIndex):
        pass


class _IQ8Source(source):
    """Shared logic of the 8-bit sources: a (n, 2) uint8 view, an offset/limit window."""

    def __init__(self, pairs, sampFreq, sourceType):
        self._pairs = pairs
        self._offset = 0
        self._sampFreq = sampFreq
        self._sourceType = sourceType
        self._actualLength = pairs.shape[0]
        self._length = pairs.shape[0]

    @property
    def sampFreq(self):
        return self._sampFreq

    @property
    def sourceType(self):
        return self._sourceType

    @property
    def length(self):
        return self._length

    def _window(self, fromIndex, toIndex):
        fromIndex += self._offset
        if toIndex is None:
            toIndex = fromIndex + 1
        else:
            toIndex += self._offset
        lo, hi = fromIndex - self._offset, toIndex - self._offset
        if lo < 0 or hi < 0 or lo >= self.length or hi > self.length:
            raise ValueError("fromIndex and toIndex have invalid values")
        return fromIndex, toIndex

    def read(self, fromIndex, toIndex=None):
        """Complex IQ samples [fromIndex, toIndex) as complex64 minus (127.5 + 127.5j)."""
        a, b = self._window(fromIndex, toIndex)
        return RawIQ8(self._pairs[a:b]).to_complex64()

    def readRaw(self, fromIndex, toIndex=None):
        """The same samples as a RawIQ8 block (no conversion, no copy of the memmap)."""
        a, b = self._window(fromIndex, toIndex)
        return RawIQ8(self._pairs[a:b])

    def limitData(self, initOffset=None, finalLimit=None):
        self._offset = initOffset if initOffset is not None else 0
        self._length = (finalLimit - self._offset) if finalLimit is not None else self._actualLength


class IQwav(_IQ8Source):
    """Two-channel unsigned 8-bit WAV as recorded by SDRSharp (source.py:53-138)."""

    def __init__(self, filename, givenSampFreq=None):
        fs, data = scipy.io.wavfile.read(filename, True)
        if data.ndim != 2 or data.shape[1] != 2 or data.dtype != np.uint8:
            raise ValueError("IQ.wav must hold two unsigned 8-bit channels")
        self.memmap = np.memmap(filename, offset=44, mode="r")
        super().__init__(data, fs if givenSampFreq is None else givenSampFreq, constants.SOURCE_IQWAV)


class IQwavAlt(IQwav):
    """The reference's memmap-based variant (source.py:237-324); same data, same results."""


class IQdat(_IQ8Source):
    """Raw interleaved unsigned 8-bit I/Q file (source.py:144-230)."""

    def __init__(self, filename, givenSampFreq=None):
        self.memmap = np.memmap(filename, mode="r")
        n = int(len(self.memmap) / 2)
        pairs = self.memmap[:2 * n].reshape(n, 2)
        super().__init__(pairs, constants.IQ_SDRSAMPRATE if givenSampFreq is None else givenSampFreq,
                         constants.SOURCE_IQDAT)
