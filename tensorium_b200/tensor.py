"""TSingleTensor — minimal host-side twin of the reference's TTensor<Single> record
(source/ntensors.pas:226-902): host storage `Data`, device twin `devData`, `Groups` (= batch) and the
`lastOP` residency flag with pushToDevice / pullFromDevice (ntensors.pas:6153-6315).

Only what crosses the backend boundary is mirrored; the tensor utility methods of the
reference (printing, plotting, image I/O, ...) are out of scope.
"""
from __future__ import annotations

import numpy as np

cdCPU, cdCUDA = 0, 1

_cuda = None  # the global `cuda : TNNCuda<Single>` (ntensors.pas:1053-1056)


def initCUDART(deviceIndex: int = 0):
    """ntensors.pas:6191-6209: create the global backend instance."""
    global _cuda
    from .nncuda import TNNCuda
    if _cuda is None or _cuda.deviceIndex != deviceIndex:
        _cuda = TNNCuda(deviceIndex)
    return _cuda


def cuda():
    if _cuda is None:
        raise RuntimeError("initCUDART() has not been called")
    return _cuda


def split_view(t):
    """Where tensor t's pre-split bf16 planes live (TNNet.setSplitActivations): (device address of the hi plane's first
    element, pixel pitch, plane stride), pitch and stride in bf16 elements.  Normally its own `split` buffer with
    pitch = channels; a source of an elided route layer (TNNet.elideRoutes) is a column slice of the route's buffer."""
    v = getattr(t, "split_at", None)
    if v is not None:
        base, off, pitch, plane = v
        return base.devData + 2 * off, pitch, plane
    return t.split.devData, t.FShape[1], t.Size()


class TSingleTensor:
    def __init__(self, shape, groups: int | None = None, data=None, device: bool = True, host: bool = True,
                 devptr: int = 0, dtype=np.float32):
        self.FShape = tuple(int(s) for s in shape)
        self.Groups = int(groups) if groups is not None else (self.FShape[0] if len(self.FShape) > 1 else 1)
        self._n = int(np.prod(self.FShape)) if len(self.FShape) else 1
        self._dtype = np.dtype(dtype)  # float32, or int64 for TSizeIntTensor (max-pool indexes)
        if data is not None:
            self._data = np.ascontiguousarray(data, dtype=self._dtype).reshape(self.FShape).copy()
        else:
            # host=False: activation tensors that only ever live on the device get their host twin lazily
            self._data = np.zeros(self.FShape, dtype=self._dtype) if host else None
        self.devData = 0
        self._owns_dev = True
        self.lastOP = cdCPU if host or data is not None else cdCUDA
        if devptr:
            # non-owning view of device memory someone else allocated (a slice of the gradient bucket)
            self.devData, self._owns_dev = int(devptr), False
        elif device and self._n:
            self.devData = cuda().createDeviceBuffer(self._n * self._dtype.itemsize)  # ntensors.pas:6103-6109

    @property
    def Data(self) -> np.ndarray:
        if self._data is None:
            self._data = np.zeros(self.FShape, dtype=self._dtype)
        return self._data

    # shape helpers (ntensors.pas: Size/Volume/Area/c/h/w)
    def Size(self) -> int:
        return self._n

    size = Size

    def volume(self) -> int:
        return self.Size() // self.Groups if self.Groups else self.Size()

    def area(self) -> int:
        return self.FShape[-1] * self.FShape[-2] if len(self.FShape) >= 2 else self.Size()

    def c(self):
        return self.FShape[-3]

    def h(self):
        return self.FShape[-2]

    def w(self):
        return self.FShape[-1]

    # residency (ntensors.pas:6153-6229)
    def wasGPU(self) -> bool:
        return self.lastOP == cdCUDA

    def setCUDA(self):
        self.lastOP = cdCUDA

    def setCPU(self):
        self.lastOP = cdCPU

    def pushToDevice(self):  # ntensors.pas:6231-6262
        cuda().writeBuffer(self.devData, self.Data.nbytes, self.Data)
        self.lastOP = cdCUDA

    def pullFromDevice(self):  # ntensors.pas:6264-6315
        cuda().readBuffer(self.devData, self.Data.nbytes, self.Data)
        self.lastOP = cdCPU
        return self.Data

    def Free(self):  # ntensors.pas:6126-6129
        if self.devData and self._owns_dev:
            cuda().freeDeviceBuffer(self.devData)
        self.devData = 0

    def rebind(self, devptr: int):
        """Move the device twin into memory owned by someone else (keeps the current device contents)."""
        if self.devData and devptr != self.devData:
            cuda().copy(self._n, self.devData, 0, 1, devptr, 0, 1)
            cuda().finish()
        self.Free()
        self.devData, self._owns_dev = int(devptr), False

    def __del__(self):
        try:
            self.Free()
        except Exception:
            pass
