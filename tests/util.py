"""Shared helpers for the parity tests: device buffers through the C ABI and the error metrics of SURVEY.md §8(d)."""
import numpy as np


class Dev:
    """A device buffer created/filled/read through the C ABI only (createDeviceBuffer/writeBuffer/readBuffer)."""

    def __init__(self, cu, host=None, n=None, pad_front=0):
        self.cu = cu
        if host is not None:
            host = np.ascontiguousarray(host, dtype=np.float32)
            n = host.size
        self.n = int(n)
        self.pad = int(pad_front)
        self._base = cu.createDeviceBuffer((self.n + self.pad) * 4 + 16)
        self.ptr = self._base
        if host is not None:
            self.write(host)
        else:
            cu.fill(self.n + self.pad, self._base, 0, 0.0, 1)

    def write(self, host):
        host = np.ascontiguousarray(host, dtype=np.float32).reshape(-1)
        assert host.size == self.n
        buf = np.zeros(self.n + self.pad, dtype=np.float32)
        buf[self.pad:] = host
        self.cu.writeBuffer(self._base, buf.nbytes, buf)

    def read(self, shape=None):
        buf = np.empty(self.n + self.pad, dtype=np.float32)
        self.cu.readBuffer(self._base, buf.nbytes, buf)
        out = buf[self.pad:].copy()
        return out.reshape(shape) if shape is not None else out

    def free(self):
        if self._base:
            self.cu.freeDeviceBuffer(self._base)
            self._base = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def rel_max(g, r):
    """max|g-r| / max|r|   (tensor-level metric)."""
    g = np.asarray(g, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    den = np.abs(r).max()
    return float(np.abs(g - r).max() / (den if den > 0 else 1.0))


def rel_elem(g, r, floor=1e-3):
    """max over elements of |g-r| / max(|r|, floor*rms(r))   (elementwise metric).

    floor=1e-3 is used for single-rounding ops; contractions use floor=1.0 (error of a dot product is
    absolute in nature: K*eps*|a||b|, so elements that cancel to ~0 have no meaningful relative error —
    the fp32 CPU reference itself fails a small-floor metric against fp64 truth)."""
    g = np.asarray(g, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    rms = np.sqrt(np.mean(r * r)) if r.size else 0.0
    den = np.maximum(np.abs(r), floor * rms if rms > 0 else 1.0)
    return float((np.abs(g - r) / den).max()) if r.size else 0.0


TOL_3XTF32 = 1e-4  # BASELINE.json north_star
TOL_TF32 = 5e-3
