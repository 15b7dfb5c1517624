"""ctypes binding of libucoslam_b200.so (the C ABI declared in include/ucoslam_b200.h).

Used by tests/, bench.py and __graft_entry__.py only: the product is the shared library plus the C++ adapters in
ucoslam-cv3_b200/host/.  There is NO CPU fallback here: if the library is missing or no CUDA device can be opened the
calls raise.
"""
import ctypes, os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libucoslam_b200.so")

UCO_KNN_HEAP, UCO_KNN_SORTED = 0, 1
_c = ctypes
_vp, _i, _sz, _u64 = _c.c_void_p, _c.c_int, _c.c_size_t, _c.c_uint64

# name -> (restype, argtypes); must list every symbol include/ucoslam_b200.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "uco_b200_create": (_vp, [_i, _i]),
    "uco_b200_destroy": (None, [_vp]),
    "uco_b200_last_error": (_c.c_char_p, [_vp]),
    "uco_b200_stream": (_vp, [_vp]),
    "uco_b200_sync": (_i, [_vp]),
    "uco_b200_launch_count": (_u64, [_vp]),
    "uco_b200_version": (_i, []),
    "uco_b200_hamming_knn": (_i, [_vp, _vp, _i, _sz, _vp, _i, _sz, _i, _i, _vp, _vp]),
    "uco_b200_hamming_knn_dev": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libucoslam_b200.so not built: run python ucoslam-cv3_b200/build.py (%s)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        _lib = lib
    return _lib


class UcoError(RuntimeError):
    pass


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)


class Context:
    """One uco_b200_ctx: a CUDA stream plus workspaces. Not thread safe (one per calling thread)."""

    def __init__(self, device=0):
        self.lib = load()
        self.h = self.lib.uco_b200_create(device, 0)
        if not self.h:
            raise UcoError("uco_b200_create failed: no usable CUDA device (there is no CPU fallback)")

    def close(self):
        if self.h:
            self.lib.uco_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise UcoError("uco_b200 error %d: %s" % (rc, self.lib.uco_b200_last_error(self.h).decode()))

    @property
    def stream(self):
        return self.lib.uco_b200_stream(self.h)

    def sync(self):
        self._chk(self.lib.uco_b200_sync(self.h))

    def launch_count(self):
        return int(self.lib.uco_b200_launch_count(self.h))

    # -- K7 ------------------------------------------------------------------------------------------------------
    def hamming_knn(self, q, t, k, order=UCO_KNN_HEAP):
        """q: (nq,32) uint8 host rows (may be strided), t: (nt,32). Returns (idx, dist) int32 (nq,k)."""
        q = np.asarray(q)
        t = np.asarray(t)
        assert q.dtype == np.uint8 and t.dtype == np.uint8 and q.ndim == 2 and t.ndim == 2
        assert q.shape[1] == 32 and t.shape[1] == 32
        assert q.strides[1] == 1 and t.strides[1] == 1
        nq, nt = q.shape[0], t.shape[0]
        idx = np.empty((nq, k), np.int32)
        dist = np.empty((nq, k), np.int32)
        qs = q.strides[0] if nq > 0 else 32
        ts = t.strides[0] if nt > 0 else 32
        self._chk(self.lib.uco_b200_hamming_knn(self.h, _p(q), nq, qs, _p(t), nt, ts, k, order, _p(idx), _p(dist)))
        return idx, dist

    def hamming_knn_dev(self, q_dev, nq, t_dev, nt, k, order, idx_dev, dist_dev):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr()); asynchronous on the context stream."""
        self._chk(self.lib.uco_b200_hamming_knn_dev(self.h, q_dev, nq, t_dev, nt, k, order, idx_dev, dist_dev))
