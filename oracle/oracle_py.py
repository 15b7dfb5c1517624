"""TEST INFRASTRUCTURE ONLY — Python loaders for the CPU oracles (oracle/_build/liboracle.so, oracle/_ref/*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes, os, subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_sz = ctypes.c_size_t


def build_oracle():
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)


def build_ref():
    """Compile the reference's own sources (only where /root/reference exists)."""
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


_oracle = None


def load_oracle():
    global _oracle
    if _oracle is None:
        p = os.path.join(HERE, "_build", "liboracle.so")
        if not os.path.exists(p):
            build_oracle()
        _oracle = ctypes.CDLL(p)
    return _oracle


def load_ref(name):
    p = os.path.join(HERE, "_ref", name)
    return ctypes.CDLL(p) if os.path.exists(p) else None


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def hamming_knn(q, t, k, order=0):
    """oracle/knn_oracle.c on (nq,32)/(nt,32) uint8 arrays -> (idx, dist) int32 (nq,k)."""
    lib = load_oracle()
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.empty((len(q), k), np.int32)
    dist = np.empty((len(q), k), np.int32)
    lib.oracle_hamming_knn(_p(q), len(q), _sz(32), _p(t), len(t), _sz(32), k, order, _p(idx), _p(dist))
    return idx, dist


def ref_xflann_knn(q, t, k, kind=0, max_checks=-1, sorted_=0):
    """The reference's xflann (oracle/_ref/libref_xflann.so). kind 0 = linear (exact), 1 = HKMeans(32,0)."""
    lib = load_ref("libref_xflann.so")
    if lib is None:
        return None
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.empty((len(q), k), np.int32)
    dist = np.empty((len(q), k), np.int32)
    rc = lib.ref_xflann_knn(_p(q), len(q), _p(t), len(t), k, kind, max_checks, sorted_, _p(idx), _p(dist))
    if rc != 0:
        raise RuntimeError("ref_xflann_knn failed rc=%d" % rc)
    return idx, dist


def synth_descriptors(seed, nt, nq, max_flips=40):
    """uniform 256-bit train rows; queries = train rows with 0..max_flips random bit flips."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    if nt == 0:
        return t, rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    src = rng.integers(0, nt, nq)
    q = t[src].copy()
    nflip = rng.integers(0, max_flips + 1, nq)
    for i in range(nq):
        bits = rng.integers(0, 256, nflip[i])
        np.bitwise_xor.at(q[i], bits >> 3, (1 << (bits & 7)).astype(np.uint8))
    return t, q
