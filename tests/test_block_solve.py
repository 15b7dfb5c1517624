"""The reduced-camera linear solve on its own (LinearSolver<PoseMatrixType>::solve, 3rdparty/g2o/g2o/core/linear_solver.h:44-90; the
reference plugs LinearSolverEigen in, solvers/eigen/linear_solver_eigen.h:92-123): the two-level block-envelope Cholesky of
csrc/ba_band.cu against a dense numpy solve of the same system.

CPU: the planner (ordering, fronts, storage map, gather lists) executed by the library's host-only inspection hook.
GPU: the kernels through uco_b200_block_solve.  Tolerance: relative 1e-9 on x (f64 Cholesky of systems with condition <= 1e4; measured
<= 1e-12) — the solvers differ from Eigen's SimplicialLDLT in elimination order only."""
import numpy as np
import pytest
import ucoslam_b200


def block_system(rng, nb, edges, shift=2.0):
    """SPD block-sparse S: every (i, j) of `edges` gets a random 6x6 coupling, diagonals made dominant; returns upper-triangle blocks + dense S"""
    edges = sorted({(min(i, j), max(i, j)) for i, j in edges if i != j})
    S = np.zeros((6 * nb, 6 * nb))
    for i, j in edges:
        Bk = rng.normal(0, 1, (6, 6))
        S[6 * i:6 * i + 6, 6 * j:6 * j + 6] = Bk
        S[6 * j:6 * j + 6, 6 * i:6 * i + 6] = Bk.T
    for i in range(nb):
        A = rng.normal(0, 1, (6, 6))
        S[6 * i:6 * i + 6, 6 * i:6 * i + 6] = A @ A.T
    w = np.abs(S).sum(1)
    S += np.diag(w * (shift - 1.0) + 1.0)
    ij = [(i, i) for i in range(nb)] + edges
    ij.sort()
    blocks = np.array([S[6 * i:6 * i + 6, 6 * j:6 * j + 6].reshape(36) for i, j in ij])
    return np.array(ij, np.int32), blocks, S


def graphs(rng):
    def ring(n, w):
        return [(i, (i + d) % n) for i in range(n) for d in range(1, w + 1)]
    def chain(n, w):
        return [(i, i + d) for i in range(n) for d in range(1, w + 1) if i + d < n]
    out = {
        "ring_498_w5": (498, ring(498, 5)),                  # BASELINE config 5's keyframe graph (closed loop, 6 consecutive observers)
        "chain_240_w5": (240, chain(240, 5)),
        "chain_with_loop_closures": (300, chain(300, 4) + [(10, 290), (11, 291), (12, 289), (100, 200), (101, 201)]),
        "two_components": (130, chain(70, 3) + [(70 + i, 70 + j) for i, j in chain(60, 2)]),
        "dense_20": (20, [(i, j) for i in range(20) for j in range(i + 1, 20)]),
        "diagonal_only": (9, []),
        "single": (1, []),
        "grid_12x12": (144, [(12 * r + c, 12 * r + c + 1) for r in range(12) for c in range(11)] + [(12 * r + c, 12 * r + c + 12) for r in range(11) for c in range(12)]),
        "random_sparse": (200, [(int(a), int(b)) for a, b in rng.integers(0, 200, (500, 2))]),
        "star": (60, [(0, i) for i in range(1, 60)]),
    }
    return out


@pytest.mark.parametrize("name", list(graphs(np.random.default_rng(0)).keys()))
@pytest.mark.parametrize("force_k", [-1, 0, 3])
def test_planner_host_execution_matches_dense_solve(name, force_k):
    rng = np.random.default_rng(7)
    nb, edges = graphs(rng)[name]
    ij, blocks, S = block_system(rng, nb, edges)
    b = rng.normal(0, 1, 6 * nb)
    x, info = ucoslam_b200.probe_block_solve(nb, ij, blocks, b, force_k=force_k)
    assert x is not None
    ref = np.linalg.solve(S, b)
    assert np.abs(x - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    if force_k == 0:
        assert info[1] == 0


def test_planner_cuts_the_loop_of_config_5():
    """a closed trajectory: the cost model must cut it (a chain of 498 pivots is what the solver is there to avoid)"""
    rng = np.random.default_rng(1)
    nb, edges = graphs(rng)["ring_498_w5"]
    ij, blocks, S = block_system(rng, nb, edges)
    _, info = ucoslam_b200.probe_block_solve(nb, ij, blocks, np.zeros(6 * nb), solve=False)
    fronts, K, root_n, root_W, max_n, max_W, max_br = [int(v) for v in info[:7]]
    assert K >= 3 and fronts >= 2 * K
    assert root_n + max_n <= 200          # chain of dependent pivots: longest interior front + separators, from 498
    assert max_br <= 12 and max_W <= 7     # an arc of the loop touches 5 + 5 separator keyframes and has a band of 5 blocks


def test_not_positive_definite_is_reported():
    rng = np.random.default_rng(2)
    ij, blocks, S = block_system(rng, 30, [(i, i + 1) for i in range(29)])
    blocks = blocks.copy()
    k = int(np.nonzero((ij[:, 0] == 17) & (ij[:, 1] == 17))[0][0])
    blocks[k] = -blocks[k]
    x, info = ucoslam_b200.probe_block_solve(30, ij, blocks, np.ones(180))
    assert x is None


def test_bad_input_rejected():
    with pytest.raises(ucoslam_b200.UcoError):
        ucoslam_b200.probe_block_solve(3, np.array([[2, 1]], np.int32), np.zeros((1, 36)), np.zeros(18))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(graphs(np.random.default_rng(0)).keys()))
@pytest.mark.parametrize("force_k", [-1, 0, 3])
def test_device_solver_matches_dense_solve(name, force_k):
    ctx = ucoslam_b200.Context(0)
    rng = np.random.default_rng(11)
    nb, edges = graphs(rng)[name]
    ij, blocks, S = block_system(rng, nb, edges)
    b = rng.normal(0, 1, 6 * nb)
    x, info = ctx.block_solve(nb, ij, blocks, b, force_k=force_k)
    assert info[7] == 0
    ref = np.linalg.solve(S, b)
    assert np.abs(x - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    x2, _ = ctx.block_solve(nb, ij, blocks, b, force_k=force_k)
    assert np.array_equal(x, x2), "fixed summation order: bitwise reproducible"
    ctx.close()


@pytest.mark.gpu
def test_device_solver_reports_indefinite_and_wide_root():
    ctx = ucoslam_b200.Context(0)
    rng = np.random.default_rng(3)
    ij, blocks, S = block_system(rng, 30, [(i, i + 1) for i in range(29)])
    bad = blocks.copy()
    k = int(np.nonzero((ij[:, 0] == 17) & (ij[:, 1] == 17))[0][0])
    bad[k] = -bad[k]
    x, info = ctx.block_solve(30, ij, bad, np.ones(180))
    assert info[7] == 1 and not x.any()
    # a dense system of 40 block rows: the window (41^2 blocks) does not fit shared memory, the root runs over a global window
    nb = 40
    ij, blocks, S = block_system(rng, nb, [(i, j) for i in range(nb) for j in range(i + 1, nb)])
    b = rng.normal(0, 1, 6 * nb)
    x, info = ctx.block_solve(nb, ij, blocks, b)
    ref = np.linalg.solve(S, b)
    assert info[7] == 0 and np.abs(x - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    ctx.close()
