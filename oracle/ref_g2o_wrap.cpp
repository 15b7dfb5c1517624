// ref_g2o_wrap.cpp — TEST INFRASTRUCTURE ONLY.  The reference's bundle adjustment on flat arrays: the vendored g2o
// (3rdparty/g2o/g2o/{core,stuff}) and the reference's OWN vertex / edge classes (src/optimization/typesg2o.h, compiled
// unchanged through oracle/shim) are built from /root/reference; this file only restates the graph assembly and the
// two-stage schedule of GlobalOptimizerG2O (src/optimization/globaloptimizer_g2o.cpp:77-401 setParams, :418-463 optimize,
// :466-538 getResults) and of PnPSolver::solvePnp (src/optimization/pnpsolver.cpp:116-408) without the Map/Frame containers.
#include "optimization/typesg2o.h"
#include "g2o/core/block_solver.h"
#include "g2o/core/optimization_algorithm_levenberg.h"
#include "g2o/core/sparse_optimizer.h"
#include "g2o/solvers/eigen/linear_solver_eigen.h"
#include <cmath>
#include <cstdint>
#include <memory>
#include <vector>

using namespace ucoslam;

static g2o::SE3Quat toSE3Quat(const float* m) {  // globaloptimizer_g2o.cpp:80-91, m = row-major 4x4 CV_32F pose_f2g
    Eigen::Matrix<double, 3, 3> R;
    R << m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10];
    Eigen::Matrix<double, 3, 1> t(m[3], m[7], m[11]);
    return g2o::SE3Quat(R, t);
}

// The reference defines this edge inside globaloptimizer_g2o.cpp (:37-66), not in a header, so it is restated here: a binary edge between
// two marker vertices (reference marker, other marker) whose 4 residuals say "both markers lie in one plane with the same normal":
// with M = inverse(ref) * other as 4x4 matrices, 10 * (M(0,2), M(1,2), 1 - M(2,2), M(2,3)).  No linearizeOplus override: g2o
// differentiates it numerically (base_binary_edge.hpp:165-233, delta = 1e-9f).
class PlanarMarkerEdge : public g2o::BaseBinaryEdge<4, Eigen::Matrix<double, 8, 1>, VertexSE3Expmap, VertexSE3Expmap> {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    bool read(std::istream&) { return false; }
    bool write(std::ostream&) const { return false; }
    void computeError() {
        const auto* a = static_cast<const VertexSE3Expmap*>(_vertices[0]);
        const auto* b = static_cast<const VertexSE3Expmap*>(_vertices[1]);
        const Eigen::Matrix<double, 4, 4> M = a->estimate().to_homogeneous_matrix().inverse() * b->estimate().to_homogeneous_matrix();
        _error.resize(4);
        _error(0) = 10. * M(0, 2);
        _error(1) = 10. * M(1, 2);
        _error(2) = 10. * (1 - M(2, 2));
        _error(3) = 10. * M(2, 3);
    }
};

extern "C" {

// obs_ur[i] is used when obs_stereo[i] != 0.  out_pose7: qx qy qz qw tx ty tz (f64); out_pose44: what getResults stores
// (CV_32F 4x4); out_chi2 / out_level / out_depth_pos: per observation state after optimize(); out_bad: the
// getBadAssociations() predicate (:506-521).  trace (optional, 64 doubles): chi2 after each outer iteration.
// Markers (globaloptimizer_g2o.cpp:157-170, 304-350): one free VertexSE3Expmap per map marker (pose_g2m) and one MarkerEdge (the
// reference's own class, typesg2o.h:108-167: 8 residuals, numeric Jacobian with delta 1e-4) per (marker, keyframe) observation with
// information = I8 * mobs_weight (frame_MarkerWeight, :276-297, computed by the caller); no robust kernel; they stay at level 0 in both
// stages (:447-451).  The InPlaneMarkers extension (:360-401) is not covered.
}  // extern "C" (the shared body has C++ linkage)
static int ba_optimize_impl(int n_poses, const float* poses44, const uint8_t* fixed, int n_points, const float* points3, int n_obs,
                    const int32_t* obs_pose, const int32_t* obs_point, const float* obs_uv, const float* obs_ur,
                    const uint8_t* obs_stereo, const float* obs_inv_sigma2, float fx, float fy, float cx, float cy, float bf,
                    int n_iters, double* out_pose7, float* out_pose44, double* out_point3, double* out_chi2,
                    uint8_t* out_level, uint8_t* out_bad, int* iters_done, double* trace,
                    int n_markers, const float* marker_pose44, const float* marker_size, int n_mobs, const int32_t* mobs_marker,
                    const int32_t* mobs_pose, const float* mobs_corners, const float* mobs_weight, double* out_marker_pose7,
                    float* out_marker_pose44, double* out_mobs_chi2,
                    const float* pose_cam /* n_poses x 5 (fx fy cx cy bf of each keyframe's ImageParams, :233-236, :262-266) or NULL */,
                    int n_plane = 0, int plane_ref = -1, const float* plane_ref_pose44 = nullptr, const int32_t* plane_other = nullptr, double plane_weight = 0);
extern "C" {


int ref_ba_optimize(int n_poses, const float* poses44, const uint8_t* fixed, int n_points, const float* points3, int n_obs,
                    const int32_t* obs_pose, const int32_t* obs_point, const float* obs_uv, const float* obs_ur,
                    const uint8_t* obs_stereo, const float* obs_inv_sigma2, float fx, float fy, float cx, float cy, float bf,
                    int n_iters, double* out_pose7, float* out_pose44, double* out_point3, double* out_chi2,
                    uint8_t* out_level, uint8_t* out_bad, int* iters_done, double* trace /* optional: per outer iteration {chi2, LM trials}, 2 x 64 */) {
    return ba_optimize_impl(n_poses, poses44, fixed, n_points, points3, n_obs, obs_pose, obs_point, obs_uv, obs_ur, obs_stereo, obs_inv_sigma2,
                            fx, fy, cx, cy, bf, n_iters, out_pose7, out_pose44, out_point3, out_chi2, out_level, out_bad, iters_done, trace,
                            0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int ref_ba_optimize_markers(int n_poses, const float* poses44, const uint8_t* fixed, int n_points, const float* points3, int n_obs,
                    const int32_t* obs_pose, const int32_t* obs_point, const float* obs_uv, const float* obs_ur,
                    const uint8_t* obs_stereo, const float* obs_inv_sigma2, float fx, float fy, float cx, float cy, float bf,
                    int n_iters, double* out_pose7, float* out_pose44, double* out_point3, double* out_chi2,
                    uint8_t* out_level, uint8_t* out_bad, int* iters_done, double* trace,
                    int n_markers, const float* marker_pose44, const float* marker_size, int n_mobs, const int32_t* mobs_marker,
                    const int32_t* mobs_pose, const float* mobs_corners, const float* mobs_weight, double* out_marker_pose7,
                    float* out_marker_pose44, double* out_mobs_chi2) {
    return ba_optimize_impl(n_poses, poses44, fixed, n_points, points3, n_obs, obs_pose, obs_point, obs_uv, obs_ur, obs_stereo, obs_inv_sigma2,
                            fx, fy, cx, cy, bf, n_iters, out_pose7, out_pose44, out_point3, out_chi2, out_level, out_bad, iters_done, trace,
                            n_markers, marker_pose44, marker_size, n_mobs, mobs_marker, mobs_pose, mobs_corners, mobs_weight, out_marker_pose7,
                            out_marker_pose44, out_mobs_chi2, nullptr);
}
// keyframes taken with different cameras in one window: every edge carries the ImageParams of ITS keyframe (:233-236, :262-266, :335-338)
int ref_ba_optimize_cams(int n_poses, const float* poses44, const uint8_t* fixed, int n_points, const float* points3, int n_obs,
                    const int32_t* obs_pose, const int32_t* obs_point, const float* obs_uv, const float* obs_ur,
                    const uint8_t* obs_stereo, const float* obs_inv_sigma2, float fx, float fy, float cx, float cy, float bf,
                    int n_iters, double* out_pose7, float* out_pose44, double* out_point3, double* out_chi2,
                    uint8_t* out_level, uint8_t* out_bad, int* iters_done, double* trace,
                    int n_markers, const float* marker_pose44, const float* marker_size, int n_mobs, const int32_t* mobs_marker,
                    const int32_t* mobs_pose, const float* mobs_corners, const float* mobs_weight, double* out_marker_pose7,
                    float* out_marker_pose44, double* out_mobs_chi2, const float* pose_cam) {
    return ba_optimize_impl(n_poses, poses44, fixed, n_points, points3, n_obs, obs_pose, obs_point, obs_uv, obs_ur, obs_stereo, obs_inv_sigma2,
                            fx, fy, cx, cy, bf, n_iters, out_pose7, out_pose44, out_point3, out_chi2, out_level, out_bad, iters_done, trace,
                            n_markers, marker_pose44, marker_size, n_mobs, mobs_marker, mobs_pose, mobs_corners, mobs_weight, out_marker_pose7,
                            out_marker_pose44, out_mobs_chi2, pose_cam);
}
// + the InPlaneMarkers edges: plane_ref = marker index of the reference marker, or -1 with its pose in plane_ref_pose44 (fixed vertex)
int ref_ba_optimize_planar(int n_poses, const float* poses44, const uint8_t* fixed, int n_points, const float* points3, int n_obs,
                    const int32_t* obs_pose, const int32_t* obs_point, const float* obs_uv, const float* obs_ur,
                    const uint8_t* obs_stereo, const float* obs_inv_sigma2, float fx, float fy, float cx, float cy, float bf,
                    int n_iters, double* out_pose7, float* out_pose44, double* out_point3, double* out_chi2,
                    uint8_t* out_level, uint8_t* out_bad, int* iters_done, double* trace,
                    int n_markers, const float* marker_pose44, const float* marker_size, int n_mobs, const int32_t* mobs_marker,
                    const int32_t* mobs_pose, const float* mobs_corners, const float* mobs_weight, double* out_marker_pose7,
                    float* out_marker_pose44, double* out_mobs_chi2, const float* pose_cam, int n_plane, int plane_ref, const float* plane_ref_pose44,
                    const int32_t* plane_other, double plane_weight) {
    return ba_optimize_impl(n_poses, poses44, fixed, n_points, points3, n_obs, obs_pose, obs_point, obs_uv, obs_ur, obs_stereo, obs_inv_sigma2,
                            fx, fy, cx, cy, bf, n_iters, out_pose7, out_pose44, out_point3, out_chi2, out_level, out_bad, iters_done, trace,
                            n_markers, marker_pose44, marker_size, n_mobs, mobs_marker, mobs_pose, mobs_corners, mobs_weight, out_marker_pose7,
                            out_marker_pose44, out_mobs_chi2, pose_cam, n_plane, plane_ref, plane_ref_pose44, plane_other, plane_weight);
}
}  // extern "C"

static int ba_optimize_impl(int n_poses, const float* poses44, const uint8_t* fixed, int n_points, const float* points3, int n_obs,
                    const int32_t* obs_pose, const int32_t* obs_point, const float* obs_uv, const float* obs_ur,
                    const uint8_t* obs_stereo, const float* obs_inv_sigma2, float fx, float fy, float cx, float cy, float bf,
                    int n_iters, double* out_pose7, float* out_pose44, double* out_point3, double* out_chi2,
                    uint8_t* out_level, uint8_t* out_bad, int* iters_done, double* trace,
                    int n_markers, const float* marker_pose44, const float* marker_size, int n_mobs, const int32_t* mobs_marker,
                    const int32_t* mobs_pose, const float* mobs_corners, const float* mobs_weight, double* out_marker_pose7,
                    float* out_marker_pose44, double* out_mobs_chi2, const float* pose_cam, int n_plane, int plane_ref, const float* plane_ref_pose44,
                    const int32_t* plane_other, double plane_weight) {
    const float Chi2D = 5.99f, Chi3D = 7.815f;
    const float thHuber2D = sqrt(Chi2D), thHuber3D = sqrt(Chi3D);
    auto Optimizer = std::make_shared<g2o::SparseOptimizer>();
    std::unique_ptr<g2o::BlockSolver_6_3::LinearSolverType> linearSolver =
        g2o::make_unique<g2o::LinearSolverEigen<g2o::BlockSolver_6_3::PoseMatrixType>>();
    auto* solver = new g2o::OptimizationAlgorithmLevenberg(g2o::make_unique<g2o::BlockSolver_6_3>(std::move(linearSolver)));
    Optimizer->setAlgorithm(solver);
    for (int i = 0; i < n_poses; i++) {
        auto* v = new VertexSE3Expmap();
        v->setEstimate(toSE3Quat(poses44 + 16 * i));
        v->setId(i);
        if (fixed[i]) v->setFixed(true);
        Optimizer->addVertex(v);
    }
    std::vector<g2o::OptimizableGraph::Edge*> edges(n_obs);
    int next_obs = 0;
    for (int p = 0; p < n_points; p++) {
        auto* v = new VertexSBAPointXYZ();
        Eigen::Matrix<double, 3, 1> x;
        x << points3[3 * p], points3[3 * p + 1], points3[3 * p + 2];
        v->setEstimate(x);
        v->setId(n_poses + p);
        v->setMarginalized(true);
        Optimizer->addVertex(v);
    }
    for (int i = 0; i < n_obs; i++) {
        (void)next_obs;
        auto* vp = dynamic_cast<g2o::OptimizableGraph::Vertex*>(Optimizer->vertex(n_poses + obs_point[i]));
        auto* vf = dynamic_cast<g2o::OptimizableGraph::Vertex*>(Optimizer->vertex(obs_pose[i]));
        if (!obs_stereo[i]) {
            Eigen::Matrix<double, 2, 1> obs;
            obs << obs_uv[2 * i], obs_uv[2 * i + 1];
            auto* e = new EdgeSE3ProjectXYZ();
            const float* pc = pose_cam ? pose_cam + 5 * obs_pose[i] : nullptr;
            e->fx = pc ? pc[0] : fx; e->fy = pc ? pc[1] : fy; e->cx = pc ? pc[2] : cx; e->cy = pc ? pc[3] : cy;
            e->setVertex(0, vp);
            e->setVertex(1, vf);
            e->setMeasurement(obs);
            e->setInformation(Eigen::Matrix2d::Identity() * obs_inv_sigma2[i]);
            auto* rk = new g2o::RobustKernelHuber();
            rk->setDelta(thHuber2D);
            e->setRobustKernel(rk);
            Optimizer->addEdge(e);
            edges[i] = e;
        } else {
            Eigen::Matrix<double, 3, 1> obs;
            obs << obs_uv[2 * i], obs_uv[2 * i + 1], obs_ur[i];
            auto* e = new EdgeStereoSE3ProjectXYZ();
            e->setVertex(0, vp);
            e->setVertex(1, vf);
            e->setMeasurement(obs);
            e->setInformation(Eigen::Matrix3d::Identity() * double(obs_inv_sigma2[i]));
            auto* rk = new g2o::RobustKernelHuber();
            rk->setDelta(thHuber3D);
            e->setRobustKernel(rk);
            const float* pc = pose_cam ? pose_cam + 5 * obs_pose[i] : nullptr;
            e->fx = pc ? pc[0] : fx; e->fy = pc ? pc[1] : fy; e->cx = pc ? pc[2] : cx; e->cy = pc ? pc[3] : cy; e->bf = pc ? pc[4] : bf;
            Optimizer->addEdge(e);
            edges[i] = e;
        }
    }
    std::vector<MarkerEdge*> marker_edges;
    for (int m = 0; m < n_markers; m++) {   // :304-313
        auto* v = new VertexSE3Expmap();
        v->setEstimate(toSE3Quat(marker_pose44 + 16 * m));
        v->setId(n_poses + n_points + m);
        Optimizer->addVertex(v);
    }
    for (int k = 0; k < n_mobs; k++) {      // :317-347
        auto* e = new MarkerEdge(marker_size[mobs_marker[k]], mobs_marker[k], mobs_pose[k]);
        Eigen::Matrix<double, 8, 1> obs;
        for (int i = 0; i < 8; i++) obs(i) = mobs_corners[8 * k + i];
        e->setMeasurement(obs);
        e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(Optimizer->vertex(n_poses + n_points + mobs_marker[k])));
        e->setVertex(1, dynamic_cast<g2o::OptimizableGraph::Vertex*>(Optimizer->vertex(mobs_pose[k])));
        const float* pc = pose_cam ? pose_cam + 5 * mobs_pose[k] : nullptr;
        e->fx = pc ? pc[0] : fx; e->fy = pc ? pc[1] : fy; e->cx = pc ? pc[2] : cx; e->cy = pc ? pc[3] : cy;
        e->setInformation(Eigen::Matrix<double, 8, 8>::Identity() * double(mobs_weight[k]));
        Optimizer->addEdge(e);
        marker_edges.push_back(e);
    }
    // InPlaneMarkers (:356-401): the reference marker (a free vertex of this window, or a fixed extra vertex when the window does not hold it)
    // is tied to every other marker by one planar edge of information plane_weight * I4 (:384-392); the edges keep level 0 and no kernel
    if (n_plane > 0) {
        g2o::OptimizableGraph::Vertex* vref = nullptr;
        if (plane_ref >= 0) vref = dynamic_cast<g2o::OptimizableGraph::Vertex*>(Optimizer->vertex(n_poses + n_points + plane_ref));
        else {
            auto* v = new VertexSE3Expmap();
            v->setEstimate(toSE3Quat(plane_ref_pose44));
            v->setId(std::numeric_limits<int>::max());
            v->setFixed(true);
            Optimizer->addVertex(v);
            vref = v;
        }
        for (int k = 0; k < n_plane; k++) {
            auto* e = new PlanarMarkerEdge();
            e->setVertex(0, vref);
            e->setVertex(1, dynamic_cast<g2o::OptimizableGraph::Vertex*>(Optimizer->vertex(n_poses + n_points + plane_other[k])));
            e->setInformation(plane_weight * Eigen::Matrix<double, 4, 4>::Identity());
            Optimizer->addEdge(e);
        }
    }
    // optimize(), :418-463
    Optimizer->initializeOptimization();
    Optimizer->setVerbose(false);
    if (trace) Optimizer->setComputeBatchStatistics(true);  // g2o's own per-iteration record (extra error evaluations only)
    int ntrace = 0;
    auto dump = [&](int its) {
        for (int i = 0; i < its && trace && ntrace < 64; i++, ntrace++) {
            trace[2 * ntrace] = Optimizer->batchStatistics()[i].chi2;
            trace[2 * ntrace + 1] = Optimizer->batchStatistics()[i].levenbergIterations;
        }
    };
    int it1 = Optimizer->optimize(n_iters, 1);
    dump(it1);
    for (int i = 0; i < n_obs; i++) {
        if (obs_stereo[i]) {
            auto* e = (EdgeStereoSE3ProjectXYZ*)edges[i];
            if (e->chi2() > Chi3D || !e->isDepthPositive()) e->setLevel(1);
            e->setRobustKernel(0);
        } else {
            auto* e = (EdgeSE3ProjectXYZ*)edges[i];
            if (e->chi2() > Chi2D || !e->isDepthPositive()) e->setLevel(1);
            e->setRobustKernel(0);
        }
    }
    for (auto* me : marker_edges) {   // :447-451
        if (me->chi2() > 15.507) me->setLevel(0);
        me->setRobustKernel(0);
    }
    Optimizer->initializeOptimization();
    int it2 = Optimizer->optimize(n_iters * 2, 1);
    dump(it2);
    if (iters_done) { iters_done[0] = it1; iters_done[1] = it2; }
    // getResults(), :466-538
    for (int i = 0; i < n_poses; i++) {
        auto* v = static_cast<VertexSE3Expmap*>(Optimizer->vertex(i));
        g2o::SE3Quat q = v->estimate();
        out_pose7[7 * i + 0] = q.rotation().x(); out_pose7[7 * i + 1] = q.rotation().y(); out_pose7[7 * i + 2] = q.rotation().z();
        out_pose7[7 * i + 3] = q.rotation().w();
        out_pose7[7 * i + 4] = q.translation()[0]; out_pose7[7 * i + 5] = q.translation()[1]; out_pose7[7 * i + 6] = q.translation()[2];
        if (fixed[i]) { for (int k = 0; k < 16; k++) out_pose44[16 * i + k] = poses44[16 * i + k]; continue; }   // :483 untouched
        Eigen::Matrix<double, 4, 4> M = q.to_homogeneous_matrix();
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out_pose44[16 * i + 4 * r + c] = (float)M(r, c);
    }
    for (int p = 0; p < n_points; p++) {
        auto* v = static_cast<VertexSBAPointXYZ*>(Optimizer->vertex(n_poses + p));
        for (int k = 0; k < 3; k++) out_point3[3 * p + k] = v->estimate()(k);
    }
    for (int m = 0; m < n_markers; m++) {   // :526-527
        g2o::SE3Quat q = static_cast<VertexSE3Expmap*>(Optimizer->vertex(n_poses + n_points + m))->estimate();
        out_marker_pose7[7 * m + 0] = q.rotation().x(); out_marker_pose7[7 * m + 1] = q.rotation().y(); out_marker_pose7[7 * m + 2] = q.rotation().z();
        out_marker_pose7[7 * m + 3] = q.rotation().w();
        for (int k = 0; k < 3; k++) out_marker_pose7[7 * m + 4 + k] = q.translation()[k];
        Eigen::Matrix<double, 4, 4> M = q.to_homogeneous_matrix();
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out_marker_pose44[16 * m + 4 * r + c] = (float)M(r, c);
    }
    for (int k = 0; k < n_mobs && out_mobs_chi2; k++) out_mobs_chi2[k] = marker_edges[k]->chi2();
    for (int i = 0; i < n_obs; i++) {
        bool bad = false;
        double c2;
        if (obs_stereo[i]) {
            auto* e = (EdgeStereoSE3ProjectXYZ*)edges[i];
            c2 = e->chi2();
            if (c2 > Chi3D || !e->isDepthPositive()) bad = true;
            out_level[i] = (uint8_t)e->level();
        } else {
            auto* e = (EdgeSE3ProjectXYZ*)edges[i];
            c2 = e->chi2();
            if (c2 > Chi2D) bad = true;
            out_level[i] = (uint8_t)e->level();
        }
        out_chi2[i] = c2;
        if (!bad) {  // :515-518  pincam = pose_f2g (f32, already updated) * point (f32)
            const float* m = out_pose44 + 16 * obs_pose[i];
            float px = (float)out_point3[3 * obs_point[i]], py = (float)out_point3[3 * obs_point[i] + 1],
                  pz = (float)out_point3[3 * obs_point[i] + 2];
            float z = m[8] * px + m[9] * py + m[10] * pz + m[11];
            if (z < 0) bad = true;
        }
        out_bad[i] = bad;
    }
    return 0;
}

extern "C" {

// PnPSolver::solvePnp (src/optimization/pnpsolver.cpp:116-408) on flat arrays: one free camera vertex, one unary edge per
// (keypoint, map point) match, optional fixed marker vertices with MarkerEdgeOnlyProject edges, 4 rounds x 10 LM iterations
// with inlier re-classification.  stable[i] = MapPoint::isStable(); obs_ur[i] = kpt.pt.x - mbf/depth (:226) where stereo[i].
// Returns the number of inliers (the reference's return value); bad[i] != 0 <-> map_matches[i].imgIdx = -1.
int ref_pose_only(const float* pose44, int n, const float* points3, const float* obs_uv, const float* obs_ur,
                  const uint8_t* obs_stereo, const float* obs_inv_sigma2, const uint8_t* stable, float fx, float fy, float cx,
                  float cy, float bf, int n_markers, const float* marker_pose44, const float* marker_size,
                  const float* marker_corners /* n_markers x 8 */, float* out_pose44, double* out_pose7, uint8_t* bad,
                  int* iters_done /* 4 */) {
    if (n == 0 && n_markers == 0) return 0;  // :144
    g2o::SparseOptimizer optimizer;
    std::unique_ptr<g2o::BlockSolver_6_3::LinearSolverType> linearSolver =
        g2o::make_unique<g2o::LinearSolverEigen<g2o::BlockSolver_6_3::PoseMatrixType>>();
    auto* solver = new g2o::OptimizationAlgorithmLevenberg(g2o::make_unique<g2o::BlockSolver_6_3>(std::move(linearSolver)));
    optimizer.setAlgorithm(solver);
    auto* cam = new VertexSE3Expmap();
    cam->setEstimate(toSE3Quat(pose44));
    cam->setId(0);
    cam->setFixed(false);
    optimizer.addVertex(cam);
    const float Chi2D = 5.99, Chi3D = 7.815, Chi8D = 15.507;
    const float thHuber2D = sqrt(5.99), thHuber3D = sqrt(7.815), thHuber8D = sqrt(15.507);
    struct edgeinfo { float MaxChi = 0; void* ptr; };
    std::vector<edgeinfo> edgesInfo(n);
    std::vector<bool> vBad(n, false);
    double KpWeightSum = 0;
    for (int i = 0; i < n; i++) {
        float edge_weight = 1;
        if (!stable[i]) edge_weight = 0.5;
        const float invSigma2 = obs_inv_sigma2[i];
        if (!obs_stereo[i]) {
            Eigen::Matrix<double, 2, 1> obs;
            obs << obs_uv[2 * i], obs_uv[2 * i + 1];
            auto* e = new EdgeSE3ProjectXYZOnlyPose(points3[3 * i], points3[3 * i + 1], points3[3 * i + 2], fx, fy, cx, cy);
            e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(cam));
            e->setMeasurement(obs);
            e->setInformation(Eigen::Matrix2d::Identity() * invSigma2);
            auto* rk = new WeightedHubberRobustKernel;
            rk->set(thHuber2D, edge_weight);
            e->setRobustKernel(rk);
            optimizer.addEdge(e);
            edgesInfo[i].ptr = (void*)e;
            edgesInfo[i].MaxChi = Chi2D;
        } else {
            Eigen::Matrix<double, 3, 1> obs;
            obs << obs_uv[2 * i], obs_uv[2 * i + 1], obs_ur[i];
            auto* e = new EdgeStereoSE3ProjectXYZOnlyPose();
            e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(cam));
            e->setMeasurement(obs);
            Eigen::Matrix3d Info = Eigen::Matrix3d::Identity() * invSigma2;
            e->setInformation(Info);
            edge_weight *= 2;
            auto* rk = new WeightedHubberRobustKernel;
            rk->set(thHuber3D, edge_weight);
            e->setRobustKernel(rk);
            e->fx = fx; e->fy = fy; e->cx = cx; e->cy = cy; e->bf = bf;
            e->Xw[0] = points3[3 * i]; e->Xw[1] = points3[3 * i + 1]; e->Xw[2] = points3[3 * i + 2];
            optimizer.addEdge(e);
            edgesInfo[i].ptr = (void*)e;
            edgesInfo[i].MaxChi = Chi3D;
        }
        KpWeightSum += edge_weight;
    }
    std::vector<MarkerEdgeOnlyProject*> marker_edges;
    float w_markers = 0.3;
    int totalNEdges = n + n_markers;
    double weight_marker = ((w_markers * totalNEdges) / (1. - w_markers)) / float(KpWeightSum);
    uint32_t vid = 1;
    for (int m = 0; m < n_markers; m++) {
        auto* vm = new VertexSE3Expmap();
        vm->setEstimate(toSE3Quat(marker_pose44 + 16 * m));
        vm->setFixed(true);
        vm->setId(vid++);
        optimizer.addVertex(vm);
        auto* e = new MarkerEdgeOnlyProject(marker_size[m]);
        Eigen::Matrix<double, 8, 1> obs;
        for (int i = 0; i < 8; i++) obs(i) = marker_corners[8 * m + i];
        e->setMeasurement(obs);
        e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(vm));
        e->setVertex(1, dynamic_cast<g2o::OptimizableGraph::Vertex*>(cam));
        e->fx = fx; e->fy = fy; e->cx = cx; e->cy = cy;
        e->setInformation(Eigen::Matrix<double, 8, 8>::Identity());
        auto* rk = new WeightedHubberRobustKernel;
        e->setRobustKernel(rk);
        rk->set(thHuber8D, weight_marker);
        optimizer.addEdge(e);
        marker_edges.push_back(e);
    }
    for (int it = 0; it < 4; it++) {
        if (iters_done) iters_done[it] = 0;
    }
    for (int it = 0; it < 4; it++) {
        cam->setEstimate(toSE3Quat(pose44));
        optimizer.initializeOptimization(0);
        optimizer.setVerbose(false);
        int r = optimizer.optimize(10);
        if (iters_done) iters_done[it] = r;
        int nGood = 0;
        for (int i = 0; i < n; i++) {
            auto* e = (EdgeSE3ProjectXYZOnlyPose*)edgesInfo[i].ptr;
            if (vBad[i]) e->computeError();
            vBad[i] = e->chi2() > edgesInfo[i].MaxChi;
            e->setLevel(vBad[i] ? 1 : 0);
            if (it >= 2) e->setRobustKernel(nullptr);
            if (!vBad[i]) nGood++;
        }
        for (auto me : marker_edges) {
            me->computeError();
            if (me->chi2() > Chi8D || it >= 2) me->setRobustKernel(nullptr);
        }
        if (nGood < 10 && n_markers == 0) break;
    }
    g2o::SE3Quat q = static_cast<VertexSE3Expmap*>(optimizer.vertex(0))->estimate();
    Eigen::Matrix<double, 4, 4> M = q.to_homogeneous_matrix();
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) out_pose44[4 * r + c] = (float)M(r, c);
    out_pose7[0] = q.rotation().x(); out_pose7[1] = q.rotation().y(); out_pose7[2] = q.rotation().z(); out_pose7[3] = q.rotation().w();
    for (int k = 0; k < 3; k++) out_pose7[4 + k] = q.translation()[k];
    int nbad = 0;
    for (int i = 0; i < n; i++) { bad[i] = vBad[i]; nbad += vBad[i]; }
    return n - nbad;
}
}
