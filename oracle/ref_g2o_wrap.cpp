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

extern "C" {

// obs_ur[i] is used when obs_stereo[i] != 0.  out_pose7: qx qy qz qw tx ty tz (f64); out_pose44: what getResults stores
// (CV_32F 4x4); out_chi2 / out_level / out_depth_pos: per observation state after optimize(); out_bad: the
// getBadAssociations() predicate (:506-521).  trace (optional, 64 doubles): chi2 after each outer iteration.
int ref_ba_optimize(int n_poses, const float* poses44, const uint8_t* fixed, int n_points, const float* points3, int n_obs,
                    const int32_t* obs_pose, const int32_t* obs_point, const float* obs_uv, const float* obs_ur,
                    const uint8_t* obs_stereo, const float* obs_inv_sigma2, float fx, float fy, float cx, float cy, float bf,
                    int n_iters, double* out_pose7, float* out_pose44, double* out_point3, double* out_chi2,
                    uint8_t* out_level, uint8_t* out_bad, int* iters_done, double* trace /* optional: per outer iteration {chi2, LM trials}, 2 x 64 */) {
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
            e->fx = fx; e->fy = fy; e->cx = cx; e->cy = cy;
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
            e->fx = fx; e->fy = fy; e->cx = cx; e->cy = cy; e->bf = bf;
            Optimizer->addEdge(e);
            edges[i] = e;
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
}
