// oracle/pose_opt.cpp -- TEST INFRASTRUCTURE ONLY: CPU restatement of Optimizer::PoseOptimization(Frame*) (src/Optimizer.cc:239-451), the motion-only
// bundle adjustment that follows every matcher call of the tracking thread.  PINNED against the reference's own src/Optimizer.cc + src/Converter.cc + vendored
// Thirdparty/g2o compiled unmodified (oracle/_ref/liboptimizer_ref.so, on the Eigen stand-in of g2o_shim/ because this image has no Eigen): same inlier count, outlier
// flags and float32 pose on every case of tests/test_optimizer_ref.py.  What is restated,
// from the vendored sources, file:line relative to Thirdparty/g2o/g2o:
//   * graph set-up, the four rounds of ten iterations from the INITIAL pose, chi-square re-classification (5.991 mono / 7.815 stereo, float
//     compares), robust kernel dropped after the third round, early exit when fewer than ten edges exist          src/Optimizer.cc:239-451
//   * OptimizationAlgorithmLevenberg::solve incl. lambda initialisation (tau = 1e-5), the rho / scale rule, up to ten trials per iteration and
//     the ORB-SLAM2 "_nBad" termination                                                       core/optimization_algorithm_levenberg.cpp:61-210
//   * SparseOptimizer::optimize (stops after the first non-OK iteration)                                       core/sparse_optimizer.cpp:354-419
//   * BaseUnaryEdge::constructQuadraticForm with the Huber kernel (rho' weighting only)   core/base_unary_edge.hpp:43-72, robust_kernel_impl.cpp:78-91
//   * EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose error and Jacobians (the stereo projection divides in float)
//                                                                                               types/types_six_dof_expmap.h:143-205, .cpp:262-364
//   * SE3Quat (exp, composition, normalisation) and Eigen's quaternion <-> matrix conversions                         types/se3quat.h:40-285
//   * a quirk that decides outlier flags: after a REJECTED trial g2o restores the estimate (pop) but not the edges' error vectors, so the
//     classification at the end of a round reads the errors of the last TRIED estimate for the edges that were active.
// The 6x6 system is solved by an LDL^T factorisation (Eigen::LDLT in g2o: pivoted; here unpivoted -- same solution to rounding).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#define SGO_API extern "C" __attribute__((visibility("default")))

namespace {

struct Quat { double x, y, z, w; };
struct SE3 { Quat r; double t[3]; };

Quat quat_from_matrix(const double R[9]) {       // Eigen::Quaterniond(Matrix3d)
    Quat q;
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t; t = 0.5 / t;
        q.x = (R[7] - R[5]) * t; q.y = (R[2] - R[6]) * t; q.z = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[3 * i + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[3 * i + i] - R[3 * j + j] - R[3 * k + k] + 1.0);
        double c[3];
        c[i] = 0.5 * t; t = 0.5 / t;
        q.w = (R[3 * k + j] - R[3 * j + k]) * t;
        c[j] = (R[3 * j + i] + R[3 * i + j]) * t; c[k] = (R[3 * k + i] + R[3 * i + k]) * t;
        q.x = c[0]; q.y = c[1]; q.z = c[2];
    }
    return q;
}

void quat_normalize_rotation(Quat& q) {          // SE3Quat::normalizeRotation
    if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

Quat quat_mul(const Quat& a, const Quat& b) {    // Eigen quaternion product
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}

void quat_rotate(const Quat& q, const double v[3], double out[3]) {     // Eigen QuaternionBase::_transformVector
    double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    const double c[3] = {q.y * uv[2] - q.z * uv[1], q.z * uv[0] - q.x * uv[2], q.x * uv[1] - q.y * uv[0]};
    for (int i = 0; i < 3; i++) out[i] = v[i] + q.w * uv[i] + c[i];
}

void quat_to_matrix(const Quat& q, double R[9]) {                       // Eigen toRotationMatrix
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

SE3 se3_from_Rt(const double R[9], const double t[3]) { SE3 s; s.r = quat_from_matrix(R); quat_normalize_rotation(s.r); std::memcpy(s.t, t, 24); return s; }

SE3 se3_exp(const double u[6]) {                 // SE3Quat::exp, se3quat.h:223-257
    const double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]};
    const double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    double O2[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double v = 0; for (int k = 0; k < 3; k++) v += O[3 * i + k] * O[3 * k + j]; O2[3 * i + j] = v; }
    double R[9], V[9];
    if (theta < 0.00001) {
        for (int i = 0; i < 9; i++) { R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta), c = (theta - std::sin(theta)) / std::pow(theta, 3);
        for (int i = 0; i < 9; i++) { const double I = i % 4 == 0 ? 1.0 : 0.0; R[i] = I + a * O[i] + b * O2[i]; V[i] = I + b * O[i] + c * O2[i]; }
    }
    double t[3];
    for (int i = 0; i < 3; i++) t[i] = V[3 * i] * up[0] + V[3 * i + 1] * up[1] + V[3 * i + 2] * up[2];
    SE3 s; s.r = quat_from_matrix(R); quat_normalize_rotation(s.r); std::memcpy(s.t, t, 24);
    return s;
}

SE3 se3_mul(const SE3& a, const SE3& b) {        // SE3Quat::operator*
    SE3 r = a;
    double rt[3]; quat_rotate(a.r, b.t, rt);
    for (int i = 0; i < 3; i++) r.t[i] += rt[i];
    r.r = quat_mul(a.r, b.r);
    quat_normalize_rotation(r.r);
    return r;
}

void se3_map(const SE3& s, const double X[3], double out[3]) { quat_rotate(s.r, X, out); for (int i = 0; i < 3; i++) out[i] += s.t[i]; }

bool ldlt_solve6(const double Hin[36], const double b[6], double x[6]) {
    double L[36], D[6];
    std::memset(L, 0, sizeof L);
    for (int j = 0; j < 6; j++) {
        double d = Hin[6 * j + j];
        for (int k = 0; k < j; k++) d -= L[6 * j + k] * L[6 * j + k] * D[k];
        if (!(d > 0)) return false;                 // LDLT::isPositive
        D[j] = d; L[6 * j + j] = 1;
        for (int i = j + 1; i < 6; i++) {
            double v = Hin[6 * i + j];
            for (int k = 0; k < j; k++) v -= L[6 * i + k] * L[6 * j + k] * D[k];
            L[6 * i + j] = v / d;
        }
    }
    double y[6];
    for (int i = 0; i < 6; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= L[6 * i + k] * y[k]; y[i] = v; }
    for (int i = 0; i < 6; i++) y[i] /= D[i];
    for (int i = 5; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < 6; k++) v -= L[6 * k + i] * x[k]; x[i] = v; }
    return true;
}

struct Edge {
    int idx; bool stereo; double obs[3]; double Xw[3]; double info; double delta; bool robust; int level; double err[3];
};

struct Problem {
    double fx, fy, cx, cy, bf;
    std::vector<Edge> edges;
    std::vector<int> active;
    SE3 est;

    void compute_error(Edge& e) const {
        double p[3]; se3_map(est, e.Xw, p);
        if (!e.stereo) {
            e.err[0] = e.obs[0] - (p[0] / p[2] * fx + cx);            // project2d then * fx + cx
            e.err[1] = e.obs[1] - (p[1] / p[2] * fy + cy);
            e.err[2] = 0;
        } else {
            const float invz = 1.0f / p[2];                            // types_six_dof_expmap.cpp:302: float
            const double r0 = p[0] * invz * fx + cx, r1 = p[1] * invz * fy + cy, r2 = r0 - bf * invz;
            e.err[0] = e.obs[0] - r0; e.err[1] = e.obs[1] - r1; e.err[2] = e.obs[2] - r2;
        }
    }
    static double chi2(const Edge& e) { return (e.err[0] * e.err[0] + e.err[1] * e.err[1] + (e.stereo ? e.err[2] * e.err[2] : 0.0)) * e.info; }
    void compute_active_errors() { for (int a : active) compute_error(edges[a]); }
    double active_robust_chi2() const {
        double chi = 0;
        for (int a : active) {
            const Edge& e = edges[a];
            const double c = chi2(e);
            if (e.robust) { const double dsqr = e.delta * e.delta; chi += c <= dsqr ? c : 2 * std::sqrt(c) * e.delta - dsqr; }
            else chi += c;
        }
        return chi;
    }
    void build_system(double H[36], double b[6]) const {
        std::memset(H, 0, 36 * 8); std::memset(b, 0, 6 * 8);
        for (int a : active) {
            const Edge& e = edges[a];
            double p[3]; se3_map(est, e.Xw, p);
            const double x = p[0], y = p[1], invz = 1.0 / p[2], invz_2 = invz * invz;
            double J[3][6];
            J[0][0] = x * y * invz_2 * fx; J[0][1] = -(1 + (x * x * invz_2)) * fx; J[0][2] = y * invz * fx; J[0][3] = -invz * fx; J[0][4] = 0; J[0][5] = x * invz_2 * fx;
            J[1][0] = (1 + y * y * invz_2) * fy; J[1][1] = -x * y * invz_2 * fy; J[1][2] = -x * invz * fy; J[1][3] = 0; J[1][4] = -invz * fy; J[1][5] = y * invz_2 * fy;
            const int D = e.stereo ? 3 : 2;
            if (e.stereo) {
                J[2][0] = J[0][0] - bf * y * invz_2; J[2][1] = J[0][1] + bf * x * invz_2; J[2][2] = J[0][2]; J[2][3] = J[0][3]; J[2][4] = 0; J[2][5] = J[0][5] - bf * invz_2;
            }
            double w = 1.0;                                  // rho[1]
            if (e.robust) { const double c = chi2(e), dsqr = e.delta * e.delta; if (c > dsqr) w = e.delta / std::sqrt(c); }
            for (int r = 0; r < 6; r++) {
                double s = 0;
                for (int d = 0; d < D; d++) s += J[d][r] * e.info * e.err[d];
                b[r] -= w * s;
                for (int c = 0; c < 6; c++) { double h = 0; for (int d = 0; d < D; d++) h += J[d][r] * (w * e.info) * J[d][c]; H[6 * r + c] += h; }
            }
        }
    }
};

}  // namespace

// has_mp[i]: mvpMapPoints[i] != NULL; xyz: GetWorldPos(); kp_xy / octave: mvKeysUn; uright: mvuRight (< 0: monocular observation).
// outlier (n, in/out is irrelevant: every participating entry is reset): mvbOutlier.  Returns nInitialCorrespondences - nBad.
SGO_API int sgo_pose_optimization(const float* Tcw_in, int n, const uint8_t* has_mp, const float* xyz, const float* kp_xy, const int32_t* octave, const float* uright,
                                  const float* inv_level_sigma2, float fx, float fy, float cx, float cy, float bf, float* Tcw_out, uint8_t* outlier) {
    Problem P;
    P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy; P.bf = bf;
    const float deltaMono = std::sqrt(5.991), deltaStereo = std::sqrt(7.815);
    int nInitial = 0;
    for (int i = 0; i < n; i++) {
        if (!has_mp[i]) continue;
        nInitial++;
        outlier[i] = 0;
        Edge e;
        e.idx = i; e.stereo = !(uright[i] < 0);
        e.obs[0] = kp_xy[2 * i]; e.obs[1] = kp_xy[2 * i + 1]; e.obs[2] = e.stereo ? uright[i] : 0;
        e.Xw[0] = xyz[3 * i]; e.Xw[1] = xyz[3 * i + 1]; e.Xw[2] = xyz[3 * i + 2];
        e.info = inv_level_sigma2[octave[i]];
        e.delta = e.stereo ? deltaStereo : deltaMono; e.robust = true; e.level = 0; e.err[0] = e.err[1] = e.err[2] = 0;
        P.edges.push_back(e);
    }
    for (int i = 0; i < 16; i++) Tcw_out[i] = Tcw_in[i];
    if (nInitial < 3) return 0;
    double R0[9], t0[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R0[3 * r + c] = Tcw_in[4 * r + c]; t0[r] = Tcw_in[4 * r + 3]; }
    const float chi2Mono[4] = {5.991f, 5.991f, 5.991f, 5.991f}, chi2Stereo[4] = {7.815f, 7.815f, 7.815f, 7.815f};
    int nBad = 0;
    for (int it = 0; it < 4; it++) {
        P.est = se3_from_Rt(R0, t0);
        P.active.clear();
        for (size_t a = 0; a < P.edges.size(); a++) if (P.edges[a].level == 0) P.active.push_back((int)a);
        if (!P.active.empty()) {
            // SparseOptimizer::optimize(10) with OptimizationAlgorithmLevenberg
            double lambda = 0, ni = 2; int nbad_lm = 0;
            bool ok = true;
            for (int i = 0; i < 10 && ok; i++) {
                P.compute_active_errors();
                double currentChi = P.active_robust_chi2(), tempChi = currentChi;
                const double iniChi = currentChi;
                double H[36], b[6];
                P.build_system(H, b);
                if (i == 0) {
                    double md = 0; for (int j = 0; j < 6; j++) md = std::max(std::fabs(H[6 * j + j]), md);
                    lambda = 1e-5 * md; ni = 2; nbad_lm = 0;
                }
                double rho = 0; int qmax = 0;
                do {
                    const SE3 backup = P.est;
                    double Hl[36]; std::memcpy(Hl, H, sizeof Hl);
                    for (int j = 0; j < 6; j++) Hl[6 * j + j] += lambda;
                    double x[6] = {0, 0, 0, 0, 0, 0};
                    const bool ok2 = ldlt_solve6(Hl, b, x);
                    P.est = se3_mul(se3_exp(x), P.est);
                    P.compute_active_errors();
                    tempChi = P.active_robust_chi2();
                    if (!ok2) tempChi = std::numeric_limits<double>::max();
                    rho = currentChi - tempChi;
                    double scale = 0;
                    for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
                    scale += 1e-3;
                    rho /= scale;
                    if (rho > 0 && std::isfinite(tempChi)) {
                        double alpha = 1. - std::pow((2 * rho - 1), 3);
                        alpha = std::min(alpha, 2. / 3.);
                        const double sf = std::max(1. / 3., alpha);
                        lambda *= sf; ni = 2; currentChi = tempChi;
                    } else {
                        lambda *= ni; ni *= 2;
                        P.est = backup;                          // pop(): the estimate comes back, the edges keep the errors of the rejected trial
                    }
                    qmax++;
                } while (rho < 0 && qmax < 10);
                if (qmax == 10 || rho == 0) { ok = false; break; }
                if ((iniChi - currentChi) * 1e3 < iniChi) nbad_lm++; else nbad_lm = 0;
                if (nbad_lm >= 3) { ok = false; break; }
            }
        }
        nBad = 0;
        for (Edge& e : P.edges) {
            if (outlier[e.idx]) P.compute_error(e);
            const float chi2 = (float)Problem::chi2(e);
            if (chi2 > (e.stereo ? chi2Stereo[it] : chi2Mono[it])) { outlier[e.idx] = 1; e.level = 1; nBad++; }
            else { outlier[e.idx] = 0; e.level = 0; }
            if (it == 2) e.robust = false;
        }
        if (P.edges.size() < 10) break;
    }
    double R[9]; quat_to_matrix(P.est.r, R);
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) Tcw_out[4 * r + c] = (float)R[3 * r + c]; Tcw_out[4 * r + 3] = (float)P.est.t[r]; }
    Tcw_out[12] = Tcw_out[13] = Tcw_out[14] = 0.f; Tcw_out[15] = 1.f;
    return nInitial - nBad;
}
