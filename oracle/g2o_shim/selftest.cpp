// C entry points over the Eigen stand-in (g2o_shim/Eigen) so that tests/test_mini_eigen.py can check its primitives against numpy: the reference pin of
// Optimizer::PoseOptimization (oracle/_ref/liboptimizer_ref.so) is only as good as this header.  TEST INFRASTRUCTURE.
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <cstring>
using namespace Eigen;
#define ME_API extern "C" __attribute__((visibility("default")))

static MatrixXd load(const double* a, int r, int c) { MatrixXd m(r, c); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m(i, j) = a[i * c + j]; return m; }
static void store(const MatrixXd& m, double* o) { for (int i = 0; i < m.rows(); ++i) for (int j = 0; j < m.cols(); ++j) o[i * m.cols() + j] = m(i, j); }

// x = A^-1 b by LLT, pivoted LDLT and LU (row-major inputs); returns LDLT's isPositive
ME_API int me_solve(int n, const double* A, const double* b, double* x_llt, double* x_ldlt, double* x_lu) {
    MatrixXd a = load(A, n, n); VectorXd rhs(n); for (int i = 0; i < n; ++i) rhs[i] = b[i];
    VectorXd x1 = a.llt().solve(rhs); LDLT<MatrixXd> ld; ld.compute(a); VectorXd x2 = ld.solve(rhs); VectorXd x3 = a.lu().solve(rhs);
    for (int i = 0; i < n; ++i) { x_llt[i] = x1[i]; x_ldlt[i] = x2[i]; x_lu[i] = x3[i]; }
    return ld.isPositive() ? 1 : 0;
}
ME_API double me_inverse_det(int n, const double* A, double* inv) { MatrixXd a = load(A, n, n); store(a.inverse(), inv); return a.determinant(); }
ME_API void me_inverse3(const double* A, double* inv, double* det) { Matrix3d a; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a(i, j) = A[3 * i + j]; Matrix3d r = a.inverse(); *det = a.determinant(); for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) inv[3 * i + j] = r(i, j); }
// C = A(r x k) * B(k x c) + D^T where D is c x r; also exercises blocks, transposes, noalias and Map
ME_API void me_gemm(int r, int k, int c, const double* A, const double* B, const double* D, double* out) {
    MatrixXd a = load(A, r, k), b = load(B, k, c), d = load(D, c, r);
    MatrixXd res = MatrixXd::Zero(r, c);
    res.noalias() += a * b;
    res.block(0, 0, r, c) += d.transpose();
    std::vector<double> colmajor((size_t)r * c);
    Map<MatrixXd> m(colmajor.data(), r, c); m = res;
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) out[i * c + j] = colmajor[i + (size_t)j * r];
}
// rotation matrix -> quaternion (x, y, z, w) -> rotation matrix; q * v; (q1 * q2) as a matrix
ME_API void me_quaternion(const double* R, const double* v, const double* R2, double* q4, double* Rback, double* qv, double* R12) {
    Matrix3d r, r2; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { r(i, j) = R[3 * i + j]; r2(i, j) = R2[3 * i + j]; }
    Quaterniond q(r), q2(r2);
    q4[0] = q.x(); q4[1] = q.y(); q4[2] = q.z(); q4[3] = q.w();
    Matrix3d rb = q.toRotationMatrix(), m12 = (q * q2).toRotationMatrix();
    Vector3d w = q * Vector3d(v[0], v[1], v[2]);
    for (int i = 0; i < 3; ++i) { qv[i] = w[i]; for (int j = 0; j < 3; ++j) { Rback[3 * i + j] = rb(i, j); R12[3 * i + j] = m12(i, j); } }
}
ME_API void me_eigenvalues(int n, const double* A, double* ev) { MatrixXd a = load(A, n, n); SelfAdjointEigenSolver<MatrixXd> es; es.compute(a, EigenvaluesOnly); for (int i = 0; i < n; ++i) ev[i] = es.eigenvalues()(i); }
// fixed-size path: H (6x6) += J^T W J, b -= J^T W e for a 2x6 Jacobian (the shapes of BaseUnaryEdge::constructQuadraticForm); comma initialiser; diagonal().array()
ME_API void me_quadratic_form(const double* J, const double* W, const double* e, double lambda, double* H, double* b) {
    Matrix<double, 2, 6> j; Matrix2d w; Vector2d err;
    for (int i = 0; i < 2; ++i) { for (int k = 0; k < 6; ++k) j(i, k) = J[6 * i + k]; for (int k = 0; k < 2; ++k) w(i, k) = W[2 * i + k]; }
    err << e[0], e[1];
    Matrix<double, 6, 6> h = Matrix<double, 6, 6>::Zero(); Matrix<double, 6, 1> g; g.setZero();
    g.noalias() -= j.transpose() * w * err;
    h.noalias() += j.transpose() * w * j;
    h.diagonal().array() += lambda;
    for (int i = 0; i < 6; ++i) { b[i] = g[i]; for (int k = 0; k < 6; ++k) H[6 * i + k] = h(i, k); }
}
