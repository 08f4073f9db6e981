// Quaternion / AngleAxis / Transform of the Eigen stand-in (see ../Core).  Formulas as published in Eigen 3 (Geometry/Quaternion.h, RotationBase): rotation matrix ->
// quaternion by the trace / largest-diagonal branches, quaternion -> matrix with the doubled products, q * v = v + w*(2 q_v x v) + q_v x (2 q_v x v).  TEST INFRASTRUCTURE.
#ifndef SGS_MINI_EIGEN_GEOMETRY
#define SGS_MINI_EIGEN_GEOMETRY
namespace Eigen {

template <class S> class AngleAxis;

template <class S> class Quaternion {
    Matrix<S, 4, 1> c;            // x, y, z, w
public:
    typedef S Scalar;
    typedef Matrix<S, 3, 1> Vector3;
    typedef Matrix<S, 3, 3> Matrix3;
    Quaternion() {}
    Quaternion(const S& w, const S& x, const S& y, const S& z) { c[0] = x; c[1] = y; c[2] = z; c[3] = w; }
    Quaternion(const Quaternion& o) : c(o.c) {}
    template <class D> explicit Quaternion(const MatrixBase<D>& m) { *this = m; }
    explicit Quaternion(const AngleAxis<S>& aa);
    Quaternion& operator=(const Quaternion& o) { c = o.c; return *this; }
    template <class D> Quaternion& operator=(const MatrixBase<D>& mat) {
        if (mat.rows() == 4 && mat.cols() == 1) { for (int i = 0; i < 4; ++i) c[i] = mat.lin(i); return *this; }
        // Eigen: quaternionbase_assign_impl<Other, 3, 3>
        S t = mat.coeff(0, 0) + mat.coeff(1, 1) + mat.coeff(2, 2);
        if (t > S(0)) {
            t = std::sqrt(t + S(1.0));
            w() = S(0.5) * t;
            t = S(0.5) / t;
            x() = (mat.coeff(2, 1) - mat.coeff(1, 2)) * t;
            y() = (mat.coeff(0, 2) - mat.coeff(2, 0)) * t;
            z() = (mat.coeff(1, 0) - mat.coeff(0, 1)) * t;
        } else {
            int i = 0;
            if (mat.coeff(1, 1) > mat.coeff(0, 0)) i = 1;
            if (mat.coeff(2, 2) > mat.coeff(i, i)) i = 2;
            int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(mat.coeff(i, i) - mat.coeff(j, j) - mat.coeff(k, k) + S(1.0));
            c[i] = S(0.5) * t;
            t = S(0.5) / t;
            w() = (mat.coeff(k, j) - mat.coeff(j, k)) * t;
            c[j] = (mat.coeff(j, i) + mat.coeff(i, j)) * t;
            c[k] = (mat.coeff(k, i) + mat.coeff(i, k)) * t;
        }
        return *this;
    }
    static Quaternion Identity() { return Quaternion(S(1), S(0), S(0), S(0)); }
    Quaternion& setIdentity() { c[0] = c[1] = c[2] = S(0); c[3] = S(1); return *this; }
    S x() const { return c[0]; } S y() const { return c[1]; } S z() const { return c[2]; } S w() const { return c[3]; }
    S& x() { return c[0]; } S& y() { return c[1]; } S& z() { return c[2]; } S& w() { return c[3]; }
    Matrix<S, 4, 1>& coeffs() { return c; }
    const Matrix<S, 4, 1>& coeffs() const { return c; }
    Block<S, 3, 1> vec() const { return c.template head<3>(); }
    S squaredNorm() const { return c.squaredNorm(); }
    S norm() const { return c.norm(); }
    void normalize() { c.normalize(); }
    Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
    Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
    Quaternion inverse() const {
        S n2 = squaredNorm();
        if (n2 > S(0)) { Quaternion q = conjugate(); q.c /= n2; return q; }
        Quaternion q; q.c.setZero(); return q;
    }
    S dot(const Quaternion& o) const { return c.dot(o.c); }
    Quaternion operator*(const Quaternion& b) const {
        const Quaternion& a = *this;
        return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                          a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                          a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                          a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    }
    Quaternion& operator*=(const Quaternion& b) { *this = *this * b; return *this; }
    template <class D> Vector3 operator*(const MatrixBase<D>& v) const { return _transformVector(v); }
    template <class D> Vector3 _transformVector(const MatrixBase<D>& v) const {
        Vector3 qv; qv[0] = x(); qv[1] = y(); qv[2] = z();
        Vector3 vv; vv[0] = v.lin(0); vv[1] = v.lin(1); vv[2] = v.lin(2);
        Vector3 uv = qv.cross(vv);
        uv += uv;
        return vv + w() * uv + qv.cross(uv);
    }
    Matrix3 toRotationMatrix() const {
        Matrix3 res;
        const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
        const S twx = tx * w(), twy = ty * w(), twz = tz * w();
        const S txx = tx * x(), txy = ty * x(), txz = tz * x();
        const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        res(0, 0) = S(1) - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
        res(1, 0) = txy + twz; res(1, 1) = S(1) - (txx + tzz); res(1, 2) = tyz - twx;
        res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = S(1) - (txx + tyy);
        return res;
    }
    Matrix3 matrix() const { return toRotationMatrix(); }
    template <class T> Quaternion<T> cast() const { return Quaternion<T>((T)w(), (T)x(), (T)y(), (T)z()); }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <class S> class AngleAxis {
    Matrix<S, 3, 1> ax; S ang;
public:
    AngleAxis() : ang(0) {}
    template <class D> AngleAxis(const S& a, const MatrixBase<D>& v) : ax(v), ang(a) {}
    explicit AngleAxis(const Quaternion<S>& q) {
        S n = q.vec().norm();
        if (n < std::numeric_limits<S>::epsilon()) n = std::sqrt(q.vec().squaredNorm());
        if (n != S(0)) { ang = S(2) * std::atan2(n, std::abs(q.w())); if (q.w() < S(0)) n = -n; ax = q.vec() / n; }
        else { ang = S(0); ax.setZero(); ax[0] = S(1); }
    }
    S angle() const { return ang; }
    const Matrix<S, 3, 1>& axis() const { return ax; }
    Matrix<S, 3, 3> toRotationMatrix() const {
        Matrix<S, 3, 3> res;
        const S sn = std::sin(ang), cs = std::cos(ang);
        Matrix<S, 3, 1> sin_axis = sn * ax, cos1_axis = (S(1) - cs) * ax;
        S tmp;
        tmp = cos1_axis.x() * ax.y(); res(0, 1) = tmp - sin_axis.z(); res(1, 0) = tmp + sin_axis.z();
        tmp = cos1_axis.x() * ax.z(); res(0, 2) = tmp + sin_axis.y(); res(2, 0) = tmp - sin_axis.y();
        tmp = cos1_axis.y() * ax.z(); res(1, 2) = tmp - sin_axis.x(); res(2, 1) = tmp + sin_axis.x();
        res(0, 0) = cos1_axis.x() * ax.x() + cs; res(1, 1) = cos1_axis.y() * ax.y() + cs; res(2, 2) = cos1_axis.z() * ax.z() + cs;
        return res;
    }
};
typedef AngleAxis<double> AngleAxisd;
template <class S> Quaternion<S>::Quaternion(const AngleAxis<S>& aa) {
    const S ha = S(0.5) * aa.angle();
    w() = std::cos(ha); const S s = std::sin(ha);
    x() = s * aa.axis().x(); y() = s * aa.axis().y(); z() = s * aa.axis().z();
}

// (Dim+1) x (Dim+1) homogeneous transform; only what g2o's SE3Quat -> Isometry3d conversion and the eigen_types typedefs need
template <class S, int Dim, int Mode, int Options = 0> class Transform {
    Matrix<S, Dim + 1, Dim + 1> m;
public:
    Transform() { m.setIdentity(); }
    template <class Q> explicit Transform(const Quaternion<Q>& q) { m.setIdentity(); m.template block<3, 3>(0, 0) = q.toRotationMatrix(); }
    template <class D> explicit Transform(const MatrixBase<D>& o) { m = o; }
    static Transform Identity() { return Transform(); }
    Matrix<S, Dim + 1, Dim + 1>& matrix() { return m; }
    const Matrix<S, Dim + 1, Dim + 1>& matrix() const { return m; }
    Block<S, Dim, 1> translation() const { return m.template block<Dim, 1>(0, Dim); }
    Block<S, Dim, Dim> linear() const { return m.template block<Dim, Dim>(0, 0); }
    Block<S, Dim, Dim> rotation() const { return linear(); }
    S operator()(int i, int j) const { return m(i, j); }
    S& operator()(int i, int j) { return m(i, j); }
    Transform operator*(const Transform& o) const { Transform t; t.m = m * o.m; return t; }
    template <class D> Matrix<S, Dim, 1> operator*(const MatrixBase<D>& v) const { Matrix<S, Dim, 1> r = linear() * v + translation(); return r; }
    Transform inverse() const { Transform t; t.m = m.inverse(); return t; }
};
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<double, 2, Isometry> Isometry2d;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<double, 2, Affine> Affine2d;

}  // namespace Eigen
#endif
