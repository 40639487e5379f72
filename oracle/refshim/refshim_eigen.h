// refshim — TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// The subset of the Eigen 3.3 API that the reference's sources use (include/liodom/factors.hpp,
// src/laser_odometry.cc, src/map.cc, src/stats.cc, src/shared_data.cc), so that those files compile
// UNMODIFIED where they lie under /root/reference.  Eigen is a third-party dependency that is not
// vendored in the reference and not installed in this image; what is restated here is its published
// arithmetic (Eigen 3.3.7: quaternion <-> matrix, quaternion * vector, slerp, isometry product /
// inverse), value types only, no expression templates.  Nothing under liodom_b200/ includes this.
#pragma once
#include <cmath>
#include <cstddef>
#include <limits>

namespace Eigen {

template <typename T>
struct NumTraits {
  static T epsilon() { return std::numeric_limits<T>::epsilon(); }
};

namespace refshim_detail {
using std::abs;
using std::acos;
using std::sin;
using std::sqrt;
template <typename T> inline T do_sqrt(const T& x) { return sqrt(x); }   // ADL finds ceres::sqrt for Jets
template <typename T> inline T do_sin(const T& x) { return sin(x); }
template <typename T> inline T do_acos(const T& x) { return acos(x); }
template <typename T> inline T do_abs(const T& x) { return abs(x); }
}  // namespace refshim_detail

template <typename T, int R, int C>
class Matrix {
 public:
  typedef T Scalar;
  T m[R * C];   // row-major storage (an implementation detail: only coefficient access is exposed)
  Matrix() { for (int i = 0; i < R * C; ++i) m[i] = T(0); }
  // fixed-size vector constructors (Vector3d(x,y,z), Matrix<T,3,1>{a,b,c})
  Matrix(const T& a, const T& b, const T& c) { static_assert(R * C == 3, "3-vector"); m[0] = a; m[1] = b; m[2] = c; }
  Matrix(const T& a, const T& b, const T& c, const T& d) { static_assert(R * C == 4, "4-vector"); m[0] = a; m[1] = b; m[2] = c; m[3] = d; }
  static Matrix Zero() { return Matrix(); }
  static Matrix Identity() { Matrix I; for (int i = 0; i < (R < C ? R : C); ++i) I.m[i * C + i] = T(1); return I; }
  T& operator()(int i, int j) { return m[i * C + j]; }
  const T& operator()(int i, int j) const { return m[i * C + j]; }
  T& operator()(int i) { return m[i]; }
  const T& operator()(int i) const { return m[i]; }
  T& operator[](int i) { return m[i]; }
  const T& operator[](int i) const { return m[i]; }
  T& x() { return m[0]; } T& y() { return m[1]; } T& z() { return m[2]; }
  const T& x() const { return m[0]; } const T& y() const { return m[1]; } const T& z() const { return m[2]; }
  Matrix operator+(const Matrix& o) const { Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = m[i] + o.m[i]; return r; }
  Matrix operator-(const Matrix& o) const { Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = m[i] - o.m[i]; return r; }
  Matrix operator-() const { Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = -m[i]; return r; }
  Matrix operator/(const T& s) const { Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = m[i] / s; return r; }
  Matrix operator*(const T& s) const { Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = m[i] * s; return r; }
  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; ++i) m[i] = m[i] + o.m[i]; return *this; }
  Matrix<T, C, R> transpose() const { Matrix<T, C, R> r; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) r.m[j * R + i] = m[i * C + j]; return r; }
  // coefficient-based product, inner index ascending: ((a0 b0 + a1 b1) + a2 b2)
  template <int K>
  Matrix<T, R, K> operator*(const Matrix<T, C, K>& o) const {
    Matrix<T, R, K> r;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < K; ++j) {
        T s = m[i * C] * o.m[j];
        for (int k = 1; k < C; ++k) s = s + m[i * C + k] * o.m[k * K + j];
        r.m[i * K + j] = s;
      }
    return r;
  }
  Matrix cross(const Matrix& b) const {
    static_assert(R * C == 3, "3-vector");
    return Matrix(m[1] * b.m[2] - m[2] * b.m[1], m[2] * b.m[0] - m[0] * b.m[2], m[0] * b.m[1] - m[1] * b.m[0]);
  }
  T dot(const Matrix& o) const { T s = m[0] * o.m[0]; for (int i = 1; i < R * C; ++i) s = s + m[i] * o.m[i]; return s; }
  T squaredNorm() const { return dot(*this); }
  T norm() const { return refshim_detail::do_sqrt(squaredNorm()); }
  T trace() const { T s = m[0]; for (int i = 1; i < (R < C ? R : C); ++i) s = s + m[i * C + i]; return s; }
};
template <typename T, int R, int C>
inline Matrix<T, R, C> operator*(const T& s, const Matrix<T, R, C>& a) { Matrix<T, R, C> r; for (int i = 0; i < R * C; ++i) r.m[i] = s * a.m[i]; return r; }

typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 4, 4> Matrix4f;

// Eigen/src/Geometry/Quaternion.h
template <typename T>
class Quaternion {
 public:
  typedef Matrix<T, 3, 1> Vector3;
  typedef Matrix<T, 3, 3> Matrix3;
  T c[4];   // x, y, z, w (Eigen's coefficient order)
  Quaternion() { c[0] = c[1] = c[2] = T(0); c[3] = T(1); }
  Quaternion(const T& w, const T& x, const T& y, const T& z) { c[0] = x; c[1] = y; c[2] = z; c[3] = w; }
  explicit Quaternion(const Matrix3& mat) {   // quaternionbase_assign_impl<Other,3,3>
    T t = mat.trace();
    if (t > T(0)) {
      t = refshim_detail::do_sqrt(t + T(1.0));
      c[3] = T(0.5) * t;
      t = T(0.5) / t;
      c[0] = (mat(2, 1) - mat(1, 2)) * t;
      c[1] = (mat(0, 2) - mat(2, 0)) * t;
      c[2] = (mat(1, 0) - mat(0, 1)) * t;
    } else {
      int i = 0;
      if (mat(1, 1) > mat(0, 0)) i = 1;
      if (mat(2, 2) > mat(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = refshim_detail::do_sqrt(mat(i, i) - mat(j, j) - mat(k, k) + T(1.0));
      c[i] = T(0.5) * t;
      t = T(0.5) / t;
      c[3] = (mat(k, j) - mat(j, k)) * t;
      c[j] = (mat(j, i) + mat(i, j)) * t;
      c[k] = (mat(k, i) + mat(i, k)) * t;
    }
  }
  static Quaternion Identity() { return Quaternion(T(1), T(0), T(0), T(0)); }
  T& x() { return c[0]; } T& y() { return c[1]; } T& z() { return c[2]; } T& w() { return c[3]; }
  const T& x() const { return c[0]; } const T& y() const { return c[1]; } const T& z() const { return c[2]; } const T& w() const { return c[3]; }
  Vector3 vec() const { return Vector3(c[0], c[1], c[2]); }
  T dot(const Quaternion& o) const { return ((c[0] * o.c[0] + c[1] * o.c[1]) + c[2] * o.c[2]) + c[3] * o.c[3]; }
  Matrix3 toRotationMatrix() const {
    Matrix3 res;
    const T tx = T(2) * c[0], ty = T(2) * c[1], tz = T(2) * c[2];
    const T twx = tx * c[3], twy = ty * c[3], twz = tz * c[3];
    const T txx = tx * c[0], txy = ty * c[0], txz = tz * c[0];
    const T tyy = ty * c[1], tyz = tz * c[1], tzz = tz * c[2];
    res(0, 0) = T(1) - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
    res(1, 0) = txy + twz; res(1, 1) = T(1) - (txx + tzz); res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = T(1) - (txx + tyy);
    return res;
  }
  // QuaternionBase::_transformVector
  Vector3 operator*(const Vector3& v) const {
    Vector3 uv = vec().cross(v);
    uv += uv;
    return v + c[3] * uv + vec().cross(uv);
  }
  // QuaternionBase::slerp (Eigen 3.3)
  Quaternion slerp(const T& t, const Quaternion& other) const {
    static const T one = T(1) - NumTraits<T>::epsilon();
    const T d = dot(other);
    const T absD = refshim_detail::do_abs(d);
    T scale0, scale1;
    if (absD >= one) {
      scale0 = T(1) - t;
      scale1 = t;
    } else {
      const T theta = refshim_detail::do_acos(absD);
      const T sinTheta = refshim_detail::do_sin(theta);
      scale0 = refshim_detail::do_sin((T(1) - t) * theta) / sinTheta;
      scale1 = refshim_detail::do_sin(t * theta) / sinTheta;
    }
    if (d < T(0)) scale1 = -scale1;
    Quaternion r;
    for (int k = 0; k < 4; ++k) r.c[k] = scale0 * c[k] + scale1 * other.c[k];
    return r;
  }
};
typedef Quaternion<double> Quaterniond;

enum TransformTraits { Isometry = 1, Affine = 2 };

// Eigen::Transform<double,3,Isometry>: 4x4 matrix whose last row stays (0,0,0,1)
template <typename T, int Dim, int Mode>
class Transform {
  static_assert(Dim == 3, "3-D transforms only");
 public:
  Matrix<T, 4, 4> M;
  Transform() : M(Matrix<T, 4, 4>::Identity()) {}
  explicit Transform(const Matrix<T, 4, 4>& mat) : M(mat) {}
  static Transform Identity() { return Transform(); }
  struct LinearRef {
    Matrix<T, 4, 4>* M;
    LinearRef& operator=(const Matrix<T, 3, 3>& r) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) (*M)(i, j) = r(i, j); return *this; }
    operator Matrix<T, 3, 3>() const { Matrix<T, 3, 3> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = (*M)(i, j); return r; }
  };
  struct TranslationRef {
    Matrix<T, 4, 4>* M;
    TranslationRef& operator=(const Matrix<T, 3, 1>& t) { for (int i = 0; i < 3; ++i) (*M)(i, 3) = t(i); return *this; }
    operator Matrix<T, 3, 1>() const { return Matrix<T, 3, 1>((*M)(0, 3), (*M)(1, 3), (*M)(2, 3)); }
    T& x() { return (*M)(0, 3); } T& y() { return (*M)(1, 3); } T& z() { return (*M)(2, 3); }
  };
  LinearRef linear() { return LinearRef{&M}; }
  Matrix<T, 3, 3> linear() const { Matrix<T, 3, 3> r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = M(i, j); return r; }
  Matrix<T, 3, 3> rotation() const { return linear(); }   // Isometry mode: rotation() is linear()
  TranslationRef translation() { return TranslationRef{&M}; }
  Matrix<T, 3, 1> translation() const { return Matrix<T, 3, 1>(M(0, 3), M(1, 3), M(2, 3)); }
  Matrix<T, 4, 4>& matrix() { return M; }
  const Matrix<T, 4, 4>& matrix() const { return M; }
  T operator()(int i, int j) const { return M(i, j); }
  // transform_transform_product_impl: linear = La Lb; translation = La tb + ta
  Transform operator*(const Transform& o) const {
    Transform res;
    const Matrix<T, 3, 3> L = linear() * o.linear();
    const Matrix<T, 3, 1> t = linear() * o.translation() + translation();
    res.linear() = L;
    res.translation() = t;
    return res;
  }
  // Transform::inverse(Isometry): (R', -(R' t))
  Transform inverse() const {
    Transform res;
    const Matrix<T, 3, 3> Rt = linear().transpose();
    res.linear() = Rt;
    res.translation() = -(Rt * translation());
    return res;
  }
};
typedef Transform<double, 3, Isometry> Isometry3d;

// Eigen::SelfAdjointEigenSolver<Matrix3d>: eigenvalues ascending.  Eigen tridiagonalises and runs
// implicit symmetric QR steps; any FP64 method of equal accuracy gives the same values to ~eps |A|
// (SURVEY.md App. A.6).  Cyclic Jacobi here.
template <typename MatrixType>
class SelfAdjointEigenSolver {
 public:
  explicit SelfAdjointEigenSolver(const MatrixType& A) {
    double a00 = A(0, 0), a01 = A(1, 0), a02 = A(2, 0), a11 = A(1, 1), a12 = A(2, 1), a22 = A(2, 2);   // lower triangle, as Eigen reads it
    for (int sweep = 0; sweep < 12; ++sweep) {
      const double off = a01 * a01 + a02 * a02 + a12 * a12;
      const double diag = a00 * a00 + a11 * a11 + a22 * a22;
      if (off <= 1e-32 * diag || off == 0.0) break;
      rotate(a00, a11, a01, a02, a12);   // (p,q) = (0,1)
      rotate(a00, a22, a02, a01, a12);   // (0,2)
      rotate(a11, a22, a12, a01, a02);   // (1,2)
    }
    double e0 = a00, e1 = a11, e2 = a22, t;
    if (e0 > e1) { t = e0; e0 = e1; e1 = t; }
    if (e1 > e2) { t = e1; e1 = e2; e2 = t; }
    if (e0 > e1) { t = e0; e0 = e1; e1 = t; }
    w_ = Matrix<double, 3, 1>(e0, e1, e2);
  }
  const Matrix<double, 3, 1>& eigenvalues() const { return w_; }

 private:
  // Jacobi rotation annihilating a_pq; x and y are the other two off-diagonal entries, the one that
  // shares index p's partner first (a_pr, a_qr).
  static void rotate(double& app, double& aqq, double& apq, double& x, double& y) {
    if (apq == 0.0) return;
    const double theta = (aqq - app) / (2.0 * apq);
    const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
    const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
    const double npp = app - t * apq, nqq = aqq + t * apq;
    const double nx = c * x - s * y, ny = s * x + c * y;
    app = npp; aqq = nqq; apq = 0.0; x = nx; y = ny;
  }
  Matrix<double, 3, 1> w_;
};

inline void initParallel() {}

}  // namespace Eigen
