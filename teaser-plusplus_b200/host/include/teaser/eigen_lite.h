// eigen_lite.h — a deliberately tiny stand-in for the handful of Eigen types that appear in the public
// signatures of teaser::RobustRegistrationSolver (teaser/include/teaser/registration.h in the reference).
// It is used ONLY when the real Eigen is not installed (`__has_include(<Eigen/Core>)` fails — as in this
// image); with Eigen present the façade uses the real types and this file is not included.
// Dense, column-major, value semantics; enough for callers written like examples/teaser_cpp_ply.cc.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <memory>

#ifndef EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#endif

namespace Eigen {

constexpr int Dynamic = -1;
using Index = std::ptrdiff_t;

template <typename T, int R, int C>
class Matrix {
 public:
  using Scalar = T;
  Matrix() : rows_(R > 0 ? R : 0), cols_(C > 0 ? C : 0) { alloc(); }
  Matrix(Index r, Index c) : rows_(R > 0 ? R : r), cols_(C > 0 ? C : c) { alloc(); }
  explicit Matrix(Index n) : rows_(R > 0 ? R : (C == 1 ? n : 1)), cols_(C > 0 ? C : (C == 1 ? 1 : n)) { alloc(); }
  // fixed-size vectors: Vector3d(x, y, z)
  Matrix(T x, T y, T z) : rows_(R > 0 ? R : 3), cols_(C > 0 ? C : 1) {
    alloc();
    assert(size() == 3);
    d_[0] = x;
    d_[1] = y;
    d_[2] = z;
  }
  Matrix(const Matrix& o) : rows_(o.rows_), cols_(o.cols_) {
    alloc();
    std::copy(o.d_.get(), o.d_.get() + size(), d_.get());
  }
  Matrix(Matrix&& o) noexcept = default;
  template <typename U, int R2, int C2>
  Matrix(const Matrix<U, R2, C2>& o) : rows_(o.rows()), cols_(o.cols()) {
    assert((R < 0 || R == o.rows()) && (C < 0 || C == o.cols()));
    alloc();
    for (Index i = 0; i < size(); ++i) d_[i] = static_cast<T>(o.data()[i]);
  }
  Matrix& operator=(const Matrix& o) {
    if (this != &o) {
      resize(o.rows_, o.cols_);
      std::copy(o.d_.get(), o.d_.get() + size(), d_.get());
    }
    return *this;
  }
  Matrix& operator=(Matrix&& o) noexcept = default;

  Index rows() const { return rows_; }
  Index cols() const { return cols_; }
  Index size() const { return rows_ * cols_; }
  T* data() { return d_.get(); }
  const T* data() const { return d_.get(); }

  void resize(Index r, Index c) {
    if (R > 0) r = R;
    if (C > 0) c = C;
    if (r == rows_ && c == cols_ && d_) return;
    rows_ = r;
    cols_ = c;
    alloc();
  }
  void resize(Index n) { (C == 1) ? resize(n, 1) : resize(1, n); }

  T& operator()(Index r, Index c) { return d_[c * rows_ + r]; }
  const T& operator()(Index r, Index c) const { return d_[c * rows_ + r]; }
  T& operator()(Index i) { return d_[i]; }
  const T& operator()(Index i) const { return d_[i]; }
  T& operator[](Index i) { return d_[i]; }
  const T& operator[](Index i) const { return d_[i]; }

  Matrix& setZero() {
    std::fill(d_.get(), d_.get() + size(), T(0));
    return *this;
  }
  Matrix& setOnes() {
    std::fill(d_.get(), d_.get() + size(), T(1));
    return *this;
  }
  Matrix& setConstant(T v) {
    std::fill(d_.get(), d_.get() + size(), v);
    return *this;
  }
  Matrix& setIdentity() {
    setZero();
    for (Index i = 0; i < std::min(rows_, cols_); ++i) (*this)(i, i) = T(1);
    return *this;
  }
  static Matrix Zero() { return Matrix().setZero(); }
  static Matrix Zero(Index r, Index c) { return Matrix(r, c).setZero(); }
  static Matrix Ones(Index r, Index c) { return Matrix(r, c).setOnes(); }
  static Matrix Identity() { return Matrix().setIdentity(); }
  static Matrix Identity(Index r, Index c) { return Matrix(r, c).setIdentity(); }

  // ---- column access (read / assign a whole column)
  class ColXpr {
   public:
    ColXpr(Matrix& m, Index j) : m_(m), j_(j) {}
    template <int R2>
    ColXpr& operator=(const Matrix<T, R2, 1>& v) {
      assert(v.rows() == m_.rows());
      for (Index r = 0; r < m_.rows(); ++r) m_(r, j_) = v(r);
      return *this;
    }
    ColXpr& operator=(const ColXpr& o) {
      for (Index r = 0; r < m_.rows(); ++r) m_(r, j_) = o.m_(r, o.j_);
      return *this;
    }
    operator Matrix<T, R, 1>() const {
      Matrix<T, R, 1> v(m_.rows(), 1);
      for (Index r = 0; r < m_.rows(); ++r) v(r) = m_(r, j_);
      return v;
    }
    T& operator()(Index r) { return m_(r, j_); }
    T& operator[](Index r) { return m_(r, j_); }
    // `m.col(j) << x, y, z;`
    class Comma {
     public:
      Comma(ColXpr& c, T v) : c_(c), k_(0) { c_(k_++) = v; }
      Comma& operator,(T v) {
        c_(k_++) = v;
        return *this;
      }

     private:
      ColXpr& c_;
      Index k_;
    };
    Comma operator<<(T v) { return Comma(*this, v); }

   private:
    Matrix& m_;
    Index j_;
  };
  ColXpr col(Index j) { return ColXpr(*this, j); }
  Matrix<T, R, 1> col(Index j) const {
    Matrix<T, R, 1> v(rows_, 1);
    for (Index r = 0; r < rows_; ++r) v(r) = (*this)(r, j);
    return v;
  }

  // ---- `m << a, b, c, ...;` (row-major fill order, like Eigen's comma initialiser)
  class CommaInit {
   public:
    CommaInit(Matrix& m, T v) : m_(m), k_(0) { put(v); }
    CommaInit& operator,(T v) {
      put(v);
      return *this;
    }

   private:
    void put(T v) {
      const Index r = k_ / m_.cols(), c = k_ % m_.cols();
      m_(r, c) = v;
      ++k_;
    }
    Matrix& m_;
    Index k_;
  };
  CommaInit operator<<(T v) { return CommaInit(*this, v); }

  // ---- arithmetic (only what registration callers need)
  Matrix<T, C, R> transpose() const {
    Matrix<T, C, R> t(cols_, rows_);
    for (Index r = 0; r < rows_; ++r)
      for (Index c = 0; c < cols_; ++c) t(c, r) = (*this)(r, c);
    return t;
  }
  T trace() const {
    T s = T(0);
    for (Index i = 0; i < std::min(rows_, cols_); ++i) s += (*this)(i, i);
    return s;
  }
  T squaredNorm() const {
    T s = T(0);
    for (Index i = 0; i < size(); ++i) s += d_[i] * d_[i];
    return s;
  }
  T norm() const { return std::sqrt(squaredNorm()); }
  T sum() const {
    T s = T(0);
    for (Index i = 0; i < size(); ++i) s += d_[i];
    return s;
  }
  T determinant() const {
    assert(rows_ == 3 && cols_ == 3);
    const Matrix& m = *this;
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
  }
  template <typename U>
  Matrix<U, R, C> cast() const {
    Matrix<U, R, C> o(rows_, cols_);
    for (Index i = 0; i < size(); ++i) o.data()[i] = static_cast<U>(d_[i]);
    return o;
  }
  Matrix operator+(const Matrix& o) const {
    Matrix r(*this);
    for (Index i = 0; i < size(); ++i) r.d_[i] += o.d_[i];
    return r;
  }
  Matrix operator-(const Matrix& o) const {
    Matrix r(*this);
    for (Index i = 0; i < size(); ++i) r.d_[i] -= o.d_[i];
    return r;
  }
  Matrix operator*(T s) const {
    Matrix r(*this);
    for (Index i = 0; i < size(); ++i) r.d_[i] *= s;
    return r;
  }
  Matrix& operator*=(T s) {
    for (Index i = 0; i < size(); ++i) d_[i] *= s;
    return *this;
  }
  friend Matrix operator*(T s, const Matrix& m) { return m * s; }
  template <int C2>
  Matrix<T, R, C2> operator*(const Matrix<T, C, C2>& o) const {
    assert(cols_ == o.rows());
    Matrix<T, R, C2> r(rows_, o.cols());
    for (Index i = 0; i < rows_; ++i)
      for (Index j = 0; j < o.cols(); ++j) {
        T s = T(0);
        for (Index k = 0; k < cols_; ++k) s += (*this)(i, k) * o(k, j);
        r(i, j) = s;
      }
    return r;
  }
  bool operator==(const Matrix& o) const {
    return rows_ == o.rows_ && cols_ == o.cols_ && std::equal(d_.get(), d_.get() + size(), o.d_.get());
  }

 private:
  void alloc() { d_.reset(size() > 0 ? new T[static_cast<size_t>(size())]() : nullptr); }
  Index rows_, cols_;
  std::unique_ptr<T[]> d_;
};

using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;
using Vector3d = Matrix<double, 3, 1>;
using Vector3f = Matrix<float, 3, 1>;
using VectorXd = Matrix<double, Dynamic, 1>;
using RowVectorXd = Matrix<double, 1, Dynamic>;
using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using MatrixXi = Matrix<int, Dynamic, Dynamic>;

}  // namespace Eigen
