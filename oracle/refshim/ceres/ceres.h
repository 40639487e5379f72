// refshim — TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// The subset of the Ceres Solver 1.14 API that src/laser_odometry.cc and include/liodom/factors.hpp
// use, so that they compile UNMODIFIED: Jet automatic differentiation, AutoDiffCostFunction, HuberLoss,
// EigenQuaternionParameterization, Problem, and Solve() as the trust-region Levenberg-Marquardt loop
// with the options the reference sets (DENSE_QR, max_num_iterations = 4) and Ceres' defaults for the
// rest.  Ceres is a third-party dependency that is neither vendored in the
// reference nor installed here; its published algorithm is restated.  Written independently of
// oracle/liodom_oracle.cc: this one is generic over parameter / residual blocks and differentiates the
// reference's OWN functor (Point2LineFactor::operator()) through Jets.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "../refshim_eigen.h"

namespace ceres {

// ---- jet.h ----------------------------------------------------------------------------------------
template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a() { for (int i = 0; i < N; ++i) v[i] = T(); }
  Jet(const T& value) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(); }   // NOLINT (implicit, as in Ceres)
  Jet(const T& value, int k) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(); v[k] = T(1.0); }
};
#define REFSHIM_JET template <typename T, int N> inline
REFSHIM_JET Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
REFSHIM_JET Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
REFSHIM_JET Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
REFSHIM_JET Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a + s; return h; }
REFSHIM_JET Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a = f.a + s; return h; }
REFSHIM_JET Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
REFSHIM_JET Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a - s; return h; }
REFSHIM_JET Jet<T, N> operator-(T s, const Jet<T, N>& f) { Jet<T, N> h; h.a = s - f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
REFSHIM_JET Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
REFSHIM_JET Jet<T, N> operator*(const Jet<T, N>& f, T s) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
REFSHIM_JET Jet<T, N> operator*(T s, const Jet<T, N>& f) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
REFSHIM_JET Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  // jet.h: g_a_inverse = 1 / g.a; f_a_by_g_a = f.a * g_a_inverse; v = (f.v - f_a_by_g_a * g.v) * g_a_inverse
  Jet<T, N> h;
  const T g_a_inverse = T(1.0) / g.a;
  const T f_a_by_g_a = f.a * g_a_inverse;
  h.a = f_a_by_g_a;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
  return h;
}
REFSHIM_JET Jet<T, N> operator/(const Jet<T, N>& f, T s) { const T inv = T(1.0) / s; Jet<T, N> h; h.a = f.a * inv; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * inv; return h; }
REFSHIM_JET Jet<T, N> operator/(T s, const Jet<T, N>& g) { const T minus_s_g_a_inverse2 = -s / (g.a * g.a); Jet<T, N> h; h.a = s / g.a; for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * minus_s_g_a_inverse2; return h; }
#define REFSHIM_JET_CMP(op)                                                                         \
  REFSHIM_JET bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; }      \
  REFSHIM_JET bool operator op(const T& s, const Jet<T, N>& g) { return s op g.a; }                \
  REFSHIM_JET bool operator op(const Jet<T, N>& f, const T& s) { return f.a op s; }
REFSHIM_JET_CMP(<) REFSHIM_JET_CMP(<=) REFSHIM_JET_CMP(>) REFSHIM_JET_CMP(>=) REFSHIM_JET_CMP(==) REFSHIM_JET_CMP(!=)
#undef REFSHIM_JET_CMP

using std::abs;
using std::acos;
using std::cos;
using std::exp;
using std::pow;
using std::sin;
using std::sqrt;
REFSHIM_JET Jet<T, N> abs(const Jet<T, N>& f) { return f.a < T(0.0) ? -f : f; }
REFSHIM_JET Jet<T, N> sqrt(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sqrt(f.a); const T two_a_inverse = T(1.0) / (T(2.0) * h.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse; return h; }
REFSHIM_JET Jet<T, N> sin(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sin(f.a); const T c = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
REFSHIM_JET Jet<T, N> cos(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::cos(f.a); const T s = -std::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
REFSHIM_JET Jet<T, N> acos(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::acos(f.a); const T tmp = -T(1.0) / std::sqrt(T(1.0) - f.a * f.a); for (int i = 0; i < N; ++i) h.v[i] = tmp * f.v[i]; return h; }
REFSHIM_JET Jet<T, N> exp(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::exp(f.a); for (int i = 0; i < N; ++i) h.v[i] = h.a * f.v[i]; return h; }
#undef REFSHIM_JET

// ---- loss_function.h ------------------------------------------------------------------------------
class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};
class HuberLoss : public LossFunction {
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  void Evaluate(double s, double rho[3]) const override {
    if (s > b_) {   // outlier region: rho(s) = 2 a sqrt(s) - b
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a_ * r - b_;
      rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else {
      rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
    }
  }
 private:
  const double a_, b_;
};

// ---- local_parameterization.h ---------------------------------------------------------------------
class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;   // GlobalSize x LocalSize, row-major
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};
// Quaternion stored (x, y, z, w); Plus(x, delta) = [sin|d|/|d| d ; cos|d|] (x) x  (Eigen product order)
class EigenQuaternionParameterization : public LocalParameterization {
 public:
  bool Plus(const double* x_ptr, const double* delta, double* out) const override {
    const double norm_delta = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    if (norm_delta > 0.0) {
      const double sin_delta_by_delta = std::sin(norm_delta) / norm_delta;
      const Eigen::Quaterniond dq(std::cos(norm_delta), sin_delta_by_delta * delta[0], sin_delta_by_delta * delta[1], sin_delta_by_delta * delta[2]);
      const Eigen::Quaterniond x(x_ptr[3], x_ptr[0], x_ptr[1], x_ptr[2]);
      // Eigen quat product a * b
      const double aw = dq.w(), ax = dq.x(), ay = dq.y(), az = dq.z(), bw = x.w(), bx = x.x(), by = x.y(), bz = x.z();
      out[3] = aw * bw - ax * bx - ay * by - az * bz;
      out[0] = aw * bx + ax * bw + ay * bz - az * by;
      out[1] = aw * by + ay * bw + az * bx - ax * bz;
      out[2] = aw * bz + az * bw + ax * by - ay * bx;
    } else {
      for (int k = 0; k < 4; ++k) out[k] = x_ptr[k];
    }
    return true;
  }
  bool ComputeJacobian(const double* x, double* j) const override {
    j[0] = x[3];  j[1] = x[2];   j[2] = -x[1];
    j[3] = -x[2]; j[4] = x[3];   j[5] = x[0];
    j[6] = x[1];  j[7] = -x[0];  j[8] = x[3];
    j[9] = -x[0]; j[10] = -x[1]; j[11] = -x[2];
    return true;
  }
  int GlobalSize() const override { return 4; }
  int LocalSize() const override { return 3; }
};

// ---- cost_function.h / autodiff_cost_function.h ---------------------------------------------------
class CostFunction {
 public:
  virtual ~CostFunction() {}
  // jacobians[i]: num_residuals x block_size(i), row-major, or NULL
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  int num_residuals() const { return num_residuals_; }
  const std::vector<int>& parameter_block_sizes() const { return sizes_; }
 protected:
  int num_residuals_ = 0;
  std::vector<int> sizes_;
};

template <typename Functor, int kNumResiduals, int N0, int N1>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : functor_(f) { num_residuals_ = kNumResiduals; sizes_ = {N0, N1}; }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    if (!jacobians) return (*functor_)(parameters[0], parameters[1], residuals);
    typedef Jet<double, N0 + N1> JetT;
    JetT x0[N0], x1[N1], out[kNumResiduals];
    for (int k = 0; k < N0; ++k) x0[k] = JetT(parameters[0][k], k);
    for (int k = 0; k < N1; ++k) x1[k] = JetT(parameters[1][k], N0 + k);
    if (!(*functor_)(x0, x1, out)) return false;
    for (int r = 0; r < kNumResiduals; ++r) {
      residuals[r] = out[r].a;
      if (jacobians[0]) for (int k = 0; k < N0; ++k) jacobians[0][r * N0 + k] = out[r].v[k];
      if (jacobians[1]) for (int k = 0; k < N1; ++k) jacobians[1][r * N1 + k] = out[r].v[N0 + k];
    }
    return true;
  }
 private:
  std::unique_ptr<Functor> functor_;
};

// ---- problem.h --------------------------------------------------------------------------------------
class Problem {
 public:
  struct Options {};
  Problem() {}
  explicit Problem(const Options&) {}
  Problem(const Problem&) = delete;
  ~Problem() {
    std::set<LossFunction*> losses;
    for (auto& rb : residual_blocks_) { delete rb.cost; if (rb.loss) losses.insert(rb.loss); }
    for (LossFunction* l : losses) delete l;
    std::set<LocalParameterization*> ps;
    for (auto& pb : parameter_blocks_) if (pb.param) ps.insert(pb.param);
    for (LocalParameterization* p : ps) delete p;
  }
  void AddParameterBlock(double* values, int size, LocalParameterization* p = nullptr) {
    for (auto& pb : parameter_blocks_) if (pb.values == values) { if (p) pb.param = p; return; }
    parameter_blocks_.push_back(ParameterBlock{values, size, p});
  }
  void* AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, double* x1) {
    double* xs[2] = {x0, x1};
    ResidualBlock rb; rb.cost = cost; rb.loss = loss;
    for (int k = 0; k < 2; ++k) {
      int found = -1;
      for (size_t i = 0; i < parameter_blocks_.size(); ++i) if (parameter_blocks_[i].values == xs[k]) found = (int)i;
      if (found < 0) { AddParameterBlock(xs[k], cost->parameter_block_sizes()[k]); found = (int)parameter_blocks_.size() - 1; }
      rb.block[k] = found;
    }
    residual_blocks_.push_back(rb);
    return nullptr;
  }
  int NumResidualBlocks() const { return (int)residual_blocks_.size(); }

  struct ParameterBlock {
    double* values; int size; LocalParameterization* param;
    int local_size() const { return param ? param->LocalSize() : size; }
  };
  struct ResidualBlock { CostFunction* cost; LossFunction* loss; int block[2]; };
  std::vector<ParameterBlock> parameter_blocks_;
  std::vector<ResidualBlock> residual_blocks_;
};

// ---- solver.h / types.h -----------------------------------------------------------------------------
enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };

class Solver {
 public:
  struct Options {
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    int max_num_iterations = 50;
    bool minimizer_progress_to_stdout = false;
    int num_threads = 1;
    // defaults of Ceres 1.14 that the reference leaves untouched
    double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    int max_num_consecutive_invalid_steps = 5;
    bool jacobi_scaling = true;
  };
  struct Summary {
    TerminationType termination_type = NO_CONVERGENCE;
    std::string message;
    double initial_cost = 0.0, final_cost = 0.0;
    int num_successful_steps = 0, num_unsuccessful_steps = 0;
    int num_residual_blocks = 0;
    int num_iterations = 0;            // iterations of the minimizer loop (excluding iteration 0)
    int num_cost_evaluations = 0, num_jacobian_evaluations = 0;
    int refshim_termination_code = 0;  // 0 max-iter, 1 gradient, 2 parameter, 3 function tol, 4 no residuals, 5 failure
    std::string BriefReport() const { return "refshim: " + message; }
  };
};

namespace refshim_detail {

// Householder QR least squares min ||A y - b|| for a dense row-major m x n matrix (A, b destroyed).
inline bool householder_lsq(std::vector<double>& A, std::vector<double>& b, int m, int n, double* y) {
  for (int k = 0; k < n; ++k) {
    double sigma = 0.0;
    for (int i = k; i < m; ++i) sigma += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    const double norm = std::sqrt(sigma);
    if (norm == 0.0) return false;
    const double alpha = A[(size_t)k * n + k] > 0.0 ? -norm : norm;
    const double v0 = A[(size_t)k * n + k] - alpha;
    double vtv = v0 * v0;
    for (int i = k + 1; i < m; ++i) vtv += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    if (vtv == 0.0) return false;
    const double beta = 2.0 / vtv;
    for (int j = k + 1; j <= n; ++j) {   // column n = right-hand side
      double s = v0 * (j < n ? A[(size_t)k * n + j] : b[k]);
      for (int i = k + 1; i < m; ++i) s += A[(size_t)i * n + k] * (j < n ? A[(size_t)i * n + j] : b[i]);
      s *= beta;
      if (j < n) { A[(size_t)k * n + j] -= s * v0; for (int i = k + 1; i < m; ++i) A[(size_t)i * n + j] -= s * A[(size_t)i * n + k]; }
      else { b[k] -= s * v0; for (int i = k + 1; i < m; ++i) b[i] -= s * A[(size_t)i * n + k]; }
    }
    A[(size_t)k * n + k] = alpha;
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s -= A[(size_t)k * n + j] * y[j];
    if (A[(size_t)k * n + k] == 0.0) return false;
    y[k] = s / A[(size_t)k * n + k];
  }
  for (int k = 0; k < n; ++k) if (!std::isfinite(y[k])) return false;
  return true;
}

// Program view of a Problem: parameter blocks in insertion order (the reference adds q, then t).
struct Program {
  const Problem* pb;
  std::vector<int> goff, loff;   // offsets of each block in the global / local state vectors
  int nglobal = 0, nlocal = 0, nres = 0;
  explicit Program(const Problem* p) : pb(p) {
    for (auto& b : p->parameter_blocks_) { goff.push_back(nglobal); loff.push_back(nlocal); nglobal += b.size; nlocal += b.local_size(); }
    for (auto& rb : p->residual_blocks_) nres += rb.cost->num_residuals();
  }
  void plus(const double* x, const double* delta, double* out) const {
    for (size_t i = 0; i < pb->parameter_blocks_.size(); ++i) {
      const auto& b = pb->parameter_blocks_[i];
      if (b.param) b.param->Plus(x + goff[i], delta + loff[i], out + goff[i]);
      else for (int k = 0; k < b.size; ++k) out[goff[i] + k] = x[goff[i] + k] + delta[loff[i] + k];
    }
  }
  // ResidualBlock::Evaluate for every block: cost = sum 0.5 rho(s); residuals / Jacobian (nres x nlocal,
  // row-major) corrected for the loss (Corrector with rho'' <= 0: both scaled by sqrt(rho')).
  bool evaluate(const double* x, double* cost, double* residuals, double* jacobian) const {
    *cost = 0.0;
    if (jacobian) std::fill(jacobian, jacobian + (size_t)nres * nlocal, 0.0);
    int row = 0;
    std::vector<double> r, jg[2], jl;
    for (const auto& rb : pb->residual_blocks_) {
      const int nr = rb.cost->num_residuals();
      const double* params[2] = {x + goff[rb.block[0]], x + goff[rb.block[1]]};
      r.assign(nr, 0.0);
      double* jac_ptr[2] = {nullptr, nullptr};
      if (jacobian) for (int k = 0; k < 2; ++k) { jg[k].assign((size_t)nr * pb->parameter_blocks_[rb.block[k]].size, 0.0); jac_ptr[k] = jg[k].data(); }
      if (!rb.cost->Evaluate(params, r.data(), jacobian ? jac_ptr : nullptr)) return false;
      double sq = 0.0;
      for (int i = 0; i < nr; ++i) sq += r[i] * r[i];
      double rho[3] = {sq, 1.0, 0.0};
      if (rb.loss) rb.loss->Evaluate(sq, rho);
      *cost += 0.5 * rho[0];
      const double sqrt_rho1 = std::sqrt(rho[1]);
      // Corrector: alpha = 0 when sq == 0 or rho'' <= 0 (always for Huber)
      double alpha_sq_norm = 0.0, residual_scaling = sqrt_rho1;
      if (sq != 0.0 && rho[2] > 0.0) {
        const double D = 1.0 + 2.0 * sq * rho[2] / rho[1];
        const double alpha = 1.0 - std::sqrt(D);
        residual_scaling = sqrt_rho1 / (1.0 - alpha);
        alpha_sq_norm = alpha / sq;
      }
      if (jacobian) {
        for (int k = 0; k < 2; ++k) {
          const auto& b = pb->parameter_blocks_[rb.block[k]];
          const int gs = b.size, ls = b.local_size();
          jl.assign((size_t)nr * ls, 0.0);
          if (b.param) {
            std::vector<double> pj((size_t)gs * ls);
            b.param->ComputeJacobian(params[k], pj.data());
            for (int i = 0; i < nr; ++i)
              for (int j = 0; j < ls; ++j) { double s = 0.0; for (int c = 0; c < gs; ++c) s += jg[k][(size_t)i * gs + c] * pj[(size_t)c * ls + j]; jl[(size_t)i * ls + j] = s; }
          } else jl = jg[k];
          if (rb.loss) {
            if (alpha_sq_norm == 0.0) for (double& v : jl) v *= sqrt_rho1;
            else for (int j = 0; j < ls; ++j) {
              double rtj = 0.0; for (int i = 0; i < nr; ++i) rtj += r[i] * jl[(size_t)i * ls + j];
              for (int i = 0; i < nr; ++i) jl[(size_t)i * ls + j] = sqrt_rho1 * (jl[(size_t)i * ls + j] - alpha_sq_norm * r[i] * rtj);
            }
          }
          for (int i = 0; i < nr; ++i) for (int j = 0; j < ls; ++j) jacobian[(size_t)(row + i) * nlocal + loff[rb.block[k]] + j] = jl[(size_t)i * ls + j];
        }
      }
      if (residuals) for (int i = 0; i < nr; ++i) residuals[row + i] = (rb.loss ? residual_scaling : 1.0) * r[i];
      row += nr;
    }
    return true;
  }
};

}  // namespace refshim_detail

// TrustRegionMinimizer + LevenbergMarquardtStrategy + DenseQRSolver (Ceres 1.14, trust_region_minimizer.cc)
inline void Solve(const Solver::Options& opt, Problem* problem, Solver::Summary* summary) {
  using refshim_detail::Program;
  Solver::Summary S;
  Program prog(problem);
  S.num_residual_blocks = problem->NumResidualBlocks();
  if (prog.nres == 0) {   // preprocessor: nothing to optimise, parameters untouched
    S.termination_type = CONVERGENCE; S.refshim_termination_code = 4;
    S.message = "Function tolerance reached. No non-constant parameter blocks found.";
    if (summary) *summary = S;
    return;
  }
  const int n = prog.nlocal, ng = prog.nglobal, m = prog.nres;
  std::vector<double> x(ng), xc(ng), xp(ng);
  for (size_t i = 0; i < problem->parameter_blocks_.size(); ++i)
    std::memcpy(x.data() + prog.goff[i], problem->parameter_blocks_[i].values, sizeof(double) * problem->parameter_blocks_[i].size);
  std::vector<double> r(m), J((size_t)m * n), scale(n, 1.0), diagonal(n), g(n), step(n), delta(n), neg_g(n);
  double x_cost = 0.0, radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int num_consecutive_invalid = 0;
  auto norm2 = [](const std::vector<double>& v) { double s = 0; for (double e : v) s += e * e; return std::sqrt(s); };

  // EvaluateGradientAndJacobian: J is column-scaled in place after iteration 0 computed the scaling
  auto evaluate_gradient_and_jacobian = [&](bool first) -> double {
    prog.evaluate(x.data(), &x_cost, r.data(), J.data());
    S.num_jacobian_evaluations++;
    for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < m; ++i) s += J[(size_t)i * n + j] * r[i]; g[j] = s; }
    if (first && opt.jacobi_scaling)
      for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < m; ++i) s += J[(size_t)i * n + j] * J[(size_t)i * n + j]; scale[j] = 1.0 / (1.0 + std::sqrt(s)); }
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) J[(size_t)i * n + j] *= scale[j];
    for (int j = 0; j < n; ++j) neg_g[j] = -g[j];
    prog.plus(x.data(), neg_g.data(), xp.data());
    double mx = 0; for (int k = 0; k < ng; ++k) mx = std::max(mx, std::fabs(xp[k] - x[k]));
    return mx;   // gradient_max_norm = ||x - Plus(x, -g)||_inf
  };

  double gradient_max_norm = evaluate_gradient_and_jacobian(true);
  S.initial_cost = x_cost;
  double x_norm = norm2(x);
  bool step_is_successful = true;
  int iteration = 0, code = 0;
  for (;;) {
    if (iteration >= opt.max_num_iterations) { code = 0; S.termination_type = NO_CONVERGENCE; S.message = "Maximum number of iterations reached."; break; }
    if (step_is_successful && gradient_max_norm <= opt.gradient_tolerance) { code = 1; S.termination_type = CONVERGENCE; S.message = "Gradient tolerance reached."; break; }
    if (radius < opt.min_trust_region_radius) { code = 5; S.termination_type = CONVERGENCE; S.message = "Minimum trust region radius reached."; break; }
    iteration++;
    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal)
      for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < m; ++i) s += J[(size_t)i * n + j] * J[(size_t)i * n + j]; diagonal[j] = std::min(std::max(s, opt.min_lm_diagonal), opt.max_lm_diagonal); }
    std::vector<double> A((size_t)(m + n) * n, 0.0), rhs(m + n, 0.0), y(n, 0.0);
    std::memcpy(A.data(), J.data(), sizeof(double) * (size_t)m * n);
    for (int j = 0; j < n; ++j) A[(size_t)(m + j) * n + j] = std::sqrt(diagonal[j] / radius);
    std::memcpy(rhs.data(), r.data(), sizeof(double) * m);
    const bool solved = refshim_detail::householder_lsq(A, rhs, m + n, n, y.data());
    reuse_diagonal = true;
    double model_cost_change = 0.0;
    bool step_valid = false;
    if (solved) {
      for (int j = 0; j < n; ++j) step[j] = -y[j];
      double acc = 0.0;   // -(J step)'(r + J step / 2)
      for (int i = 0; i < m; ++i) { double mr = 0; for (int j = 0; j < n; ++j) mr += J[(size_t)i * n + j] * step[j]; acc += mr * (r[i] + mr / 2.0); }
      model_cost_change = -acc;
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) {
      if (++num_consecutive_invalid >= opt.max_num_consecutive_invalid_steps) { code = 5; S.termination_type = FAILURE; S.message = "Too many consecutive invalid steps."; break; }
      radius *= 0.5; step_is_successful = false; S.num_unsuccessful_steps++;
      continue;
    }
    num_consecutive_invalid = 0;
    for (int j = 0; j < n; ++j) delta[j] = step[j] * scale[j];
    prog.plus(x.data(), delta.data(), xc.data());
    double candidate_cost = 0.0;
    if (!prog.evaluate(xc.data(), &candidate_cost, nullptr, nullptr) || !std::isfinite(candidate_cost)) candidate_cost = x_cost;
    S.num_cost_evaluations++;
    double step_norm = 0.0; for (int k = 0; k < ng; ++k) step_norm += (x[k] - xc[k]) * (x[k] - xc[k]);
    step_norm = std::sqrt(step_norm);
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { code = 2; S.termination_type = CONVERGENCE; S.message = "Parameter tolerance reached."; break; }
    const double cost_change = x_cost - candidate_cost;
    if (std::fabs(cost_change) <= opt.function_tolerance * x_cost) { code = 3; S.termination_type = CONVERGENCE; S.message = "Function tolerance reached."; break; }
    const double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > opt.min_relative_decrease) {
      x = xc;
      x_norm = norm2(x);
      gradient_max_norm = evaluate_gradient_and_jacobian(false);
      step_is_successful = true; S.num_successful_steps++;
      radius = std::min(opt.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3)));
      decrease_factor = 2.0; reuse_diagonal = false;
    } else {
      step_is_successful = false; S.num_unsuccessful_steps++;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
  }
  S.num_iterations = iteration; S.final_cost = x_cost; S.refshim_termination_code = code;
  for (size_t i = 0; i < problem->parameter_blocks_.size(); ++i)
    std::memcpy(problem->parameter_blocks_[i].values, x.data() + prog.goff[i], sizeof(double) * problem->parameter_blocks_[i].size);
  if (summary) *summary = S;
}

}  // namespace ceres

namespace Eigen {
template <typename T, int N>
struct NumTraits<ceres::Jet<T, N> > {
  static ceres::Jet<T, N> epsilon() { return ceres::Jet<T, N>(std::numeric_limits<T>::epsilon()); }
};
}  // namespace Eigen
