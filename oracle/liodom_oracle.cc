// liodom oracle — TEST INFRASTRUCTURE ONLY (see liodom_oracle.h for scope and the
// "parity unpinned" statement).  CPU restatement of the LiODOM hot path, written from
// the reference's behaviour; every function cites the reference file:line it follows.
//
// Build flags mirror the reference's CMakeLists.txt:13 (-O3 -g, C++17, OpenMP) plus
// -ffp-contract=off and no -march=native so that no FMA contraction can change the
// float/double rounding sequence of the curvature and distance arithmetic.
#include "liodom_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <queue>
#include <unordered_map>
#include <vector>
#include <omp.h>
#include <unistd.h>

namespace {

using Clock = std::chrono::steady_clock;
inline double us_since(Clock::time_point t0) {
  return std::chrono::duration<double, std::micro>(Clock::now() - t0).count();
}

struct P4 { float x, y, z, i; };

// ---------------------------------------------------------------------------------
// A1  FeatureExtractor::isValidPoint  (src/feature_extractor.cc:84-102)
// ---------------------------------------------------------------------------------
inline bool is_valid_point(const OrcParams& p, double x, double y, double z, double* dist) {
  bool valid = true;
  if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) valid = false;
  *dist = std::sqrt(x * x + y * y);
  if (*dist > p.max_range || *dist < p.min_range) valid = false;
  return valid;
}

// A2 ring id for Velodyne models (src/feature_extractor.cc:126-151). -1: rejected.
inline int velodyne_ring(const OrcParams& p, double z, double distance) {
  int scan_id = -1;
  double angle = atan(z / distance) * 180 / M_PI;
  if (p.scan_lines == 64) {
    if (angle >= -8.83)
      scan_id = int((2 - angle) * 3.0 + 0.5);
    else
      scan_id = p.scan_lines / 2 + int((-8.83 - angle) * 2.0 + 0.5);
    if (angle > 2 || angle < -24.33 || scan_id > 63 || scan_id < 0) return -1;
  } else if (p.scan_lines == 32) {
    scan_id = int((angle + 92.0 / 3.0) * 3.0 / 4.0);
    if (scan_id > (p.scan_lines - 1) || scan_id < 0) return -1;
  } else if (p.scan_lines == 16) {
    scan_id = int((angle + 15) / 2 + 0.5);
    if (scan_id > (p.scan_lines - 1) || scan_id < 0) return -1;
  } else {
    return -1;  // ROS_ERROR_ONCE("Invalid scan lines"), scan_id stays -1
  }
  return scan_id;
}

// ---------------------------------------------------------------------------------
// A3/A4 helpers
// ---------------------------------------------------------------------------------
struct SmoothnessItem {  // include/liodom/feature_extractor.h:44-60
  int point_index;
  double smoothness;
  bool operator<(const SmoothnessItem& s) const { return smoothness > s.smoothness; }
};

struct Extractor {
  std::vector<uint8_t> picked;  // bool picked_[400000] (feature_extractor.h:75), persistent
  Extractor() : picked(400000, 0) {}
};

inline int extractor_threads(const OrcParams& p) {
  if (p.omp_threads > 0) return p.omp_threads;
  int n = 2;  // src/feature_extractor.cc:29-34
  int m = omp_get_max_threads() - 5;
  if (m > 1) n = m;
  return n;
}

// src/feature_extractor.cc:256-313
void extract_region(const OrcParams& p, Extractor& ex, const P4* pc, std::vector<SmoothnessItem>& smooths,
                    int sort_mode, int ring, std::vector<P4>& out, std::vector<int>& out_ring,
                    std::vector<int>& out_idx) {
  if (sort_mode == 0) {
    std::sort(smooths.begin(), smooths.end());
  } else {
    std::sort(smooths.begin(), smooths.end(), [](const SmoothnessItem& a, const SmoothnessItem& b) {
      if (a.smoothness != b.smoothness) return a.smoothness > b.smoothness;
      return a.point_index < b.point_index;
    });
  }
  int picked_edges = 0;
  for (size_t i = 0; i < smooths.size(); i++) {
    int point_index = smooths[i].point_index;
    if (!ex.picked[point_index]) {
      if (smooths[i].smoothness < 0.1 || picked_edges > p.edges_per_region) break;
      out.push_back(pc[point_index]);
      out_ring.push_back(ring);
      out_idx.push_back(point_index);
      picked_edges++;
      ex.picked[point_index] = 1;
      for (int l = 1; l <= 5; l++) {
        double diff_x = pc[point_index + l].x - pc[point_index + l - 1].x;
        double diff_y = pc[point_index + l].y - pc[point_index + l - 1].y;
        double diff_z = pc[point_index + l].z - pc[point_index + l - 1].z;
        if (diff_x * diff_x + diff_y * diff_y + diff_z * diff_z > 0.05) break;
        ex.picked[point_index + l] = 1;
      }
      for (int l = -1; l >= -5; l--) {
        double diff_x = pc[point_index + l].x - pc[point_index + l + 1].x;
        double diff_y = pc[point_index + l].y - pc[point_index + l + 1].y;
        double diff_z = pc[point_index + l].z - pc[point_index + l + 1].z;
        if (diff_x * diff_x + diff_y * diff_y + diff_z * diff_z > 0.05) break;
        ex.picked[point_index + l] = 1;
      }
    }
  }
}

// src/feature_extractor.cc:181-254
void extract_features(const OrcParams& p, Extractor& ex, const P4* rings, const int32_t* off,
                      int sort_mode, double* keys, std::vector<P4>& out, std::vector<int>& out_ring,
                      std::vector<int>& out_idx) {
  const size_t min_points_per_scan = (size_t)(p.scan_regions * p.edges_per_region + 10);  // params.cc:63
  const int nthreads = extractor_threads(p);
  for (int i = 0; i < p.scan_lines; i++) {
    const P4* pts = rings + off[i];
    const size_t n = (size_t)(off[i + 1] - off[i]);
    if (n < min_points_per_scan) continue;
    std::vector<SmoothnessItem> smooths_aux(n, SmoothnessItem{-1, -1.0});
#pragma omp parallel for num_threads(nthreads)
    for (size_t j = 5; j < n - 5; j++) {
      // float arithmetic, left to right, 10*x a float multiply; widened on assignment.
      double diff_x = pts[j - 5].x + pts[j - 4].x + pts[j - 3].x + pts[j - 2].x + pts[j - 1].x -
                      10 * pts[j].x + pts[j + 1].x + pts[j + 2].x + pts[j + 3].x + pts[j + 4].x +
                      pts[j + 5].x;
      double diff_y = pts[j - 5].y + pts[j - 4].y + pts[j - 3].y + pts[j - 2].y + pts[j - 1].y -
                      10 * pts[j].y + pts[j + 1].y + pts[j + 2].y + pts[j + 3].y + pts[j + 4].y +
                      pts[j + 5].y;
      double diff_z = pts[j - 5].z + pts[j - 4].z + pts[j - 3].z + pts[j - 2].z + pts[j - 1].z -
                      10 * pts[j].z + pts[j + 1].z + pts[j + 2].z + pts[j + 3].z + pts[j + 4].z +
                      pts[j + 5].z;
      SmoothnessItem item{(int)j, diff_x * diff_x + diff_y * diff_y + diff_z * diff_z};
      ex.picked[j] = 0;
      smooths_aux[j] = item;
    }
    if (keys)
      for (size_t j = 5; j < n - 5; j++) keys[off[i] + j] = smooths_aux[j].smoothness;
    std::vector<SmoothnessItem> smooths(smooths_aux.begin() + 5, smooths_aux.end() - 5);
    int total_points = (int)n - 10;
    int sector_length = (int)(total_points / p.scan_regions);
    for (int j = 0; j < p.scan_regions; j++) {
      int region_start = sector_length * j;
      int region_end = sector_length * (j + 1);
      if (j == p.scan_regions - 1) region_end = total_points;
      std::vector<SmoothnessItem> smooths_sub(smooths.begin() + region_start, smooths.begin() + region_end);
      extract_region(p, ex, pts, smooths_sub, sort_mode, i, out, out_ring, out_idx);
    }
  }
}

// ---------------------------------------------------------------------------------
// A.1 pcl::transformPointCloud, double matrix, generic path
// ---------------------------------------------------------------------------------
inline P4 transform_point(const double* T, const P4& s) {
  const double px = s.x, py = s.y, pz = s.z;
  P4 o;
  o.x = static_cast<float>(T[0] * px + T[1] * py + T[2] * pz + T[3]);
  o.y = static_cast<float>(T[4] * px + T[5] * py + T[6] * pz + T[7]);
  o.z = static_cast<float>(T[8] * px + T[9] * py + T[10] * pz + T[11]);
  o.i = s.i;
  return o;
}

// ---------------------------------------------------------------------------------
// A.2 exact k-NN (k=5) under flann::L2_Simple<float>
// ---------------------------------------------------------------------------------
inline float l2_simple(const P4& a, const P4& b) {
  float result = 0.f, diff;
  diff = a.x - b.x; result += diff * diff;
  diff = a.y - b.y; result += diff * diff;
  diff = a.z - b.z; result += diff * diff;
  return result;
}

struct Knn5 {  // ascending (d2, idx); keeps a 6th entry to flag ties at the boundary
  float d[6]; int id[6]; int n;
  Knn5() : n(0) { for (int k = 0; k < 6; ++k) { d[k] = std::numeric_limits<float>::infinity(); id[k] = -1; } }
  inline float worst() const { return d[5]; }
  inline void push(float dist, int idx) {
    if (dist > d[5] || (dist == d[5] && idx > id[5] && id[5] >= 0)) return;
    int k = 5;
    while (k > 0 && (d[k - 1] > dist || (d[k - 1] == dist && (id[k - 1] > idx || id[k - 1] < 0)))) {
      d[k] = d[k - 1]; id[k] = id[k - 1]; --k;
    }
    d[k] = dist; id[k] = idx;
  }
};

class KdTree {  // single kd-tree, leaf size 15, points reordered (FLANN KDTreeSingleIndex shape)
 public:
  KdTree(const P4* pts, int n) : src_(pts), n_(n) {
    ind_.resize(n);
    std::iota(ind_.begin(), ind_.end(), 0);
    // PCL skips non-finite points when building the index (kdtree_flann.hpp convertCloudToArray).
    ind_.erase(std::remove_if(ind_.begin(), ind_.end(), [&](int i) {
      return !std::isfinite(pts[i].x) || !std::isfinite(pts[i].y) || !std::isfinite(pts[i].z); }), ind_.end());
    n_ = (int)ind_.size();
    nodes_.reserve(2 * n_ / 8 + 4);
    if (n_ > 0) {
      float lo[3], hi[3];
      bbox(0, n_, lo, hi);
      for (int k = 0; k < 3; ++k) { root_lo_[k] = lo[k]; root_hi_[k] = hi[k]; }
      build(0, n_, lo, hi);
      data_.resize(n_);
      for (int i = 0; i < n_; ++i) data_[i] = src_[ind_[i]];
    }
  }
  void knn(const P4& q, Knn5& res) const {
    if (n_ == 0) return;
    double dists[3] = {0, 0, 0};
    double dsq = 0;
    const float qv[3] = {q.x, q.y, q.z};
    for (int k = 0; k < 3; ++k) {
      if (qv[k] < root_lo_[k]) { double d = (double)qv[k] - root_lo_[k]; dists[k] = d * d; dsq += dists[k]; }
      if (qv[k] > root_hi_[k]) { double d = (double)qv[k] - root_hi_[k]; dists[k] = d * d; dsq += dists[k]; }
    }
    search(0, q, qv, dsq, dists, res);
  }

 private:
  struct Node { int left, right, lo, hi, dim; float divlow, divhigh; };
  void bbox(int b, int e, float* lo, float* hi) const {
    for (int k = 0; k < 3; ++k) { lo[k] = std::numeric_limits<float>::infinity(); hi[k] = -lo[k]; }
    for (int i = b; i < e; ++i) {
      const P4& p = src_[ind_[i]];
      const float v[3] = {p.x, p.y, p.z};
      for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], v[k]); hi[k] = std::max(hi[k], v[k]); }
    }
  }
  inline float coord(int i, int dim) const { const P4& p = src_[ind_[i]]; return dim == 0 ? p.x : (dim == 1 ? p.y : p.z); }
  int build(int b, int e, float* lo, float* hi) {
    int me = (int)nodes_.size();
    nodes_.push_back(Node{-1, -1, b, e, 0, 0, 0});
    if (e - b <= 15) return me;
    int dim = 0; float span = hi[0] - lo[0];
    for (int k = 1; k < 3; ++k) if (hi[k] - lo[k] > span) { span = hi[k] - lo[k]; dim = k; }
    // median split on the widest dimension keeps the tree balanced for any input
    int mid = (b + e) / 2;
    std::nth_element(ind_.begin() + b, ind_.begin() + mid, ind_.begin() + e,
                     [&](int a, int c) { float va = comp(a, dim), vc = comp(c, dim); return va < vc || (va == vc && a < c); });
    float cut = coord(mid, dim);
    float llo[3], lhi[3], rlo[3], rhi[3];
    bbox(b, mid, llo, lhi); bbox(mid, e, rlo, rhi);
    int l = build(b, mid, llo, lhi);
    int r = build(mid, e, rlo, rhi);
    nodes_[me].left = l; nodes_[me].right = r; nodes_[me].dim = dim;
    nodes_[me].divlow = lhi[dim]; nodes_[me].divhigh = rlo[dim];
    (void)cut;
    return me;
  }
  inline float comp(int idx, int dim) const { const P4& p = src_[idx]; return dim == 0 ? p.x : (dim == 1 ? p.y : p.z); }
  void search(int ni, const P4& q, const float* qv, double mindsq, double* dists, Knn5& res) const {
    const Node& nd = nodes_[ni];
    if (nd.left < 0) {
      for (int i = nd.lo; i < nd.hi; ++i) res.push(l2_simple(q, data_[i]), ind_[i]);
      return;
    }
    int dim = nd.dim;
    double val = qv[dim];
    double diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
    int best, other; double cut;
    if (diff1 + diff2 < 0) { best = nd.left; other = nd.right; cut = diff2 * diff2; }
    else { best = nd.right; other = nd.left; cut = diff1 * diff1; }
    search(best, q, qv, mindsq, dists, res);
    double saved = dists[dim];
    double nm = mindsq + cut - saved;
    // Lower bound in exact arithmetic; prune only when it clearly exceeds the worst
    // float distance kept (1e-6 slack >> float rounding of l2_simple).
    if (nm * (1.0 - 1e-6) <= (double)res.worst()) {
      dists[dim] = cut;
      search(other, q, qv, nm, dists, res);
      dists[dim] = saved;
    }
  }
  const P4* src_; int n_;
  std::vector<int> ind_;
  std::vector<P4> data_;
  std::vector<Node> nodes_;
  float root_lo_[3], root_hi_[3];
};

inline bool knn_tie(const Knn5& r) {
  for (int k = 0; k < 5; ++k)
    if (r.id[k] >= 0 && r.id[k + 1] >= 0 && r.d[k] == r.d[k + 1]) return true;
  return false;
}

// ---------------------------------------------------------------------------------
// A.6 symmetric 3x3 eigenvalues (stands in for Eigen::SelfAdjointEigenSolver):
// cyclic Jacobi in double using only + - * / sqrt, so the GPU can replay it bit for bit.
// ---------------------------------------------------------------------------------
void sym3_eigenvalues(const double* A, double* w) {
  double a00 = A[0], a01 = A[1], a02 = A[2], a11 = A[4], a12 = A[5], a22 = A[8];
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = a01 * a01 + a02 * a02 + a12 * a12;
    double diag = a00 * a00 + a11 * a11 + a22 * a22;
    if (off <= 1e-32 * diag || off == 0.0) break;
    // rotate (0,1)
    if (a01 != 0.0) {
      double theta = (a11 - a00) / (2.0 * a01);
      double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      double n00 = a00 - t * a01, n11 = a11 + t * a01;
      double n02 = c * a02 - s * a12, n12 = s * a02 + c * a12;
      a00 = n00; a11 = n11; a01 = 0.0; a02 = n02; a12 = n12;
    }
    // rotate (0,2)
    if (a02 != 0.0) {
      double theta = (a22 - a00) / (2.0 * a02);
      double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      double n00 = a00 - t * a02, n22 = a22 + t * a02;
      double n01 = c * a01 - s * a12, n12 = s * a01 + c * a12;
      a00 = n00; a22 = n22; a02 = 0.0; a01 = n01; a12 = n12;
    }
    // rotate (1,2)
    if (a12 != 0.0) {
      double theta = (a22 - a11) / (2.0 * a12);
      double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      double n11 = a11 - t * a12, n22 = a22 + t * a12;
      double n01 = c * a01 - s * a02, n02 = s * a01 + c * a02;
      a11 = n11; a22 = n22; a12 = 0.0; a01 = n01; a02 = n02;
    }
  }
  double e0 = a00, e1 = a11, e2 = a22, tmp;
  if (e0 > e1) { tmp = e0; e0 = e1; e1 = tmp; }
  if (e1 > e2) { tmp = e1; e1 = e2; e2 = tmp; }
  if (e0 > e1) { tmp = e0; e0 = e1; e1 = tmp; }
  w[0] = e0; w[1] = e1; w[2] = e2;
}

// src/laser_odometry.cc:325-344: centroid, scatter, eigen gate. nn: 5 neighbours in kNN order.
inline bool line_gate(const P4* nn, double* eig) {
  double cx = 0, cy = 0, cz = 0;
  for (int j = 0; j < 5; j++) { cx = cx + (double)nn[j].x; cy = cy + (double)nn[j].y; cz = cz + (double)nn[j].z; }
  cx = cx / 5.0; cy = cy / 5.0; cz = cz / 5.0;
  double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 5; j++) {
    double dx = (double)nn[j].x - cx, dy = (double)nn[j].y - cy, dz = (double)nn[j].z - cz;
    C[0] = C[0] + dx * dx; C[1] = C[1] + dx * dy; C[2] = C[2] + dx * dz;
    C[4] = C[4] + dy * dy; C[5] = C[5] + dy * dz; C[8] = C[8] + dz * dz;
  }
  C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
  sym3_eigenvalues(C, eig);
  return eig[2] > 3 * eig[1];
}

// ---------------------------------------------------------------------------------
// A9  Point2LineFactor through forward-mode dual numbers (what ceres::AutoDiffCostFunction
// evaluates), include/liodom/factors.hpp:71-105.
// ---------------------------------------------------------------------------------
template <int N> struct Jet {
  double a; double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  explicit Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& g) { Jet<N> h; h.a = s - g.a; for (int i = 0; i < N; ++i) h.v[i] = -g.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double fg = f.a * gi; h.a = fg;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * gi;
  return h;
}
template <int N> inline Jet<N> jsqrt(const Jet<N>& f) { Jet<N> h; h.a = std::sqrt(f.a); const double t = 1.0 / (2.0 * h.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * t; return h; }
template <int N> inline Jet<N> jsin(const Jet<N>& f) { Jet<N> h; h.a = std::sin(f.a); const double c = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <int N> inline Jet<N> jacos(const Jet<N>& f) { Jet<N> h; h.a = std::acos(f.a); const double t = -1.0 / std::sqrt(1.0 - f.a * f.a); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
template <int N> inline Jet<N> jabs(const Jet<N>& f) { return f.a < 0.0 ? -f : f; }
inline double jsqrt(double f) { return std::sqrt(f); }
inline double jsin(double f) { return std::sin(f); }
inline double jacos(double f) { return std::acos(f); }
inline double jabs(double f) { return std::fabs(f); }
template <typename T> inline double jval(const T& f) { return f.a; }
template <> inline double jval<double>(const double& f) { return f; }
template <typename T> inline T jconst(double s) { return T(s); }

template <typename T> struct V3 { T x, y, z; };
template <typename T> inline V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return V3<T>{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Point2LineFactor::operator() restated for T in {double, Jet<7>}. q = (x,y,z,w) storage.
template <typename T>
void point2line(const double* c, const double* a, const double* b, double min_d, double max_d,
                const T* q, const T* t, T* residual) {
  V3<T> cp{T(c[0]), T(c[1]), T(c[2])};
  V3<T> lpa{T(a[0]), T(a[1]), T(a[2])};
  V3<T> lpb{T(b[0]), T(b[1]), T(b[2])};
  // Eigen::Quaternion<T> q_last_curr{q[3], q[0], q[1], q[2]}; slerp(T(1)) from identity (Eigen 3.3 slerp).
  T qw = q[3], qx = q[0], qy = q[1], qz = q[2];
  {
    const double one = 1.0 - std::numeric_limits<double>::epsilon();
    T d = T(1.0) * qw + T(0.0) * qx + T(0.0) * qy + T(0.0) * qz;
    T absD = jabs(d);
    T scale0, scale1;
    const T tt = T(1.0);
    if (jval(absD) >= one) {
      scale0 = T(1.0) - tt; scale1 = tt;
    } else {
      T theta = jacos(absD);
      T sinTheta = jsin(theta);
      scale0 = jsin((T(1.0) - tt) * theta) / sinTheta;
      scale1 = jsin(tt * theta) / sinTheta;
    }
    if (jval(d) < 0.0) scale1 = -scale1;
    T nw = scale0 * T(1.0) + scale1 * qw, nx = scale0 * T(0.0) + scale1 * qx;
    T ny = scale0 * T(0.0) + scale1 * qy, nz = scale0 * T(0.0) + scale1 * qz;
    qw = nw; qx = nx; qy = ny; qz = nz;
  }
  V3<T> tl{T(1.0) * t[0], T(1.0) * t[1], T(1.0) * t[2]};
  // Eigen quaternion * vector: uv = vec x v; uv += uv; v + w*uv + vec x uv
  V3<T> qv{qx, qy, qz};
  V3<T> uv = cross(qv, cp);
  uv = V3<T>{uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  V3<T> c2 = cross(qv, uv);
  V3<T> lp{cp.x + qw * uv.x + c2.x + tl.x, cp.y + qw * uv.y + c2.y + tl.y, cp.z + qw * uv.z + c2.z + tl.z};
  V3<T> la{lp.x - lpa.x, lp.y - lpa.y, lp.z - lpa.z};
  V3<T> lb{lp.x - lpb.x, lp.y - lpb.y, lp.z - lpb.z};
  V3<T> nu = cross(la, lb);
  V3<T> de{lpa.x - lpb.x, lpa.y - lpb.y, lpa.z - lpb.z};
  V3<T> cpl{c[0] - t[0], c[1] - t[1], c[2] - t[2]};
  T d = jsqrt(cpl.x * cpl.x + cpl.y * cpl.y);
  d = (d - T(min_d)) / (T(max_d) - T(min_d));
  T w = T(1.01) - d;
  T den = jsqrt(de.x * de.x + de.y * de.y + de.z * de.z);
  residual[0] = w * (nu.x / den);
  residual[1] = w * (nu.y / den);
  residual[2] = w * (nu.z / den);
}

// residual + local (tangent) Jacobian 3x6: autodiff global 3x7 times the
// EigenQuaternionParameterization plus-Jacobian (Ceres local_parameterization.cc).
void factor_eval(const double* c, const double* a, const double* b, double min_d, double max_d,
                 const double* q, const double* t, double* r, double* J /*3x6 or null*/) {
  if (!J) {
    point2line<double>(c, a, b, min_d, max_d, q, t, r);
    return;
  }
  typedef Jet<7> J7;
  J7 jq[4], jt[3], jr[3];
  for (int k = 0; k < 4; ++k) jq[k] = J7(q[k], k);
  for (int k = 0; k < 3; ++k) jt[k] = J7(t[k], 4 + k);
  point2line<J7>(c, a, b, min_d, max_d, jq, jt, jr);
  // plus Jacobian (4x3 row-major) for storage (x,y,z,w)
  const double P[12] = {q[3], q[2], -q[1], -q[2], q[3], q[0], q[1], -q[0], q[3], -q[0], -q[1], -q[2]};
  for (int i = 0; i < 3; ++i) {
    r[i] = jr[i].a;
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += jr[i].v[k] * P[k * 3 + j];
      J[i * 6 + j] = s;
    }
    for (int j = 0; j < 3; ++j) J[i * 6 + 3 + j] = jr[i].v[4 + j];
  }
}

// EigenQuaternionParameterization::Plus, storage (x,y,z,w): x_plus = dq (x) x
void quat_plus(const double* x, const double* delta, double* out) {
  const double norm_delta = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (norm_delta > 0.0) {
    const double s = std::sin(norm_delta) / norm_delta;
    const double dw = std::cos(norm_delta), dx = s * delta[0], dy = s * delta[1], dz = s * delta[2];
    const double xw = x[3], xx = x[0], xy = x[1], xz = x[2];
    // Eigen quaternion product a*b
    out[3] = dw * xw - dx * xx - dy * xy - dz * xz;
    out[0] = dw * xx + dx * xw + dy * xz - dz * xy;
    out[1] = dw * xy + dy * xw + dz * xx - dx * xz;
    out[2] = dw * xz + dz * xw + dx * xy - dy * xx;
  } else {
    out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3];
  }
}
inline void state_plus(const double* x7, const double* d6, double* o7) {
  quat_plus(x7, d6, o7);
  o7[4] = x7[4] + d6[3]; o7[5] = x7[5] + d6[4]; o7[6] = x7[6] + d6[5];
}

// HuberLoss(a) (ceres/loss_function.cc): rho[0..2]
inline void huber(double a, double s, double* rho) {
  const double b = a * a;
  if (s > b) {
    const double r = std::sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

struct Problem {
  const double* cab; int n; double min_d, max_d; int threads;
  // Evaluate cost (and optionally corrected residuals r[3n] and Jacobian J[3n x 6]) at x7=(q,t).
  double evaluate(const double* x7, double* r, double* J) const {
    double cost = 0;
#pragma omp parallel for num_threads(threads) reduction(+ : cost) if (n > 256)
    for (int i = 0; i < n; ++i) {
      double ri[3], Ji[18];
      factor_eval(cab + 9 * i, cab + 9 * i + 3, cab + 9 * i + 6, min_d, max_d, x7, x7 + 4, ri, J ? Ji : nullptr);
      double s = ri[0] * ri[0] + ri[1] * ri[1] + ri[2] * ri[2];
      double rho[3]; huber(0.2, s, rho);
      cost += 0.5 * rho[0];
      // Corrector with rho'' <= 0: scale residual and Jacobian rows by sqrt(rho').
      const double sc = std::sqrt(rho[1]);
      if (r) { r[3 * i] = sc * ri[0]; r[3 * i + 1] = sc * ri[1]; r[3 * i + 2] = sc * ri[2]; }
      if (J) for (int k = 0; k < 18; ++k) J[18 * i + k] = sc * Ji[k];
    }
    return cost;
  }
};

// Householder QR least squares: minimise ||A y - b||, A m x 6 row-major (destroyed).
bool qr_solve6(std::vector<double>& A, std::vector<double>& b, int m, double* y) {
  const int n = 6;
  for (int k = 0; k < n; ++k) {
    double norm = 0;
    for (int i = k; i < m; ++i) norm += A[i * n + k] * A[i * n + k];
    norm = std::sqrt(norm);
    if (norm == 0.0) return false;
    double alpha = A[k * n + k] > 0 ? -norm : norm;
    double v0 = A[k * n + k] - alpha;
    // v = (v0, A[k+1..m-1][k]); beta = 2 / (v'v)
    double vtv = v0 * v0;
    for (int i = k + 1; i < m; ++i) vtv += A[i * n + k] * A[i * n + k];
    if (vtv == 0.0) return false;
    double beta = 2.0 / vtv;
    for (int j = k + 1; j < n; ++j) {
      double s = v0 * A[k * n + j];
      for (int i = k + 1; i < m; ++i) s += A[i * n + k] * A[i * n + j];
      s *= beta;
      A[k * n + j] -= s * v0;
      for (int i = k + 1; i < m; ++i) A[i * n + j] -= s * A[i * n + k];
    }
    {
      double s = v0 * b[k];
      for (int i = k + 1; i < m; ++i) s += A[i * n + k] * b[i];
      s *= beta;
      b[k] -= s * v0;
      for (int i = k + 1; i < m; ++i) b[i] -= s * A[i * n + k];
    }
    A[k * n + k] = alpha;
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s -= A[k * n + j] * y[j];
    if (A[k * n + k] == 0.0) return false;
    y[k] = s / A[k * n + k];
  }
  for (int k = 0; k < n; ++k) if (!std::isfinite(y[k])) return false;
  return true;
}

bool chol_solve6(const double* H, const double* g, double* y) {
  double L[36];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = H[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
      if (i == j) { if (!(s > 0.0)) return false; L[i * 6 + i] = std::sqrt(s); }
      else L[i * 6 + j] = s / L[j * 6 + j];
    }
  double z[6];
  for (int i = 0; i < 6; ++i) { double s = g[i]; for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * z[k]; z[i] = s / L[i * 6 + i]; }
  for (int i = 5; i >= 0; --i) { double s = z[i]; for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * y[k]; y[i] = s / L[i * 6 + i]; }
  for (int k = 0; k < 6; ++k) if (!std::isfinite(y[k])) return false;
  return true;
}

// ceres::Solve restated: TrustRegionMinimizer + LevenbergMarquardtStrategy + DenseQRSolver,
// Ceres 1.14 defaults, max_num_iterations = 4 (src/laser_odometry.cc:212-218).
void lm_solve(const Problem& pb, double* q, double* t, int linear_solver, OrcSolveSummary* sum) {
  OrcSolveSummary S; std::memset(&S, 0, sizeof(S));
  S.num_residual_blocks = pb.n;
  if (pb.n == 0) { S.termination = 4; if (sum) *sum = S; return; }
  const int n = pb.n, m = 3 * n;
  const int max_iterations = 4;
  const double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const double max_radius = 1e16, min_radius = 1e-32;
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int num_consecutive_invalid = 0;

  double x[7] = {q[0], q[1], q[2], q[3], t[0], t[1], t[2]};
  std::vector<double> r(m), J((size_t)m * 6), Js((size_t)m * 6);
  double scale[6], diagonal[6], g[6];
  double x_cost;

  auto x_norm_of = [](const double* v) { double s = 0; for (int k = 0; k < 7; ++k) s += v[k] * v[k]; return std::sqrt(s); };
  auto eval_grad_jac = [&](bool first) {
    x_cost = pb.evaluate(x, r.data(), J.data());
    S.jac_evals++;
    for (int j = 0; j < 6; ++j) { double s = 0; for (int i = 0; i < m; ++i) s += J[(size_t)i * 6 + j] * r[i]; g[j] = s; }
    if (first) {
      for (int j = 0; j < 6; ++j) { double s = 0; for (int i = 0; i < m; ++i) s += J[(size_t)i * 6 + j] * J[(size_t)i * 6 + j];
        scale[j] = 1.0 / (1.0 + std::sqrt(s)); }
    }
    for (int i = 0; i < m; ++i) for (int j = 0; j < 6; ++j) Js[(size_t)i * 6 + j] = J[(size_t)i * 6 + j] * scale[j];
    // gradient_max_norm = || x - Plus(x, -g) ||_inf
    double ng[6], xp[7]; for (int j = 0; j < 6; ++j) ng[j] = -g[j];
    state_plus(x, ng, xp);
    double mx = 0; for (int k = 0; k < 7; ++k) mx = std::max(mx, std::fabs(xp[k] - x[k]));
    return mx;
  };

  double gradient_max_norm = eval_grad_jac(true);
  S.initial_cost = x_cost;
  double x_norm = x_norm_of(x);
  bool step_is_successful = true;  // IterationZero()
  int iteration = 0;
  S.termination = 0;
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iteration >= max_iterations) { S.termination = 0; break; }
    if (step_is_successful && gradient_max_norm <= gradient_tolerance) { S.termination = 1; break; }
    if (radius < min_radius) { S.termination = 5; break; }
    iteration++;
    // ComputeTrustRegionStep: LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      for (int j = 0; j < 6; ++j) { double s = 0; for (int i = 0; i < m; ++i) s += Js[(size_t)i * 6 + j] * Js[(size_t)i * 6 + j];
        diagonal[j] = std::min(std::max(s, min_lm_diagonal), max_lm_diagonal); }
    }
    double D[6]; for (int j = 0; j < 6; ++j) D[j] = std::sqrt(diagonal[j] / radius);
    double y[6]; bool ok;
    if (linear_solver == 0) {
      std::vector<double> A((size_t)(m + 6) * 6, 0.0), rhs(m + 6, 0.0);
      std::memcpy(A.data(), Js.data(), sizeof(double) * (size_t)m * 6);
      for (int j = 0; j < 6; ++j) A[(size_t)(m + j) * 6 + j] = D[j];
      std::memcpy(rhs.data(), r.data(), sizeof(double) * m);
      ok = qr_solve6(A, rhs, m + 6, y);
    } else {
      double H[36], gs[6];
      for (int a = 0; a < 6; ++a) { for (int b = 0; b < 6; ++b) { double s = 0; for (int i = 0; i < m; ++i) s += Js[(size_t)i * 6 + a] * Js[(size_t)i * 6 + b]; H[a * 6 + b] = s; }
        H[a * 6 + a] += D[a] * D[a]; double s = 0; for (int i = 0; i < m; ++i) s += Js[(size_t)i * 6 + a] * r[i]; gs[a] = s; }
      ok = chol_solve6(H, gs, y);
    }
    reuse_diagonal = true;
    bool step_valid = false; double model_cost_change = 0; double step[6], delta[6];
    if (ok) {
      for (int j = 0; j < 6; ++j) step[j] = -y[j];
      // model_cost_change = -(J step)'(r + J step / 2)
      double acc = 0;
      for (int i = 0; i < m; ++i) { double mr = 0; for (int j = 0; j < 6; ++j) mr += Js[(size_t)i * 6 + j] * step[j]; acc += mr * (r[i] + mr / 2.0); }
      model_cost_change = -acc;
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) {  // HandleInvalidStep
      if (++num_consecutive_invalid >= 5) { S.termination = 5; break; }
      radius *= 0.5; reuse_diagonal = true; step_is_successful = false;
      continue;
    }
    num_consecutive_invalid = 0;
    for (int j = 0; j < 6; ++j) delta[j] = step[j] * scale[j];
    // ComputeCandidatePointAndEvaluateCost
    double xc[7]; state_plus(x, delta, xc);
    double candidate_cost = pb.evaluate(xc, nullptr, nullptr);
    S.cost_evals++;
    if (!std::isfinite(candidate_cost)) candidate_cost = x_cost;
    // ParameterToleranceReached
    double step_norm = 0; for (int k = 0; k < 7; ++k) step_norm += (x[k] - xc[k]) * (x[k] - xc[k]);
    step_norm = std::sqrt(step_norm);
    if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) { S.termination = 2; break; }
    // FunctionToleranceReached
    double cost_change = x_cost - candidate_cost;
    if (std::fabs(cost_change) <= function_tolerance * x_cost) { S.termination = 3; break; }
    // IsStepSuccessful
    double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > min_relative_decrease) {  // HandleSuccessfulStep
      std::memcpy(x, xc, sizeof(x));
      x_norm = x_norm_of(x);
      gradient_max_norm = eval_grad_jac(false);
      step_is_successful = true; S.successful_steps++;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
      radius = std::min(max_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
    } else {  // HandleUnsuccessfulStep
      step_is_successful = false;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
  }
  S.iterations = iteration;
  S.final_cost = x_cost;
  q[0] = x[0]; q[1] = x[1]; q[2] = x[2]; q[3] = x[3]; t[0] = x[4]; t[1] = x[5]; t[2] = x[6];
  if (sum) *sum = S;
}

// ---------------------------------------------------------------------------------
// Eigen pieces (A.6): isometry algebra on row-major 4x4, quaternion <-> matrix
// ---------------------------------------------------------------------------------
struct Iso { double m[16]; };
inline Iso iso_identity() { Iso I; std::memset(I.m, 0, sizeof(I.m)); I.m[0] = I.m[5] = I.m[10] = I.m[15] = 1.0; return I; }
inline Iso iso_mul(const Iso& A, const Iso& B) {  // linear = La*Lb ; translation = La*tb + ta
  Iso C = iso_identity();
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) C.m[i * 4 + j] = A.m[i * 4 + 0] * B.m[0 * 4 + j] + A.m[i * 4 + 1] * B.m[1 * 4 + j] + A.m[i * 4 + 2] * B.m[2 * 4 + j];
    C.m[i * 4 + 3] = (A.m[i * 4 + 0] * B.m[3] + A.m[i * 4 + 1] * B.m[7] + A.m[i * 4 + 2] * B.m[11]) + A.m[i * 4 + 3];
  }
  return C;
}
inline Iso iso_inverse(const Iso& A) {  // Isometry: (R', -R' t)
  Iso C = iso_identity();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C.m[i * 4 + j] = A.m[j * 4 + i];
  for (int i = 0; i < 3; ++i) C.m[i * 4 + 3] = (-C.m[i * 4 + 0]) * A.m[3] + (-C.m[i * 4 + 1]) * A.m[7] + (-C.m[i * 4 + 2]) * A.m[11];
  return C;
}
// Eigen::Quaterniond(Matrix3d) -> (x,y,z,w)
inline void quat_from_matrix(const Iso& A, double* q) {
  auto M = [&](int r, int c) { return A.m[r * 4 + c]; };
  double tr = M(0, 0) + M(1, 1) + M(2, 2);
  if (tr > 0.0) {
    double tt = std::sqrt(tr + 1.0);
    q[3] = 0.5 * tt; tt = 0.5 / tt;
    q[0] = (M(2, 1) - M(1, 2)) * tt; q[1] = (M(0, 2) - M(2, 0)) * tt; q[2] = (M(1, 0) - M(0, 1)) * tt;
  } else {
    int i = 0;
    if (M(1, 1) > M(0, 0)) i = 1;
    if (M(2, 2) > M(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double tt = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
    q[i] = 0.5 * tt; tt = 0.5 / tt;
    q[3] = (M(k, j) - M(j, k)) * tt;
    q[j] = (M(j, i) + M(i, j)) * tt;
    q[k] = (M(k, i) + M(i, k)) * tt;
  }
}
// Quaterniond::toRotationMatrix (no normalisation, as src/laser_odometry.cc:225-226)
inline void matrix_from_quat(const double* q, Iso& A) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  A.m[0] = 1 - (tyy + tzz); A.m[1] = txy - twz; A.m[2] = txz + twy;
  A.m[4] = txy + twz; A.m[5] = 1 - (txx + tzz); A.m[6] = tyz - twx;
  A.m[8] = txz - twy; A.m[9] = tyz + twx; A.m[10] = 1 - (txx + tyy);
}

// ---------------------------------------------------------------------------------
// tf (bullet LinearMath) pieces used by the IMU override and publishOdom
// (src/laser_odometry.cc:152-183, :395-446): Matrix3x3(Quaternion) = setRotation, getRPY =
// getEulerYPR solution 1, setRPY = setEulerYPR, getRotation.  Third-party (ros/geometry tf,
// Noetic 1.13), source not in /root/reference: restated from the published header; parity unpinned.
// Matrices row-major 3x3; quaternions (x,y,z,w).
// ---------------------------------------------------------------------------------
inline void tf_matrix_from_quat(const double* q, double* m) {
  const double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const double s = 2.0 / d;
  const double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
  const double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
  const double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs, yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
  m[0] = 1.0 - (yy + zz); m[1] = xy - wz; m[2] = xz + wy;
  m[3] = xy + wz; m[4] = 1.0 - (xx + zz); m[5] = yz - wx;
  m[6] = xz - wy; m[7] = yz + wx; m[8] = 1.0 - (xx + yy);
}
inline void tf_get_rpy(const double* m, double* roll, double* pitch, double* yaw) {
  const double kPi = 3.14159265358979323846;
  if (std::fabs(m[6]) >= 1.0) {
    *yaw = 0.0;
    if (m[6] < 0.0) { *pitch = kPi / 2.0; *roll = std::atan2(m[1], m[2]); }
    else { *pitch = -kPi / 2.0; *roll = std::atan2(-m[1], -m[2]); }
  } else {
    *pitch = -std::asin(m[6]);
    const double cp = std::cos(*pitch);
    *roll = std::atan2(m[7] / cp, m[8] / cp);
    *yaw = std::atan2(m[3] / cp, m[0] / cp);
  }
}
inline void tf_set_rpy(double roll, double pitch, double yaw, double* m) {
  const double ci = std::cos(roll), cj = std::cos(pitch), ch = std::cos(yaw), si = std::sin(roll), sj = std::sin(pitch), sh = std::sin(yaw);
  const double cc = ci * ch, cs = ci * sh, sc = si * ch, ss = si * sh;
  m[0] = cj * ch; m[1] = sj * sc - cs; m[2] = sj * cc + ss;
  m[3] = cj * sh; m[4] = sj * ss + cc; m[5] = sj * cs - sc;
  m[6] = -sj; m[7] = cj * si; m[8] = cj * ci;
}
inline void tf_get_rotation(const double* m, double* q) {
  const double trace = m[0] + m[4] + m[8];
  if (trace > 0.0) {
    double s = std::sqrt(trace + 1.0);
    q[3] = s * 0.5; s = 0.5 / s;
    q[0] = (m[7] - m[5]) * s; q[1] = (m[2] - m[6]) * s; q[2] = (m[3] - m[1]) * s;
  } else {
    const int i = m[0] < m[4] ? (m[4] < m[8] ? 2 : 1) : (m[0] < m[8] ? 2 : 0);
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
    q[i] = s * 0.5; s = 0.5 / s;
    q[3] = (m[k * 3 + j] - m[j * 3 + k]) * s; q[j] = (m[j * 3 + i] + m[i * 3 + j]) * s; q[k] = (m[k * 3 + i] + m[i * 3 + k]) * s;
  }
}
// src/laser_odometry.cc:152-183
inline Iso imu_override(const Iso& odom, const double* imu_q, const Iso& l2b) {
  double imu_m[9], r_imu, p_imu, y_imu;
  tf_matrix_from_quat(imu_q, imu_m);
  tf_get_rpy(imu_m, &r_imu, &p_imu, &y_imu);
  Iso bl = iso_mul(odom, l2b);
  double q[4], m[9], r_bl, p_bl, y_bl;
  quat_from_matrix(bl, q);
  tf_matrix_from_quat(q, m);
  tf_get_rpy(m, &r_bl, &p_bl, &y_bl);
  tf_set_rpy(r_imu, p_imu, y_bl, m);
  tf_get_rotation(m, q);
  Iso R = iso_identity();
  matrix_from_quat(q, R);
  for (int i = 0; i < 3; ++i) for (int jj = 0; jj < 3; ++jj) bl.m[i * 4 + jj] = R.m[i * 4 + jj];
  return iso_mul(bl, iso_inverse(l2b));
}

// ---------------------------------------------------------------------------------
// A.3 pcl::VoxelGrid<PointXYZI>::applyFilter (PCL 1.10 voxel_grid.hpp), cubic leaf,
// downsample_all_data = true, min_points_per_voxel = 0, no filter field.
// ---------------------------------------------------------------------------------
int voxel_grid(const std::vector<P4>& in, float leaf, std::vector<P4>& out) {
  out.clear();
  if (in.empty()) return 0;
  const float inv = 1.0f / leaf;
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-mn[0], -mn[1], -mn[2]};
  for (const P4& p : in) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) { out = in; return -1; }
  int minb[3], maxb[3], divb[3];
  for (int k = 0; k < 3; ++k) { minb[k] = (int)std::floor(mn[k] * inv); maxb[k] = (int)std::floor(mx[k] * inv); divb[k] = maxb[k] - minb[k] + 1; }
  const int mul1 = divb[0], mul2 = divb[0] * divb[1];
  struct IP { unsigned idx; unsigned pi; };
  std::vector<IP> iv; iv.reserve(in.size());
  for (unsigned i = 0; i < in.size(); ++i) {
    const P4& p = in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    int i0 = (int)(std::floor(p.x * inv) - (float)minb[0]);
    int i1 = (int)(std::floor(p.y * inv) - (float)minb[1]);
    int i2 = (int)(std::floor(p.z * inv) - (float)minb[2]);
    iv.push_back(IP{(unsigned)(i0 + i1 * mul1 + i2 * mul2), i});
  }
  // PCL: std::sort on idx only (unstable). Canonical choice here: stable (input order).
  std::stable_sort(iv.begin(), iv.end(), [](const IP& a, const IP& b) { return a.idx < b.idx; });
  size_t k = 0;
  while (k < iv.size()) {
    size_t e = k + 1;
    while (e < iv.size() && iv[e].idx == iv[k].idx) ++e;
    float sx = 0, sy = 0, sz = 0, si = 0;
    for (size_t u = k; u < e; ++u) { const P4& p = in[iv[u].pi]; sx += p.x; sy += p.y; sz += p.z; si += p.i; }
    const float cnt = (float)(e - k);  // CentroidPoint::get: accumulated / n
    out.push_back(P4{sx / cnt, sy / cnt, sz / cnt, si / cnt});
    k = e;
  }
  return (int)out.size();
}

}  // namespace

// ===================================================================================
// LocalMapManager (src/laser_odometry.cc:24-69)
// ===================================================================================
struct OrcLmap {
  std::vector<P4> total_points;
  size_t nframes = 0, max_nframes = 0;
  std::queue<size_t> sizes;
  void add(const P4* pc, size_t n) {
    total_points.insert(total_points.end(), pc, pc + n);
    nframes++;
    sizes.push(n);
    if (nframes > max_nframes) {
      size_t pc_size = sizes.front(); sizes.pop();
      total_points.erase(total_points.begin(), total_points.begin() + pc_size);  // ExtractIndices(negative), order kept
      nframes--;
    }
  }
};

// ===================================================================================
// LaserOdometer (src/laser_odometry.cc:100-366)
// ===================================================================================
struct OrcOdom {
  OrcParams p;
  bool init = false;
  Iso prev_odom = iso_identity(), odom = iso_identity();
  double param_q[4] = {0, 0, 0, 1}, param_t[3] = {0, 0, 0};
  bool use_imu = false;                       // params->use_imu_
  double imu_q[4] = {0, 0, 0, 1};             // SharedData::getLastIMUOri
  Iso laser_to_base = iso_identity();         // getBaseToLaserTf
  OrcLmap lmap;
  std::vector<P4> received;  // SharedData::local_map_
};

namespace {

struct AssocOut { std::vector<double> cab; int matches = 0; };

// src/laser_odometry.cc:300-366
void add_edge_constraints(const OrcParams& p, const P4* edges, int E, const std::vector<P4>& local_map,
                          const Iso& pose, int knn_method, AssocOut* blocks,
                          int32_t* knn_idx, float* knn_d2, uint8_t* gate, double* eig, float* q_world, uint8_t* tie) {
  std::vector<P4> edges_map(E);
  for (int i = 0; i < E; ++i) edges_map[i] = transform_point(pose.m, edges[i]);
  const int M = (int)local_map.size();
  std::unique_ptr<KdTree> tree;
  if (knn_method == 1) tree.reset(new KdTree(local_map.data(), M));
  std::vector<Knn5> res(E);
  // The reference's association loop is serial (kd-tree path). The brute-force path is a
  // test-only cross-check and may use all cores.
#pragma omp parallel for schedule(dynamic, 64) if (knn_method == 0)
  for (int i = 0; i < E; ++i) {
    Knn5& r = res[i];
    if (knn_method == 1) tree->knn(edges_map[i], r);
    else for (int j = 0; j < M; ++j) {
      const P4& mp = local_map[j];
      if (!std::isfinite(mp.x) || !std::isfinite(mp.y) || !std::isfinite(mp.z)) continue;
      r.push(l2_simple(edges_map[i], mp), j);
    }
  }
  for (int i = 0; i < E; ++i) {
    const Knn5& r = res[i];
    uint8_t gt = 0; double ev[3] = {0, 0, 0};
    if (r.id[4] >= 0 && r.d[4] < 1.0) {  // :324 (M<5 is UB in the reference; treated as gate failure)
      gt |= 1;
      P4 nn[5]; for (int j = 0; j < 5; ++j) nn[j] = local_map[r.id[j]];
      if (line_gate(nn, ev)) {
        gt |= 2;
        if (blocks) {
          const double v[9] = {edges[i].x, edges[i].y, edges[i].z, nn[0].x, nn[0].y, nn[0].z, nn[1].x, nn[1].y, nn[1].z};
          blocks->cab.insert(blocks->cab.end(), v, v + 9);
          blocks->matches++;
        }
      }
    }
    if (knn_idx) for (int j = 0; j < 5; ++j) knn_idx[5 * i + j] = r.id[j];
    if (knn_d2) for (int j = 0; j < 5; ++j) knn_d2[5 * i + j] = r.d[j];
    if (gate) gate[i] = gt;
    if (eig) { eig[3 * i] = ev[0]; eig[3 * i + 1] = ev[1]; eig[3 * i + 2] = ev[2]; }
    if (q_world) std::memcpy(q_world + 4 * i, &edges_map[i], 16);
    if (tie) tie[i] = knn_tie(r) ? 1 : 0;
  }
}

}  // namespace

extern "C" {

void orc_default_params(OrcParams* p) {
  p->min_range = 3.0; p->max_range = 75.0; p->lidar_type = 0; p->scan_lines = 64; p->scan_regions = 8;
  p->edges_per_region = 10; p->prev_frames = 5; p->filter_local_map = 0; p->mapping = 0; p->omp_threads = 0;
}

int orc_split(const OrcParams* p, const float* pts, int n, int stride_f, int width, int height,
              int32_t* ring_of_point, float* rings_xyzi, int32_t* ring_offsets, int32_t* src_index) {
  const int L = p->scan_lines;
  std::vector<std::vector<int>> scans(L);
  int bad = 0;
  if (p->lidar_type == 0) {
    if (L != 64 && L != 32 && L != 16) bad = 1;
    for (int i = 0; i < n; i++) {
      const float* q = pts + (size_t)i * stride_f;
      double x = q[0], y = q[1], z = q[2], distance;
      int id = -1;
      if (is_valid_point(*p, x, y, z, &distance)) id = velodyne_ring(*p, z, distance);
      if (ring_of_point) ring_of_point[i] = id;
      if (id != -1) scans[id].push_back(i);
    }
  } else if (p->lidar_type == 1) {
    // src/feature_extractor.cc:160-175; rows beyond scan_lines would overrun `scans` in the
    // reference (UB) — reported as an error here.
    if (height > L) bad = 1;
    for (int row = 0; row < height && row < L; row++)
      for (int col = 0; col < width; col++) {
        int i = row * width + col;
        const float* q = pts + (size_t)i * stride_f;
        double x = q[0], y = q[1], z = q[2], distance;
        int id = is_valid_point(*p, x, y, z, &distance) ? row : -1;
        if (ring_of_point) ring_of_point[i] = id;
        if (id != -1) scans[id].push_back(i);
      }
    if (ring_of_point) for (int i = std::min(height, L) * width; i < n; ++i) ring_of_point[i] = -1;
  } else {
    bad = 1;
    if (ring_of_point) for (int i = 0; i < n; ++i) ring_of_point[i] = -1;
  }
  int pos = 0;
  for (int r = 0; r < L; ++r) {
    if (ring_offsets) ring_offsets[r] = pos;
    for (int i : scans[r]) {
      if (rings_xyzi) std::memcpy(rings_xyzi + (size_t)pos * 4, pts + (size_t)i * stride_f, 16);
      if (src_index) src_index[pos] = i;
      ++pos;
    }
  }
  if (ring_offsets) ring_offsets[L] = pos;
  return bad ? -1 : pos;
}

int orc_extract(const OrcParams* p, const float* rings_xyzi, const int32_t* ring_offsets,
                float* edges_xyzi, int32_t* edge_ring, int32_t* edge_idx, double* keys, int sort_mode, int cap) {
  static thread_local Extractor ex;
  std::vector<P4> out; std::vector<int> oring, oidx;
  if (keys) { int tot = ring_offsets[p->scan_lines]; for (int i = 0; i < tot; ++i) keys[i] = std::numeric_limits<double>::quiet_NaN(); }
  extract_features(*p, ex, reinterpret_cast<const P4*>(rings_xyzi), ring_offsets, sort_mode, keys, out, oring, oidx);
  int n = (int)out.size();
  int w = std::min(n, cap);
  if (edges_xyzi) std::memcpy(edges_xyzi, out.data(), (size_t)w * 16);
  if (edge_ring) std::memcpy(edge_ring, oring.data(), (size_t)w * 4);
  if (edge_idx) std::memcpy(edge_idx, oidx.data(), (size_t)w * 4);
  return n;
}

int orc_extract_scan(const OrcParams* p, const float* pts, int n, int stride_f, int width, int height,
                     float* edges_xyzi, int cap, double* times_us) {
  std::vector<float> rings((size_t)n * 4);
  std::vector<int32_t> off(p->scan_lines + 1);
  auto t0 = Clock::now();
  int nv = orc_split(p, pts, n, stride_f, width, height, nullptr, rings.data(), off.data(), nullptr);
  double ts = us_since(t0);
  if (nv < 0) return 0;
  t0 = Clock::now();
  int e = orc_extract(p, rings.data(), off.data(), edges_xyzi, nullptr, nullptr, nullptr, 0, cap);
  double te = us_since(t0);
  if (times_us) { times_us[0] = ts; times_us[1] = te; }
  return e;
}

void orc_transform(const float* in_xyzi, int n, const double* T, float* out_xyzi) {
  const P4* in = reinterpret_cast<const P4*>(in_xyzi); P4* out = reinterpret_cast<P4*>(out_xyzi);
  for (int i = 0; i < n; ++i) out[i] = transform_point(T, in[i]);
}

void orc_knn5(const float* map_xyzi, int M, const float* q_xyzi, int E, int method, int32_t* idx, float* d2, uint8_t* tie) {
  const P4* mp = reinterpret_cast<const P4*>(map_xyzi); const P4* q = reinterpret_cast<const P4*>(q_xyzi);
  std::unique_ptr<KdTree> tree; if (method == 1) tree.reset(new KdTree(mp, M));
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < E; ++i) {
    Knn5 r;
    if (method == 1) tree->knn(q[i], r);
    else for (int j = 0; j < M; ++j) {
      if (!std::isfinite(mp[j].x) || !std::isfinite(mp[j].y) || !std::isfinite(mp[j].z)) continue;
      r.push(l2_simple(q[i], mp[j]), j);
    }
    for (int k = 0; k < 5; ++k) { idx[5 * i + k] = r.id[k]; d2[5 * i + k] = r.d[k]; }
    if (tie) tie[i] = knn_tie(r) ? 1 : 0;
  }
}

void orc_associate(const float* edges_xyzi, int E, const double* T, const float* map_xyzi, int M, int knn_method,
                   int32_t* knn_idx, float* knn_d2, uint8_t* gate, double* eig, float* q_world, uint8_t* tie) {
  OrcParams p; orc_default_params(&p);
  Iso pose; std::memcpy(pose.m, T, sizeof(pose.m));
  std::vector<P4> lm(reinterpret_cast<const P4*>(map_xyzi), reinterpret_cast<const P4*>(map_xyzi) + M);
  add_edge_constraints(p, reinterpret_cast<const P4*>(edges_xyzi), E, lm, pose, knn_method, nullptr,
                       knn_idx, knn_d2, gate, eig, q_world, tie);
}

/* A.6 alone: eigenvalues (ascending) of a symmetric 3x3, row-major; used to check the GPU line-gate screen. */
void orc_sym3_eigenvalues(const double* A9, double* w3) { sym3_eigenvalues(A9, w3); }

void orc_factor(const double* c, const double* a, const double* b, double min_range, double max_range,
                const double* q, const double* t, double* r3, double* J18) {
  factor_eval(c, a, b, min_range, max_range, q, t, r3, J18);
}

void orc_solve(const double* cab, int nblocks, double min_range, double max_range, double* q, double* t,
               int linear_solver, OrcSolveSummary* sum) {
  Problem pb{cab, nblocks, min_range, max_range, (int)sysconf(_SC_NPROCESSORS_ONLN)};
  lm_solve(pb, q, t, linear_solver, sum);
}

OrcOdom* orc_odom_create(const OrcParams* p) {
  OrcOdom* o = new OrcOdom;
  o->p = *p;
  o->lmap.max_nframes = (size_t)p->prev_frames;
  return o;
}
void orc_odom_destroy(OrcOdom* o) { delete o; }
void orc_odom_set_pose(OrcOdom* o, const double* odom, const double* prev_odom) {
  if (odom) std::memcpy(o->odom.m, odom, sizeof(o->odom.m));
  if (prev_odom) std::memcpy(o->prev_odom.m, prev_odom, sizeof(o->prev_odom.m));
}
void orc_odom_get_pose(const OrcOdom* o, double* odom, double* prev_odom) {
  if (odom) std::memcpy(odom, o->odom.m, sizeof(o->odom.m));
  if (prev_odom) std::memcpy(prev_odom, o->prev_odom.m, sizeof(o->prev_odom.m));
}
int orc_odom_window_size(const OrcOdom* o) { return (int)o->lmap.total_points.size(); }
int orc_odom_window_frames(const OrcOdom* o) { return (int)o->lmap.nframes; }
void orc_odom_get_window(const OrcOdom* o, float* xyzi) { std::memcpy(xyzi, o->lmap.total_points.data(), o->lmap.total_points.size() * 16); }
void orc_odom_set_window(OrcOdom* o, const float* xyzi, const int32_t* frame_sizes, int nframes) {
  size_t tot = 0; std::queue<size_t> q;
  for (int i = 0; i < nframes; ++i) { q.push((size_t)frame_sizes[i]); tot += frame_sizes[i]; }
  o->lmap.total_points.assign(reinterpret_cast<const P4*>(xyzi), reinterpret_cast<const P4*>(xyzi) + tot);
  o->lmap.sizes = q; o->lmap.nframes = (size_t)nframes; o->init = nframes > 0;
}
void orc_odom_set_imu(OrcOdom* o, int use_imu, const double* q_xyzw, const double* laser_to_base16) {
  o->use_imu = use_imu != 0;
  if (q_xyzw) std::memcpy(o->imu_q, q_xyzw, sizeof(o->imu_q));
  if (laser_to_base16) std::memcpy(o->laser_to_base.m, laser_to_base16, sizeof(o->laser_to_base.m));
}

void orc_tf_rpy(const double* q_xyzw, double* rpy3, double* q_back_xyzw) {
  double m[9], m2[9];
  tf_matrix_from_quat(q_xyzw, m);
  tf_get_rpy(m, rpy3, rpy3 + 1, rpy3 + 2);
  tf_set_rpy(rpy3[0], rpy3[1], rpy3[2], m2);
  tf_get_rotation(m2, q_back_xyzw);
}

void orc_imu_override(const double* odom16, const double* imu_q_xyzw, const double* l2b16, double* out16) {
  Iso o, l; std::memcpy(o.m, odom16, sizeof(o.m)); std::memcpy(l.m, l2b16, sizeof(l.m));
  Iso r = imu_override(o, imu_q_xyzw, l);
  std::memcpy(out16, r.m, sizeof(r.m));
}

// publishOdom (src/laser_odometry.cc:395-446): out[13] = orientation x,y,z,w, position x,y,z,
// twist linear x,y,z, twist angular x,y,z.  prev_odom16 = prev_odom_ at the time of the call.
void orc_publish_odom(const double* pose16, const double* prev_odom16, const double* l2b16, double delta_time, double* out13) {
  Iso pose, prev, l2b;
  std::memcpy(pose.m, pose16, sizeof(pose.m)); std::memcpy(prev.m, prev_odom16, sizeof(prev.m)); std::memcpy(l2b.m, l2b16, sizeof(l2b.m));
  const Iso obl = iso_mul(pose, l2b);
  quat_from_matrix(obl, out13);
  out13[4] = obl.m[3]; out13[5] = obl.m[7]; out13[6] = obl.m[11];
  const Iso delta = iso_mul(iso_inverse(iso_mul(prev, l2b)), obl);
  out13[7] = delta.m[3] / delta_time; out13[8] = delta.m[7] / delta_time; out13[9] = delta.m[11] / delta_time;
  double qd[4], m[9], r, p, y;
  quat_from_matrix(delta, qd);
  tf_matrix_from_quat(qd, m);
  tf_get_rpy(m, &r, &p, &y);
  out13[10] = r / delta_time; out13[11] = p / delta_time; out13[12] = y / delta_time;
}

void orc_odom_set_received_map(OrcOdom* o, const float* xyzi, int n) {
  o->received.assign(reinterpret_cast<const P4*>(xyzi), reinterpret_cast<const P4*>(xyzi) + n);
}

void orc_odom_process(OrcOdom* o, const float* edges_xyzi, int E, double* pose_out, OrcFrameDiag* diag) {
  const P4* feats = reinterpret_cast<const P4*>(edges_xyzi);
  OrcFrameDiag D; std::memset(&D, 0, sizeof(D)); D.n_edges = E;
  if (!o->init) {  // :108-136
    auto t0 = Clock::now();
    o->lmap.add(feats, (size_t)E);
    o->init = true;
    D.times_us[3] = us_since(t0);
    std::memcpy(D.pred_pose, o->odom.m, sizeof(D.pred_pose));
  } else {
    auto t0 = Clock::now();
    // computeLocalMap (:274-298)
    std::vector<P4> local_map_rec = o->received;
    std::vector<P4> filtered;
    const std::vector<P4>* local_map_gen = &o->lmap.total_points;
    if (o->p.filter_local_map && o->lmap.nframes == (size_t)o->p.prev_frames && !o->p.mapping) {
      voxel_grid(o->lmap.total_points, 0.4f, filtered);
      local_map_gen = &filtered;
    }
    D.times_us[0] = us_since(t0);
    // prediction (:148-150)
    Iso pred = iso_mul(o->odom, iso_mul(iso_inverse(o->prev_odom), o->odom));
    o->prev_odom = o->odom;
    o->odom = pred;
    if (o->use_imu) o->odom = imu_override(o->odom, o->imu_q, o->laser_to_base);   // :152-183
    std::memcpy(D.pred_pose, o->odom.m, sizeof(D.pred_pose));
    // initial guess (:186-195)
    quat_from_matrix(o->odom, o->param_q);
    o->param_t[0] = o->odom.m[3]; o->param_t[1] = o->odom.m[7]; o->param_t[2] = o->odom.m[11];
    for (int optim_it = 0; optim_it < 2; optim_it++) {  // :198
      t0 = Clock::now();
      std::vector<P4> local_map(*local_map_gen);  // :310-314
      if (o->p.mapping) local_map.insert(local_map.end(), local_map_rec.begin(), local_map_rec.end());
      AssocOut blocks;
      add_edge_constraints(o->p, feats, E, local_map, o->odom, 1, &blocks, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
      D.n_map[optim_it] = (int)local_map.size(); D.n_matches[optim_it] = blocks.matches;
      D.times_us[1] += us_since(t0);
      t0 = Clock::now();
      // options.num_threads = sysconf(_SC_NPROCESSORS_ONLN) (src/laser_odometry.cc:216); omp_threads > 0
      // overrides it so that several independent sequences can share the host (bench reference arm).
      const int solver_threads = o->p.omp_threads > 0 ? o->p.omp_threads : (int)sysconf(_SC_NPROCESSORS_ONLN);
      Problem pb{blocks.cab.data(), blocks.matches, o->p.min_range, o->p.max_range, solver_threads};
      lm_solve(pb, o->param_q, o->param_t, 0, &D.solve[optim_it]);
      o->odom = iso_identity();  // :222-227
      matrix_from_quat(o->param_q, o->odom);
      o->odom.m[3] = o->param_t[0]; o->odom.m[7] = o->param_t[1]; o->odom.m[11] = o->param_t[2];
      D.times_us[2] += us_since(t0);
    }
    t0 = Clock::now();
    std::vector<P4> edges_map(E);  // :231-235
    for (int i = 0; i < E; ++i) edges_map[i] = transform_point(o->odom.m, feats[i]);
    o->lmap.add(edges_map.data(), (size_t)E);
    D.times_us[3] = us_since(t0);
  }
  if (pose_out) std::memcpy(pose_out, o->odom.m, sizeof(o->odom.m));
  if (diag) *diag = D;
}

OrcLmap* orc_lmap_create(int max_frames) { OrcLmap* m = new OrcLmap; m->max_nframes = (size_t)max_frames; return m; }
void orc_lmap_destroy(OrcLmap* m) { delete m; }
void orc_lmap_add(OrcLmap* m, const float* xyzi, int n) { m->add(reinterpret_cast<const P4*>(xyzi), (size_t)n); }
int orc_lmap_size(const OrcLmap* m) { return (int)m->total_points.size(); }
int orc_lmap_frames(const OrcLmap* m) { return (int)m->nframes; }
void orc_lmap_get(const OrcLmap* m, float* xyzi) { std::memcpy(xyzi, m->total_points.data(), m->total_points.size() * 16); }
void orc_lmap_set_max_frames(OrcLmap* m, int max_frames) { m->max_nframes = (size_t)max_frames; }

void orc_decode_cloud2(const uint8_t* data, int width, int height, int point_step, int row_step,
                       int off_x, int off_y, int off_z, int off_i, float* out_xyzi) {
  for (int r = 0; r < height; ++r)
    for (int c = 0; c < width; ++c) {
      const uint8_t* src = data + (size_t)r * row_step + (size_t)c * point_step;
      float* o = out_xyzi + ((size_t)r * width + c) * 4;
      std::memcpy(o, src + off_x, 4); std::memcpy(o + 1, src + off_y, 4); std::memcpy(o + 2, src + off_z, 4);
      o[3] = 0.f;
      if (off_i >= 0) std::memcpy(o + 3, src + off_i, 4);
    }
}

int orc_voxelgrid(const float* in_xyzi, int n, float leaf, float* out_xyzi) {
  std::vector<P4> in(reinterpret_cast<const P4*>(in_xyzi), reinterpret_cast<const P4*>(in_xyzi) + n), out;
  int r = voxel_grid(in, leaf, out);
  std::memcpy(out_xyzi, out.data(), out.size() * 16);
  return r;
}

}  // extern "C"

// ===================================================================================
// Map / Cell / HashKey (src/map.cc:24-189, include/liodom/map.h:39-116)
// ===================================================================================
struct OrcCell { std::vector<P4> points; bool modified = false; int key[3]; };
struct OrcMap {
  double xy, inv_xy, xy_half, zs, inv_z, z_half, res;
  struct Key { int x, y, z; bool operator==(const Key& o) const { return x == o.x && y == o.y && z == o.z; } };
  struct KeyHash { size_t operator()(const Key& k) const {
    size_t h1 = std::hash<int>()(k.x), h2 = std::hash<int>()(k.y), h3 = std::hash<int>()(k.z);
    return (h1 ^ (h2 << 1)) ^ (h3 << 2); } };
  std::unordered_map<Key, OrcCell*, KeyHash> cells;
  std::vector<OrcCell*> cells_vector;
  ~OrcMap() { for (OrcCell* c : cells_vector) delete c; }
  inline Key key_of(double x, double y, double z) const {  // src/map.cc:103-105
    return Key{int(std::floor(x * inv_xy) * xy + xy_half), int(std::floor(y * inv_xy) * xy + xy_half),
               int(std::floor(z * inv_z) * zs + z_half)};
  }
};

extern "C" {

OrcMap* orc_map_create(double xy_size, double z_size, double resolution) {
  OrcMap* m = new OrcMap;
  m->xy = xy_size; m->inv_xy = 1.0 / xy_size; m->xy_half = xy_size / 2.0;
  m->zs = z_size; m->inv_z = 1.0 / z_size; m->z_half = z_size / 2.0; m->res = resolution;
  return m;
}
void orc_map_destroy(OrcMap* m) { delete m; }

void orc_map_update(OrcMap* m, const float* pts_xyzi, int n, const double* T) {  // src/map.cc:90-129
  const P4* in = reinterpret_cast<const P4*>(pts_xyzi);
  for (int i = 0; i < n; ++i) {
    P4 point = transform_point(T, in[i]);
    OrcMap::Key key = m->key_of(point.x, point.y, point.z);
    OrcCell* cell;
    auto it = m->cells.find(key);
    if (it == m->cells.end()) {
      cell = new OrcCell; cell->key[0] = key.x; cell->key[1] = key.y; cell->key[2] = key.z;
      m->cells[key] = cell; m->cells_vector.push_back(cell);
    } else cell = it->second;
    cell->points.push_back(point); cell->modified = true;
  }
  const float leaf = (float)m->res;  // setLeafSize(double->float)
  for (OrcCell* c : m->cells_vector)
    if (c->modified) { std::vector<P4> out; voxel_grid(c->points, leaf, out); c->points.swap(out); c->modified = false; }
}
int orc_map_size(const OrcMap* m) { size_t s = 0; for (OrcCell* c : m->cells_vector) s += c->points.size(); return (int)s; }
int orc_map_num_cells(const OrcMap* m) { return (int)m->cells_vector.size(); }
void orc_map_get(const OrcMap* m, float* xyzi) {  // src/map.cc:131-139
  size_t pos = 0;
  for (OrcCell* c : m->cells_vector) { std::memcpy(xyzi + pos * 4, c->points.data(), c->points.size() * 16); pos += c->points.size(); }
}
void orc_map_cell_info(const OrcMap* m, int i, int32_t* key3, int32_t* count) {
  const OrcCell* c = m->cells_vector[i];
  key3[0] = c->key[0]; key3[1] = c->key[1]; key3[2] = c->key[2]; *count = (int)c->points.size();
}
int orc_map_get_local(const OrcMap* m, const double* T, int cells_xy, int cells_z, float* xyzi, int cap) {  // src/map.cc:141-189
  int x = (int)T[3];
  int voxel_x = int(std::floor(x * m->inv_xy) * m->xy + m->xy_half);
  int y = (int)T[7];
  int voxel_y = int(std::floor(y * m->inv_xy) * m->xy + m->xy_half);
  int z = (int)T[11];
  int voxel_z = int(std::floor(z * m->inv_z) * m->zs + m->z_half);
  std::vector<P4> total;
  int init_x = voxel_x - cells_xy * m->xy, end_x = voxel_x + cells_xy * m->xy;
  int init_y = voxel_y - cells_xy * m->xy, end_y = voxel_y + cells_xy * m->xy;
  for (int i = init_x; i <= end_x; i += m->xy)
    for (int j = init_y; j <= end_y; j += m->xy) {
      auto it = m->cells.find(OrcMap::Key{i, j, voxel_z});
      if (it != m->cells.end()) total.insert(total.end(), it->second->points.begin(), it->second->points.end());
    }
  int init_z = voxel_z - cells_z * m->xy, end_z = voxel_z + cells_z * m->xy;
  for (int i = init_z; i <= end_z; i += m->zs) {
    auto it = m->cells.find(OrcMap::Key{voxel_x, voxel_y, i});
    if (it != m->cells.end()) total.insert(total.end(), it->second->points.begin(), it->second->points.end());
  }
  int w = std::min((int)total.size(), cap);
  if (xyzi) std::memcpy(xyzi, total.data(), (size_t)w * 16);
  return (int)total.size();
}

long orc_run_sequence(const OrcParams* p, const float* pts, const int32_t* npts, int nframes, int stride_f,
                      int width, int height, double* poses_out, double* stage_us) {
  OrcOdom* od = orc_odom_create(p);
  long total_edges = 0; size_t pos = 0;
  const int cap = p->scan_lines * p->scan_regions * (p->edges_per_region + 1);
  std::vector<float> edges((size_t)cap * 4);
  for (int f = 0; f < nframes; ++f) {
    double tt[2] = {0, 0};
    int E = orc_extract_scan(p, pts + pos * stride_f, npts[f], stride_f, width, height, edges.data(), cap, tt);
    pos += (size_t)npts[f];
    OrcFrameDiag D;
    orc_odom_process(od, edges.data(), E, poses_out ? poses_out + 16 * f : nullptr, &D);
    if (stage_us) { stage_us[0] += tt[0]; stage_us[1] += tt[1]; stage_us[2] += D.times_us[0] + D.times_us[1]; stage_us[3] += D.times_us[2]; stage_us[4] += D.times_us[3]; }
    total_edges += E;
  }
  orc_odom_destroy(od);
  return total_edges;
}

}  // extern "C"
