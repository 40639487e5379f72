// refshim — TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// The subset of PCL 1.10 (and of FLANN behind pcl::KdTreeFLANN) that the reference's sources use, so
// that they compile UNMODIFIED: PointXYZI (32-byte layout), PointCloud, copyPointCloud,
// transformPointCloud (double matrix, generic path), ExtractIndices, VoxelGrid, KdTreeFLANN::
// nearestKSearch (exact k-NN under flann::L2_Simple<float>), toROSMsg / fromROSMsg.  PCL and FLANN are
// third-party dependencies, not vendored in the reference and not installed here; their published
// algorithms are restated.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "refshim_eigen.h"
#include "refshim_ros.h"

namespace pcl {

struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };

// pcl::PointXYZI: float data[4] (x, y, z, 1) + float data_c[4] (intensity, pad), 16-byte aligned
struct alignas(16) PointXYZI {
  float x, y, z, data3;
  float intensity, pad_[3];
  PointXYZI() : x(0.f), y(0.f), z(0.f), data3(1.f), intensity(0.f), pad_{0.f, 0.f, 0.f} {}
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is 32 bytes");

struct PointIndices { PCLHeader header; std::vector<int> indices; typedef std::shared_ptr<PointIndices> Ptr; typedef std::shared_ptr<PointIndices const> ConstPtr; };

template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT> > Ptr;
  typedef std::shared_ptr<const PointCloud<PointT> > ConstPtr;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = 0; height = 0; }
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
  bool isOrganized() const { return height > 1; }
  const PointT& at(int column, int row) const {
    if (height > 1) return points.at((size_t)row * width + column);
    throw std::runtime_error("Can't use 2D indexing with an unorganized point cloud");
  }
  PointT& at(int column, int row) {
    if (height > 1) return points.at((size_t)row * width + column);
    throw std::runtime_error("Can't use 2D indexing with an unorganized point cloud");
  }
  const PointT& operator[](size_t i) const { return points[i]; }
  PointT& operator[](size_t i) { return points[i]; }
  PointCloud& operator+=(const PointCloud& rhs) {
    if (rhs.header.stamp > header.stamp) header.stamp = rhs.header.stamp;
    points.insert(points.end(), rhs.points.begin(), rhs.points.end());
    width = (uint32_t)points.size(); height = 1;
    is_dense = is_dense && rhs.is_dense;
    return *this;
  }
};

template <typename PointT>
inline void copyPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out) { out = in; }

// pcl/common/impl/transforms.hpp, Eigen::Matrix<Scalar,4,4> overload: every row evaluated in Scalar,
// left to right, result rounded to float; data[3] = 1; other fields copied.
template <typename PointT, typename Scalar>
inline void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix<Scalar, 4, 4>& tf, bool /*copy_all_fields*/ = true) {
  if (&in != &out) {
    out.header = in.header; out.is_dense = in.is_dense; out.width = in.width; out.height = in.height;
    out.points.reserve(in.points.size());
    out.points.assign(in.points.begin(), in.points.end());
  }
  for (size_t i = 0; i < out.points.size(); ++i) {
    const PointT& s = in.points[i];
    if (!in.is_dense && (!std::isfinite(s.x) || !std::isfinite(s.y) || !std::isfinite(s.z))) continue;
    const Scalar p[3] = {s.x, s.y, s.z};
    const float tx = static_cast<float>(tf(0, 0) * p[0] + tf(0, 1) * p[1] + tf(0, 2) * p[2] + tf(0, 3));
    const float ty = static_cast<float>(tf(1, 0) * p[0] + tf(1, 1) * p[1] + tf(1, 2) * p[2] + tf(1, 3));
    const float tz = static_cast<float>(tf(2, 0) * p[0] + tf(2, 1) * p[1] + tf(2, 2) * p[2] + tf(2, 3));
    out.points[i].x = tx; out.points[i].y = ty; out.points[i].z = tz; out.points[i].data3 = 1.f;
  }
}

template <typename PointT>
class PCLBase {
 public:
  typedef typename PointCloud<PointT>::Ptr PointCloudPtr;
  typedef typename PointCloud<PointT>::ConstPtr PointCloudConstPtr;
  virtual ~PCLBase() {}
  void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  void setIndices(const PointIndices::ConstPtr& ind) { indices_ = ind; }
 protected:
  PointCloudConstPtr input_;
  PointIndices::ConstPtr indices_;
};

// pcl::Filter::filter handles output == input by filtering into a temporary.
template <typename PointT>
class Filter : public PCLBase<PointT> {
 public:
  void filter(PointCloud<PointT>& output) {
    if (!this->input_) return;
    if (&output == this->input_.get()) {
      PointCloud<PointT> tmp;
      applyFilter(tmp);
      tmp.header = this->input_->header;
      output = tmp;
    } else {
      output.header = this->input_->header;
      applyFilter(output);
    }
  }
 protected:
  virtual void applyFilter(PointCloud<PointT>& output) = 0;
};

// pcl::ExtractIndices with setNegative: keeps (or drops) the indexed points, input order preserved.
template <typename PointT>
class ExtractIndices : public Filter<PointT> {
 public:
  void setNegative(bool negative) { negative_ = negative; }
 protected:
  void applyFilter(PointCloud<PointT>& output) override {
    const PointCloud<PointT>& in = *this->input_;
    std::vector<char> marked(in.points.size(), 0);
    if (this->indices_) for (int i : this->indices_->indices) if (i >= 0 && (size_t)i < marked.size()) marked[i] = 1;
    output.points.clear();
    if (!negative_) {
      if (this->indices_) for (int i : this->indices_->indices) output.points.push_back(in.points[i]);
    } else {
      for (size_t i = 0; i < in.points.size(); ++i) if (!marked[i]) output.points.push_back(in.points[i]);
    }
    output.width = (uint32_t)output.points.size(); output.height = 1; output.is_dense = in.is_dense;
  }
 private:
  bool negative_ = false;
};

// pcl::VoxelGrid<PointXYZI>::applyFilter (voxel_grid.hpp), downsample_all_data = true, no filter field,
// min_points_per_voxel = 0: bounding box -> per-point voxel index -> std::sort by index (the comparator
// looks at the index only, as cloud_point_index_idx::operator< does) -> one centroid per voxel in
// ascending index order (CentroidPoint: float sums of x, y, z, intensity divided by the count).
template <typename PointT>
class VoxelGrid : public Filter<PointT> {
 public:
  void setLeafSize(float lx, float ly, float lz) {
    leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz;
    for (int k = 0; k < 3; ++k) inv_[k] = 1.0f / leaf_[k];
  }
 protected:
  struct cloud_point_index_idx {
    unsigned int idx, cloud_point_index;
    bool operator<(const cloud_point_index_idx& p) const { return idx < p.idx; }
  };
  void applyFilter(PointCloud<PointT>& output) override {
    const PointCloud<PointT>& in = *this->input_;
    output.height = 1; output.is_dense = true;
    output.points.clear();
    if (in.points.empty()) { output.width = 0; return; }
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float mx[3] = {-mn[0], -mn[1], -mn[2]};
    for (const PointT& p : in.points) {   // getMinMax3D
      if (!in.is_dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))) continue;
      mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
      mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    const int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv_[0]) + 1;
    const int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv_[1]) + 1;
    const int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv_[2]) + 1;
    if (dx * dy * dz > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
      ROS_WARN("[pcl::VoxelGrid::applyFilter] Leaf size is too small for the input dataset. Integer indices would overflow.");
      output = in;
      return;
    }
    int min_b[3], max_b[3], div_b[3];
    for (int k = 0; k < 3; ++k) {
      min_b[k] = static_cast<int>(std::floor(mn[k] * inv_[k]));
      max_b[k] = static_cast<int>(std::floor(mx[k] * inv_[k]));
      div_b[k] = max_b[k] - min_b[k] + 1;
    }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    std::vector<cloud_point_index_idx> index_vector;
    index_vector.reserve(in.points.size());
    for (unsigned i = 0; i < in.points.size(); ++i) {
      const PointT& p = in.points[i];
      if (!in.is_dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))) continue;
      const int ijk0 = static_cast<int>(std::floor(p.x * inv_[0]) - static_cast<float>(min_b[0]));
      const int ijk1 = static_cast<int>(std::floor(p.y * inv_[1]) - static_cast<float>(min_b[1]));
      const int ijk2 = static_cast<int>(std::floor(p.z * inv_[2]) - static_cast<float>(min_b[2]));
      index_vector.push_back(cloud_point_index_idx{static_cast<unsigned>(ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2]), i});
    }
    std::sort(index_vector.begin(), index_vector.end(), std::less<cloud_point_index_idx>());
    size_t first = 0;
    while (first < index_vector.size()) {
      size_t last = first + 1;
      while (last < index_vector.size() && index_vector[last].idx == index_vector[first].idx) ++last;
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      for (size_t u = first; u < last; ++u) {
        const PointT& p = in.points[index_vector[u].cloud_point_index];
        sx += p.x; sy += p.y; sz += p.z; si += p.intensity;
      }
      const float n = static_cast<float>(last - first);
      PointT c;
      c.x = sx / n; c.y = sy / n; c.z = sz / n; c.intensity = si / n;
      output.points.push_back(c);
      first = last;
    }
    output.width = static_cast<uint32_t>(output.points.size());
  }
 private:
  float leaf_[3] = {0.f, 0.f, 0.f}, inv_[3] = {0.f, 0.f, 0.f};
};

// pcl::KdTreeFLANN<PointXYZI>::nearestKSearch: exact k nearest neighbours under flann::L2_Simple<float>
// (x, y, z only; result += diff*diff in that order), ascending squared distances.  The traversal of
// FLANN's KDTreeSingleIndex is not restated: an exhaustive scan gives the same answer except for the
// order of exactly equal distances (FLANN keeps tree traversal order; here lower index first).
// Non-finite input points are left out of the index, as PCL does.
template <typename PointT>
class KdTreeFLANN {
 public:
  typedef std::shared_ptr<KdTreeFLANN<PointT> > Ptr;
  typedef typename PointCloud<PointT>::ConstPtr PointCloudConstPtr;
  void setInputCloud(const PointCloudConstPtr& cloud) {
    input_ = cloud;
    index_.clear();
    for (size_t i = 0; i < cloud->points.size(); ++i) {
      const PointT& p = cloud->points[i];
      if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) index_.push_back((int)i);
    }
  }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    if (k > (int)index_.size()) k = (int)index_.size();
    k_indices.assign(k, 0);
    k_sqr_distances.assign(k, 0.f);
    if (k == 0) return 0;
    std::vector<std::pair<float, int> > best;   // ascending (d2, index), at most k entries
    best.reserve(k + 1);
    for (int i : index_) {
      const PointT& p = input_->points[i];
      float result = 0.f, diff;
      diff = q.x - p.x; result += diff * diff;
      diff = q.y - p.y; result += diff * diff;
      diff = q.z - p.z; result += diff * diff;
      if ((int)best.size() == k && !(result < best.back().first)) continue;   // KNNSimpleResultSet: dist >= worst rejected
      std::pair<float, int> e(result, i);
      auto pos = std::upper_bound(best.begin(), best.end(), e, [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first; });
      best.insert(pos, e);   // equal distances: after the existing ones
      if ((int)best.size() > k) best.pop_back();
    }
    for (int j = 0; j < k; ++j) { k_indices[j] = best[j].second; k_sqr_distances[j] = best[j].first; }
    return k;
  }
 private:
  PointCloudConstPtr input_;
  std::vector<int> index_;
};

// pcl_conversions: PointXYZI <-> sensor_msgs/PointCloud2 (fields x, y, z @0,4,8 and intensity @16)
template <typename PointT>
inline void toROSMsg(const PointCloud<PointT>& cloud, sensor_msgs::PointCloud2& msg) {
  msg.height = cloud.height ? cloud.height : 1;
  msg.width = cloud.height ? cloud.width : (uint32_t)cloud.points.size();
  msg.fields.clear();
  const char* names[4] = {"x", "y", "z", "intensity"};
  const uint32_t offs[4] = {0, 4, 8, 16};
  for (int k = 0; k < 4; ++k) { sensor_msgs::PointField f; f.name = names[k]; f.offset = offs[k]; f.datatype = sensor_msgs::PointField::FLOAT32; f.count = 1; msg.fields.push_back(f); }
  msg.is_bigendian = false;
  msg.point_step = sizeof(PointT);
  msg.row_step = msg.point_step * msg.width;
  msg.is_dense = cloud.is_dense;
  msg.data.resize(cloud.points.size() * sizeof(PointT));
  if (!cloud.points.empty()) std::memcpy(msg.data.data(), cloud.points.data(), msg.data.size());
  msg.header.frame_id = cloud.header.frame_id;
  msg.header.stamp.fromSec(1e-6 * (double)cloud.header.stamp);
}

template <typename PointT>
inline void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointT>& cloud) {
  int off[4] = {-1, -1, -1, -1};
  const char* names[4] = {"x", "y", "z", "intensity"};
  for (const auto& f : msg.fields)
    for (int k = 0; k < 4; ++k)
      if (f.name == names[k] && f.datatype == sensor_msgs::PointField::FLOAT32 && f.count == 1) off[k] = (int)f.offset;
  cloud.width = msg.width; cloud.height = msg.height; cloud.is_dense = msg.is_dense;
  cloud.header.frame_id = msg.header.frame_id;
  cloud.points.assign((size_t)msg.width * msg.height, PointT());
  for (uint32_t r = 0; r < msg.height; ++r)
    for (uint32_t c = 0; c < msg.width; ++c) {
      const uint8_t* src = msg.data.data() + (size_t)r * msg.row_step + (size_t)c * msg.point_step;
      PointT& p = cloud.points[(size_t)r * msg.width + c];
      if (off[0] >= 0) std::memcpy(&p.x, src + off[0], 4);
      if (off[1] >= 0) std::memcpy(&p.y, src + off[1], 4);
      if (off[2] >= 0) std::memcpy(&p.z, src + off[2], 4);
      if (off[3] >= 0) std::memcpy(&p.intensity, src + off[3], 4);
    }
}

}  // namespace pcl
