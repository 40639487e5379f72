// Host facade of the B200 hot path: the reference's class surface (include/liodom/*.h) on top
// of the C ABI (include/liodom_b200.h).  Everything numeric happens in libliodom_b200.so; this
// file is queues, ownership and parameter plumbing.  Errors follow the reference's convention:
// logged, processing continues.
#include <liodom/feature_extractor.h>
#include <liodom/laser_odometry.h>
#include <liodom/map.h>

#include "../../include/liodom_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <thread>

namespace liodom {

static bool verbose() { static const bool v = std::getenv("LIODOM_VERBOSE") != nullptr; return v; }
#define LIODOM_INFO(...) do { if (verbose()) { std::fprintf(stderr, "[liodom] " __VA_ARGS__); std::fprintf(stderr, "\n"); } } while (0)
#define LIODOM_ERROR(...) do { std::fprintf(stderr, "[liodom][error] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)

// ---------------------------------------------------------------------------------------------
// Params (src/params.cc:24-110)
// ---------------------------------------------------------------------------------------------
Params* Params::instance_{nullptr};
std::mutex Params::instance_mutex_;

Params* Params::getInstance() {
  std::lock_guard<std::mutex> lock(instance_mutex_);
  if (!instance_) instance_ = new Params();
  return instance_;
}

Params::Params()
    : min_range_(3.0), max_range_(75.0), lidar_type_(0), scan_lines_(64), scan_regions_(8), edges_per_region_(10),
      min_points_per_scan_(90), local_map_size_(5), use_imu_(false), filter_local_map_(false), mapping_(false),
      save_results_(false), publish_tf_(true), results_dir_("~/"), fixed_frame_("odom"), base_frame_("base_link"), laser_frame_("") {}

void Params::readParams(const NodeHandle& nh) {
  nh.param("min_range", min_range_, 3.0);
  nh.param("max_range", max_range_, 75.0);
  nh.param("lidar_type", lidar_type_, 0);
  nh.param("scan_lines", scan_lines_, 64);
  nh.param("scan_regions", scan_regions_, 8);
  nh.param("edges_per_region", edges_per_region_, 10);
  min_points_per_scan_ = (size_t)(scan_regions_ * edges_per_region_ + 10);
  nh.param("save_results", save_results_, false);
  nh.param<std::string>("save_results_dir", results_dir_, "~/");
  nh.param<std::string>("fixed_frame", fixed_frame_, "odom");
  nh.param<std::string>("base_frame", base_frame_, "base_link");
  nh.param<std::string>("laser_frame", laser_frame_, "");
  int pframes = 5;
  nh.param("prev_frames", pframes, 5);
  local_map_size_ = (size_t)pframes;
  nh.param("use_imu", use_imu_, false);
  nh.param("filter_local_map", filter_local_map_, false);
  nh.param("mapping", mapping_, false);
  nh.param("publish_tf", publish_tf_, true);
  LIODOM_INFO("range [%.2f, %.2f], lidar_type %d, scan_lines %d, regions %d, edges/region %d, window %zu, mapping %d",
              min_range_, max_range_, lidar_type_, scan_lines_, scan_regions_, edges_per_region_, local_map_size_, (int)mapping_);
}

// ---------------------------------------------------------------------------------------------
// SharedData (src/shared_data.cc:24-117)
// ---------------------------------------------------------------------------------------------
SharedData* SharedData::instance_{nullptr};
std::mutex SharedData::instance_mutex_;

SharedData* SharedData::getInstance() {
  std::lock_guard<std::mutex> lock(instance_mutex_);
  if (!instance_) instance_ = new SharedData();
  return instance_;
}
// Both queues share one shape: the cloud pointer and its header travel together, nothing is copied.
template <typename F>
static void fifo_push(F& q, const PointCloud::Ptr& cloud, const Header& header) {
  {
    std::lock_guard<std::mutex> lock(q.m);
    q.items.push(cloud); q.headers.push(header);
  }
  q.cv.notify_one();
}
template <typename F>
static void fifo_wait(F& q, int ms) {
  std::unique_lock<std::mutex> lock(q.m);
  q.cv.wait_for(lock, std::chrono::milliseconds(ms), [&] { return !q.items.empty(); });
}
template <typename F>
static bool fifo_pop(F& q, PointCloud::Ptr& cloud, Header& header) {
  std::lock_guard<std::mutex> lock(q.m);
  if (q.items.empty()) return false;
  cloud = q.items.front(); q.items.pop();
  header = q.headers.front(); q.headers.pop();
  return true;
}
void SharedData::pushPointCloud(const PointCloud::Ptr& pc_in, const Header& header) { fifo_push(scans_, pc_in, header); }
bool SharedData::popPointCloud(PointCloud::Ptr& pc_out, Header& header) { return fifo_pop(scans_, pc_out, header); }
void SharedData::pushFeatures(const PointCloud::Ptr& feat_in, Header& header) { fifo_push(feats_, feat_in, header); }
bool SharedData::popFeatures(PointCloud::Ptr& feat_out, Header& header) { return fifo_pop(feats_, feat_out, header); }
void SharedData::waitPointCloud(int ms) { fifo_wait(scans_, ms); }
void SharedData::waitFeatures(int ms) { fifo_wait(feats_, ms); }
void SharedData::setLocalMap(const PointCloud::Ptr& map_in) {
  std::lock_guard<std::mutex> lock(map_mutex_);
  *local_map_ = *map_in;   // pcl::copyPointCloud: deep copy
}
void SharedData::getLocalMap(PointCloud::Ptr& map_out) {
  std::lock_guard<std::mutex> lock(map_mutex_);
  *map_out = *local_map_;
}
void SharedData::setLastIMUOri(Quaterniond& imu_ori) { std::lock_guard<std::mutex> lock(imu_mutex_); last_imu_ = imu_ori; }
void SharedData::getLastIMUOri(Quaterniond& imu_ori) { std::lock_guard<std::mutex> lock(imu_mutex_); imu_ori = last_imu_; }

// ---------------------------------------------------------------------------------------------
// Stats (src/stats.cc:24-132)
// ---------------------------------------------------------------------------------------------
Stats* Stats::instance_{nullptr};
std::mutex Stats::instance_mutex_;

Stats* Stats::getInstance() {
  std::lock_guard<std::mutex> lock(instance_mutex_);
  if (!instance_) instance_ = new Stats();
  return instance_;
}
static double whole_ms(const Clock::time_point& a, const Clock::time_point& b) {
  return (double)std::chrono::duration_cast<std::chrono::milliseconds>(b - a).count();
}
void Stats::addPose(const Matrix4d& pose) { poses_.push_back(pose); }
void Stats::addFeatureExtractionTime(const Clock::time_point& start, const Clock::time_point& end) { extract_ms_.push_back(whole_ms(start, end)); }
void Stats::addLaserOdometryTime(const Clock::time_point& start, const Clock::time_point& end) { odom_ms_.push_back(whole_ms(start, end)); }
void Stats::addNumOfFeats(const size_t& nfeats) { nfeats_.push_back(nfeats); }
void Stats::startFrame(const Clock::time_point& start) { std::lock_guard<std::mutex> lock(frame_mutex_); pending_starts_.push(start); }
void Stats::stopFrame(const Clock::time_point& stop) {
  std::lock_guard<std::mutex> lock(frame_mutex_);
  if (!pending_starts_.empty()) {
    const Clock::time_point start = pending_starts_.front();
    pending_starts_.pop();
    frame_ms_.push_back(whole_ms(start, stop));
  }
}
void Stats::clear() {
  std::lock_guard<std::mutex> lock(frame_mutex_);
  poses_.clear(); extract_ms_.clear(); odom_ms_.clear(); nfeats_.clear(); frame_ms_.clear();
  while (!pending_starts_.empty()) pending_starts_.pop();
}
template <typename V> static void write_column(const std::string& path, const V& v) {
  std::ofstream f(path.c_str(), std::ios::out | std::ios::trunc);
  for (size_t i = 0; i < v.size(); ++i) f << v[i] << std::endl;
}
void Stats::writeResults(const std::string& dir) {
  {  // KITTI format: the first three rows of every pose on one line, default stream precision
    std::ofstream f((dir + "poses.txt").c_str(), std::ios::out | std::ios::trunc);
    for (size_t k = 0; k < poses_.size(); ++k)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
          f << poses_[k](i, j);
          if (i == 2 && j == 3) f << std::endl; else f << " ";
        }
  }
  write_column(dir + "feat_ext_times.txt", extract_ms_);
  write_column(dir + "laser_odom_times.txt", odom_ms_);
  write_column(dir + "nfeats.txt", nfeats_);
  write_column(dir + "frame_times.txt", frame_ms_);
}

// ---------------------------------------------------------------------------------------------
// context helpers
// ---------------------------------------------------------------------------------------------
static std::shared_ptr<liodom_ctx> make_ctx(const Params* params, int max_points, int prev_frames, int max_received) {
  liodom_params p;
  liodom_default_params(&p);
  p.min_range = params->min_range_; p.max_range = params->max_range_; p.lidar_type = params->lidar_type_;
  p.scan_lines = params->scan_lines_; p.scan_regions = params->scan_regions_; p.edges_per_region = params->edges_per_region_;
  p.prev_frames = prev_frames; p.filter_local_map = params->filter_local_map_ ? 1 : 0; p.mapping = params->mapping_ ? 1 : 0;
  p.max_points = max_points; p.max_received_map = max_received; p.use_imu = params->use_imu_ ? 1 : 0;
  int device = 0;
  if (const char* d = std::getenv("LIODOM_DEVICE")) device = std::atoi(d);
  liodom_ctx* c = nullptr;
  const int rc = liodom_ctx_create(&p, 1, device, &c);
  if (rc != LIODOM_OK) {
    // "Invalid scan lines" / "Incorrect Lidar type" are ROS_ERROR_ONCE in the reference
    // (src/feature_extractor.cc:150,177): log and carry on without output.
    LIODOM_ERROR("liodom_ctx_create failed (%d): %s", rc, liodom_last_error(nullptr));
    return std::shared_ptr<liodom_ctx>();
  }
  return std::shared_ptr<liodom_ctx>(c, [](liodom_ctx* q) { liodom_ctx_destroy(q); });
}

static void cloud_from_xyzi(const float* xyzi, int n, PointCloud* pc) {
  pc->points.resize((size_t)n);
  for (int i = 0; i < n; ++i) {
    Point& q = pc->points[(size_t)i];
    q.x = xyzi[4 * i]; q.y = xyzi[4 * i + 1]; q.z = xyzi[4 * i + 2]; q.intensity = xyzi[4 * i + 3];
  }
  pc->width = (uint32_t)n; pc->height = 1; pc->is_dense = true;
}
static void xyzi_from_cloud(const PointCloud& pc, std::vector<float>* out) {
  out->resize(pc.points.size() * 4);
  for (size_t i = 0; i < pc.points.size(); ++i) {
    const Point& q = pc.points[i];
    (*out)[4 * i] = q.x; (*out)[4 * i + 1] = q.y; (*out)[4 * i + 2] = q.z; (*out)[4 * i + 3] = q.intensity;
  }
}

// ---------------------------------------------------------------------------------------------
// FeatureExtractor (src/feature_extractor.cc:24-82)
// ---------------------------------------------------------------------------------------------
FeatureExtractor::FeatureExtractor(const NodeHandle& nh)
    : nh_(nh), sdata(SharedData::getInstance()), stats(Stats::getInstance()), params(Params::getInstance()) {
#ifdef LIODOM_FACADE_USE_PCL
  pc_edges_pub_ = nh_.advertise<sensor_msgs::PointCloud2>("edges", 10);
#endif
}
FeatureExtractor::FeatureExtractor(const FeatureExtractor& o)
    : nh_(o.nh_),
#ifdef LIODOM_FACADE_USE_PCL
      pc_edges_pub_(o.pc_edges_pub_),
#endif
      sdata(o.sdata), stats(o.stats), params(o.params), ctx_(o.ctx_), edges_cb_(o.edges_cb_), ctx_points_(o.ctx_points_) {}
FeatureExtractor::~FeatureExtractor() {}

bool FeatureExtractor::ensureContext(size_t npoints) {
  if (ctx_ && npoints <= ctx_points_) return true;
  size_t cap = 131072;
  while (cap < npoints) cap <<= 1;
  ctx_ = make_ctx(params, (int)cap, (int)params->local_map_size_, 0);
  ctx_points_ = ctx_ ? cap : 0;
  return (bool)ctx_;
}

// ---- sensor_msgs/PointCloud2 -> liodom::Point (pcl::fromROSMsg, src/liodom_node.cc:43-44, :62-63) ----
bool cloudLayoutFromFields(const PointCloud2& msg, liodom_cloud_layout* layout) {
  if (!layout || msg.is_bigendian) return false;
  int off[4] = {-1, -1, -1, -1};
  static const char* names[4] = {"x", "y", "z", "intensity"};
  for (const PointField& f : msg.fields)
    for (int k = 0; k < 4; ++k)
      if (off[k] < 0 && f.name == names[k] && f.datatype == PointField::FLOAT32 && f.count == 1) off[k] = (int)f.offset;
  if (off[0] < 0 || off[1] < 0 || off[2] < 0) return false;
  layout->point_step = (int)msg.point_step; layout->row_step = (int)msg.row_step;
  layout->off_x = off[0]; layout->off_y = off[1]; layout->off_z = off[2]; layout->off_intensity = off[3];
  layout->is_bigendian = 0;
  return true;
}

#ifndef LIODOM_FACADE_USE_PCL   // with PCL present pcl::fromROSMsg does this (src/liodom_node.cc:43-44)
bool fromROSMsg(const PointCloud2& msg, PointCloud& cloud) {
  liodom_cloud_layout lay;
  cloud.clear(); cloud.raw.reset();
  if (!cloudLayoutFromFields(msg, &lay)) { LIODOM_ERROR("fromROSMsg: no FLOAT32 x/y/z fields (or a big-endian message)"); return false; }
  const size_t row = msg.row_step ? msg.row_step : (size_t)msg.width * msg.point_step;
  if (msg.data.size() < (size_t)msg.height * row) { LIODOM_ERROR("fromROSMsg: data shorter than height * row_step"); return false; }
  cloud.points.resize((size_t)msg.width * msg.height);
  for (uint32_t r = 0; r < msg.height; ++r)
    for (uint32_t c = 0; c < msg.width; ++c) {
      const uint8_t* src = msg.data.data() + r * row + (size_t)c * msg.point_step;
      Point& p = cloud.points[(size_t)r * msg.width + c];
      std::memcpy(&p.x, src + lay.off_x, 4); std::memcpy(&p.y, src + lay.off_y, 4); std::memcpy(&p.z, src + lay.off_z, 4);
      if (lay.off_intensity >= 0) std::memcpy(&p.intensity, src + lay.off_intensity, 4);
    }
  cloud.width = msg.width; cloud.height = msg.height; cloud.is_dense = msg.is_dense; cloud.header = msg.header;
  return true;
}

bool fromROSMsgDeferred(const PointCloud2::ConstPtr& msg, PointCloud& cloud) {
  liodom_cloud_layout lay;
  cloud.clear(); cloud.raw.reset();
  if (!msg || !cloudLayoutFromFields(*msg, &lay)) { LIODOM_ERROR("fromROSMsgDeferred: no FLOAT32 x/y/z fields (or a big-endian message)"); return false; }
  cloud.raw = msg;
  cloud.width = msg->width; cloud.height = msg->height; cloud.is_dense = msg->is_dense; cloud.header = msg->header;
  return true;
}
#endif

bool FeatureExtractor::extract(const PointCloud::Ptr& pc_curr, PointCloud::Ptr& pc_edges) {
  if (!ensureContext(pc_curr->size())) return false;
  const int cap = liodom_max_edges(ctx_.get());
  std::vector<float> edges((size_t)cap * 4);
  int ne = 0;
  const int w = params->lidar_type_ == 1 ? (int)pc_curr->width : 0, h = params->lidar_type_ == 1 ? (int)pc_curr->height : 0;
  int rc;
#ifndef LIODOM_FACADE_USE_PCL
  if (pc_curr->raw) {   // message bytes straight to the device (no host-side decode)
    liodom_cloud_layout lay;
    if (!cloudLayoutFromFields(*pc_curr->raw, &lay)) { LIODOM_ERROR("extract: unusable PointCloud2 field list"); return false; }
    const int ww = (w > 0 || lay.row_step) ? (int)pc_curr->raw->width : 0, hh = (h > 0 || lay.row_step) ? (int)pc_curr->raw->height : 0;
    rc = liodom_extract_layout(ctx_.get(), 0, pc_curr->raw->data.data(), (int)pc_curr->size(), &lay, ww, hh, edges.data(), &ne, nullptr, nullptr);
  } else
#endif
    rc = liodom_extract(ctx_.get(), 0, pc_curr->points.data(), (int)pc_curr->size(), (int)sizeof(Point), w, h,
                        edges.data(), &ne, nullptr, nullptr, nullptr);
  if (rc != LIODOM_OK) { LIODOM_ERROR("liodom_extract failed (%d): %s", rc, liodom_last_error(ctx_.get())); return false; }
  cloud_from_xyzi(edges.data(), ne, pc_edges.get());
  return true;
}

void FeatureExtractor::operator()(std::atomic<bool>& running) {
  while (running) {
    PointCloud::Ptr pc_curr(new PointCloud);
    Header pc_header;
    if (sdata->popPointCloud(pc_curr, pc_header)) {
      PointCloud::Ptr pc_edges(new PointCloud);
      // The reference starts its timer after splitPointCloud (src/feature_extractor.cc:53-55); split and
      // extraction are one enqueue here, so the span covers both.
      const auto start_t = Clock::now();
      extract(pc_curr, pc_edges);
      const auto end_t = Clock::now();
      stats->addFeatureExtractionTime(start_t, end_t);
      stats->addNumOfFeats(pc_edges->size());
#ifdef LIODOM_FACADE_USE_PCL
      {   // publishing detected edges (src/feature_extractor.cc:71-74)
        sensor_msgs::PointCloud2 edges_msg;
        pcl::toROSMsg(*pc_edges, edges_msg);
        edges_msg.header = pc_header;
        pc_edges_pub_.publish(edges_msg);
      }
#endif
      if (edges_cb_) edges_cb_(pc_header, pc_edges);
      sdata->pushFeatures(pc_edges, pc_header);
    }
    sdata->waitPointCloud(2);   // the reference sleeps 2 ms here (src/feature_extractor.cc:80); see shared_data.h
  }
}

// ---------------------------------------------------------------------------------------------
// LocalMapManager (src/laser_odometry.cc:24-69)
// ---------------------------------------------------------------------------------------------
LocalMapManager::LocalMapManager(const size_t max_frames) : max_nframes_(max_frames) {}
LocalMapManager::~LocalMapManager() {}

bool LocalMapManager::ensureContext(size_t frame_points) {
  if (ctx_ && frame_points <= frame_cap_) return true;
  if (ctx_) { LIODOM_ERROR("LocalMapManager: frame of %zu points exceeds the slab capacity %zu", frame_points, frame_cap_); return false; }
  // slab capacity = scan_lines * scan_regions * (edges_per_region + 1) of the current Params
  ctx_ = make_ctx(Params::getInstance(), 2048, (int)std::max<size_t>(max_nframes_, 1), 0);
  frame_cap_ = ctx_ ? (size_t)liodom_max_edges(ctx_.get()) : 0;
  return ctx_ && frame_points <= frame_cap_;
}
void LocalMapManager::addPointCloud(const PointCloud::Ptr& pc) {
  if (!ensureContext(pc->size())) return;
  std::vector<float> buf;
  xyzi_from_cloud(*pc, &buf);
  const int rc = liodom_lmap_add(ctx_.get(), 0, buf.data(), (int)pc->size());
  if (rc != LIODOM_OK) LIODOM_ERROR("liodom_lmap_add failed (%d): %s", rc, liodom_last_error(ctx_.get()));
}
size_t LocalMapManager::getLocalMap(PointCloud::Ptr& map) {
  if (!ctx_) { map->clear(); return 0; }
  int n = 0, nf = 0;
  liodom_lmap_get(ctx_.get(), 0, nullptr, 0, &n, &nf);
  std::vector<float> buf((size_t)std::max(n, 1) * 4);
  const int rc = liodom_lmap_get(ctx_.get(), 0, buf.data(), n, &n, &nf);
  if (rc != LIODOM_OK) { LIODOM_ERROR("liodom_lmap_get failed (%d): %s", rc, liodom_last_error(ctx_.get())); return 0; }
  cloud_from_xyzi(buf.data(), n, map.get());
  return (size_t)nf;
}
void LocalMapManager::setMaxFrames(const size_t max_nframes) {
  max_nframes_ = max_nframes;
  if (ctx_) {
    const int rc = liodom_lmap_set_max_frames(ctx_.get(), 0, (int)max_nframes);
    if (rc != LIODOM_OK) LIODOM_ERROR("liodom_lmap_set_max_frames failed (%d): %s", rc, liodom_last_error(ctx_.get()));
  }
}

// ---------------------------------------------------------------------------------------------
// LaserOdometer (src/laser_odometry.cc:71-272)
// ---------------------------------------------------------------------------------------------
static double wall_now_secs() { return std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count(); }

LaserOdometer::LaserOdometer(const NodeHandle& nh)
    : nh_(nh), init_(false), odom_(Isometry3d::Identity()), sdata(SharedData::getInstance()), stats(Stats::getInstance()),
      params(Params::getInstance()), prev_odom_(Isometry3d::Identity()), laser_to_base_(Isometry3d::Identity()) {
  for (int i = 0; i < 5; ++i) { in_freqs_[i] = 20.0; out_freqs_[i] = 20.0; }   // 100 / 5 (src/laser_odometry.cc:83-90)
  last_in_time_secs_ = last_out_time_secs_ = wall_now_secs();
#ifdef LIODOM_FACADE_USE_PCL
  odom_pub_ = nh_.advertise<nav_msgs::Odometry>("odom", 10);
  twist_pub_ = nh_.advertise<geometry_msgs::TwistStamped>("twist", 10);
#endif
}
LaserOdometer::LaserOdometer(const LaserOdometer& o)
    : nh_(o.nh_),
#ifdef LIODOM_FACADE_USE_PCL
      odom_pub_(o.odom_pub_), twist_pub_(o.twist_pub_), tf_broadcaster_(o.tf_broadcaster_),
#endif
      init_(o.init_), odom_(o.odom_), sdata(o.sdata), stats(o.stats), params(o.params), ctx_(o.ctx_), odom_cb_(o.odom_cb_),
      odom_msg_cb_(o.odom_msg_cb_), prev_odom_(o.prev_odom_), laser_to_base_(o.laser_to_base_), prev_stamp_(o.prev_stamp_),
      mean_in_freq_(o.mean_in_freq_), mean_out_freq_(o.mean_out_freq_), num_freqs_(o.num_freqs_),
      last_in_time_secs_(o.last_in_time_secs_), last_out_time_secs_(o.last_out_time_secs_), rate_warnings_(o.rate_warnings_) {
  for (int i = 0; i < 5; ++i) { in_freqs_[i] = o.in_freqs_[i]; out_freqs_[i] = o.out_freqs_[i]; }
}
LaserOdometer::~LaserOdometer() {}

bool LaserOdometer::ensureContext() {
  if (ctx_) return true;
  ctx_ = make_ctx(params, 2048, (int)params->local_map_size_, 0);   // 0: the library default capacity for the received map
  if (ctx_) { double l2b[16]; detail::pose_to16(laser_to_base_, l2b); liodom_odom_set_laser_to_base(ctx_.get(), 0, l2b); }
  return (bool)ctx_;
}

bool LaserOdometer::process(const PointCloud::Ptr& feats, const Header& header, Isometry3d* pose_out) {
  if (!ensureContext()) return false;
  if (params->mapping_) {   // computeLocalMap: latest map received from the mapping process (src/laser_odometry.cc:276-278)
    PointCloud::Ptr rec(new PointCloud);
    sdata->getLocalMap(rec);
    std::vector<float> buf;
    xyzi_from_cloud(*rec, &buf);
    const int rc = liodom_set_received_map(ctx_.get(), 0, buf.data(), (int)rec->size());
    if (rc != LIODOM_OK) LIODOM_ERROR("liodom_set_received_map failed (%d): %s", rc, liodom_last_error(ctx_.get()));
  }
  if (params->use_imu_) {   // sdata->getLastIMUOri (src/laser_odometry.cc:155-157): the override itself runs on the device
    Quaterniond imu_ori;
    sdata->getLastIMUOri(imu_ori);
    const double q[4] = {imu_ori.x(), imu_ori.y(), imu_ori.z(), imu_ori.w()};
    const int rc = liodom_odom_set_imu(ctx_.get(), 0, q);
    if (rc != LIODOM_OK) LIODOM_ERROR("liodom_odom_set_imu failed (%d): %s", rc, liodom_last_error(ctx_.get()));
  }
  std::vector<float> buf;
  xyzi_from_cloud(*feats, &buf);
  double pose16[16];
  liodom_frame_diag diag;
  const int rc = liodom_register(ctx_.get(), 0, buf.data(), (int)feats->size(), pose16, &diag);
  if (rc != LIODOM_OK) { LIODOM_ERROR("liodom_register failed (%d): %s", rc, liodom_last_error(ctx_.get())); return false; }
  if (init_) prev_odom_ = odom_;   // prev_odom_ = odom_ before the prediction (src/laser_odometry.cc:149)
  odom_ = detail::pose_from16(pose16);
  init_ = true;
  LIODOM_INFO("frame %u: %d edges, map %d, matches %d/%d", header.seq, diag.n_edges, diag.n_map[0], diag.n_matches[0], diag.n_matches[1]);
  if (pose_out) *pose_out = odom_;
  return true;
}

void LaserOdometer::setLaserToBase(const Isometry3d& laser_to_base) {
  laser_to_base_ = laser_to_base;
  if (ensureContext()) {
    double l2b[16];
    detail::pose_to16(laser_to_base, l2b);
    const int rc = liodom_odom_set_laser_to_base(ctx_.get(), 0, l2b);
    if (rc != LIODOM_OK) LIODOM_ERROR("liodom_odom_set_laser_to_base failed (%d): %s", rc, liodom_last_error(ctx_.get()));
  }
}

// Eigen::Quaterniond(Matrix3d): trace / largest-diagonal branches
static Quaterniond quat_from_rotation(const Isometry3d& Tiso) {
  const auto& T = Tiso.matrix();
  const double tr = T(0, 0) + T(1, 1) + T(2, 2);
  double v[4];
  if (tr > 0.0) {
    double t = std::sqrt(tr + 1.0);
    v[3] = 0.5 * t; t = 0.5 / t;
    v[0] = (T(2, 1) - T(1, 2)) * t; v[1] = (T(0, 2) - T(2, 0)) * t; v[2] = (T(1, 0) - T(0, 1)) * t;
  } else {
    int i = 0;
    if (T(1, 1) > T(0, 0)) i = 1;
    if (T(2, 2) > T(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(T(i, i) - T(j, j) - T(k, k) + 1.0);
    v[i] = 0.5 * t; t = 0.5 / t;
    v[3] = (T(k, j) - T(j, k)) * t; v[j] = (T(j, i) + T(i, j)) * t; v[k] = (T(k, i) + T(i, k)) * t;
  }
  return Quaterniond(v[3], v[0], v[1], v[2]);   // Eigen's (w, x, y, z) constructor
}

// tf::Matrix3x3(tf::Quaternion).getRPY (getEulerYPR, solution 1)
static void rpy_from_quat(const Quaterniond& q, double* roll, double* pitch, double* yaw) {
  const double d = q.x() * q.x() + q.y() * q.y() + q.z() * q.z() + q.w() * q.w();
  const double s = 2.0 / d;
  const double xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
  const double wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
  const double xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs, yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
  const double m00 = 1.0 - (yy + zz), m01 = xy - wz, m02 = xz + wy, m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
  const double kPi = 3.14159265358979323846;
  if (std::fabs(m20) >= 1.0) {
    *yaw = 0.0;
    if (m20 < 0.0) { *pitch = kPi / 2.0; *roll = std::atan2(m01, m02); }
    else { *pitch = -kPi / 2.0; *roll = std::atan2(-m01, -m02); }
  } else {
    *pitch = -std::asin(m20);
    const double cp = std::cos(*pitch);
    *roll = std::atan2(m21 / cp, m22 / cp);
    *yaw = std::atan2(m10 / cp, m00 / cp);
  }
}

Odometry LaserOdometer::makeOdometry(const Header& header, const Isometry3d& pose, const Isometry3d& prev_odom,
                                     const Isometry3d& laser_to_base, double prev_stamp, const std::string& fixed_frame,
                                     const std::string& base_frame) {
  Odometry msg;
  msg.header.frame_id = fixed_frame; msg.header.stamp = header.stamp; msg.header.seq = header.seq;
  msg.child_frame_id = base_frame;
  const Isometry3d odom_base_link = pose * laser_to_base;   // transform to base_link before publication
  msg.orientation = quat_from_rotation(odom_base_link);
  for (int k = 0; k < 3; ++k) msg.position[k] = odom_base_link.matrix()(k, 3);
  const double delta_time = header.stamp.toSec() - prev_stamp;
  const Isometry3d delta_odom = (prev_odom * laser_to_base).inverse() * odom_base_link;
  for (int k = 0; k < 3; ++k) msg.twist_linear[k] = delta_odom.matrix()(k, 3) / delta_time;
  double roll, pitch, yaw;
  rpy_from_quat(quat_from_rotation(delta_odom), &roll, &pitch, &yaw);
  msg.twist_angular[0] = roll / delta_time; msg.twist_angular[1] = pitch / delta_time; msg.twist_angular[2] = yaw / delta_time;
  return msg;
}

void LaserOdometer::publishOdom(const Header& header, const Isometry3d& pose) {
  const Odometry m = makeOdometry(header, pose, prev_odom_, laser_to_base_, prev_stamp_, params->fixed_frame_, params->base_frame_);
#ifdef LIODOM_FACADE_USE_PCL
  {   // src/laser_odometry.cc:395-446: nav_msgs/Odometry, geometry_msgs/TwistStamped and the TF, from inside the worker
    nav_msgs::Odometry laser_odom_msg;
    laser_odom_msg.header.frame_id = params->fixed_frame_;
    laser_odom_msg.child_frame_id = params->base_frame_;
    laser_odom_msg.header.stamp = header.stamp;
    laser_odom_msg.pose.pose.orientation.x = m.orientation.x(); laser_odom_msg.pose.pose.orientation.y = m.orientation.y();
    laser_odom_msg.pose.pose.orientation.z = m.orientation.z(); laser_odom_msg.pose.pose.orientation.w = m.orientation.w();
    laser_odom_msg.pose.pose.position.x = m.position[0]; laser_odom_msg.pose.pose.position.y = m.position[1]; laser_odom_msg.pose.pose.position.z = m.position[2];
    laser_odom_msg.twist.twist.linear.x = m.twist_linear[0]; laser_odom_msg.twist.twist.linear.y = m.twist_linear[1]; laser_odom_msg.twist.twist.linear.z = m.twist_linear[2];
    laser_odom_msg.twist.twist.angular.x = m.twist_angular[0]; laser_odom_msg.twist.twist.angular.y = m.twist_angular[1]; laser_odom_msg.twist.twist.angular.z = m.twist_angular[2];
    odom_pub_.publish(laser_odom_msg);
    geometry_msgs::TwistStamped twist_msg;
    twist_msg.header.frame_id = params->base_frame_;
    twist_msg.header.stamp = header.stamp;
    twist_msg.twist = laser_odom_msg.twist.twist;
    twist_pub_.publish(twist_msg);
    if (params->publish_tf_) {
      tf::Transform transform;
      transform.setOrigin(tf::Vector3(m.position[0], m.position[1], m.position[2]));
      transform.setRotation(tf::Quaternion(m.orientation.x(), m.orientation.y(), m.orientation.z(), m.orientation.w()));
      tf_broadcaster_.sendTransform(tf::StampedTransform(transform, header.stamp, params->fixed_frame_, params->base_frame_));
    }
  }
#endif
  if (odom_msg_cb_) odom_msg_cb_(m);
  prev_stamp_ = header.stamp.toSec();   // src/laser_odometry.cc:127,264
}

#ifdef LIODOM_FACADE_USE_PCL
// src/laser_odometry.cc:368-393: the static base -> laser transform, looked up once through tf
bool LaserOdometer::getBaseToLaserTf(const std::string& frame_id) {
  tf::TransformListener tf_listener;
  tf::StampedTransform laser_to_base_tf;
  try {
    tf_listener.waitForTransform(frame_id, params->base_frame_, ros::Time(0), ros::Duration(1.0));
    tf_listener.lookupTransform(frame_id, params->base_frame_, ros::Time(0), laser_to_base_tf);
  } catch (tf::TransformException& ex) {
    LIODOM_ERROR("Could not get initial transform from base to laser frame, %s", ex.what());
    return false;
  }
  const double qx = laser_to_base_tf.getRotation().x(), qy = laser_to_base_tf.getRotation().y(), qz = laser_to_base_tf.getRotation().z(), qw = laser_to_base_tf.getRotation().w();
  double m[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1};   // Eigen::Quaterniond::toRotationMatrix
  const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  m[0] = 1.0 - (tyy + tzz); m[1] = txy - twz; m[2] = txz + twy;
  m[4] = txy + twz; m[5] = 1.0 - (txx + tzz); m[6] = tyz - twx;
  m[8] = txz - twy; m[9] = tyz + twx; m[10] = 1.0 - (txx + tyy);
  m[3] = laser_to_base_tf.getOrigin().x(); m[7] = laser_to_base_tf.getOrigin().y(); m[11] = laser_to_base_tf.getOrigin().z();
  setLaserToBase(detail::pose_from16(m));
  return true;
}
#endif

// Output-rate watchdog of the reference (src/laser_odometry.cc:239-256): moving means over 5 samples of the input
// frequency (from the scan stamps) and of the output frequency (wall clock); a warning when out < 0.8 * in.
void LaserOdometer::watchdog(double stamp_secs) {
  const double now_secs = wall_now_secs();
  mean_in_freq_ -= in_freqs_[num_freqs_];
  in_freqs_[num_freqs_] = (1.0 / (stamp_secs - last_in_time_secs_)) / 5.0;
  mean_in_freq_ += in_freqs_[num_freqs_];
  last_in_time_secs_ = stamp_secs;
  mean_out_freq_ -= out_freqs_[num_freqs_];
  out_freqs_[num_freqs_] = (1.0 / (now_secs - last_out_time_secs_)) / 5.0;
  mean_out_freq_ += out_freqs_[num_freqs_];
  last_out_time_secs_ = now_secs;
  num_freqs_ = (num_freqs_ + 1) % 5;
  if (mean_out_freq_ < mean_in_freq_ * 0.8) {
    ++rate_warnings_;
    LIODOM_INFO("Output frequency too low: %2.2f (in: %2.2f)", mean_out_freq_, mean_in_freq_);
  }
}

void LaserOdometer::operator()(std::atomic<bool>& running) {
  while (running) {
    PointCloud::Ptr feats(new PointCloud);
    Header feat_header;
    if (sdata->popFeatures(feats, feat_header)) {
      const bool first = !init_;
#ifdef LIODOM_FACADE_USE_PCL
      if (first) {   // cache the static tf from base to laser (src/laser_odometry.cc:110-119)
        if (params->laser_frame_ == "") params->laser_frame_ = feat_header.frame_id;
        if (!getBaseToLaserTf(params->laser_frame_)) { LIODOM_ERROR("Skipping point_cloud"); return; }
      }
#endif
      const auto start_t = Clock::now();
      Isometry3d pose;
      const bool ok = process(feats, feat_header, &pose);
      const auto end_t = Clock::now();
      if (ok) {
        // both branches of the reference record pose, time and frame span (src/laser_odometry.cc:130-134, :259-263)
        stats->addPose(odom_.matrix());
        stats->addLaserOdometryTime(start_t, end_t);
        stats->stopFrame(end_t);
        // first frame: prev_stamp_ is set BEFORE publishOdom (:127), so its twist divides by zero as the reference's does
        if (first) prev_stamp_ = feat_header.stamp.toSec();
        else watchdog(feat_header.stamp.toSec());
        if (odom_cb_) odom_cb_(feat_header, odom_);
        publishOdom(feat_header, odom_);
      }
    }
    sdata->waitFeatures(2);   // the reference sleeps 2 ms here (src/laser_odometry.cc:270); see shared_data.h
  }
}

// ---------------------------------------------------------------------------------------------
// Map (src/map.cc:70-211)
// ---------------------------------------------------------------------------------------------
Map::Map(const double xy_size, const double z_size, const double res) : map_(nullptr) {
  int device = 0;
  if (const char* d = std::getenv("LIODOM_DEVICE")) device = std::atoi(d);
  int max_points = 1 << 22;
  if (const char* d = std::getenv("LIODOM_MAP_MAX_POINTS")) max_points = std::atoi(d);
  const int rc = liodom_map_create(xy_size, z_size, res, device, max_points, &map_);
  if (rc != LIODOM_OK) { LIODOM_ERROR("liodom_map_create failed (%d): %s", rc, liodom_map_last_error(nullptr)); map_ = nullptr; }
}
Map::~Map() { liodom_map_destroy(map_); }

void Map::updateMap(const PointCloud::Ptr& pc_in, const Isometry3d& pose) {
  if (!map_) return;
  std::vector<float> buf;
  xyzi_from_cloud(*pc_in, &buf);
  double T16[16];
  detail::pose_to16(pose, T16);
  const int rc = liodom_map_update(map_, buf.data(), (int)pc_in->size(), T16);
  if (rc != LIODOM_OK) LIODOM_ERROR("liodom_map_update failed (%d): %s", rc, liodom_map_last_error(map_));
}
PointCloud::Ptr Map::getMap() {
  PointCloud::Ptr out(new PointCloud);
  if (!map_) return out;
  int n = 0;
  liodom_map_size(map_, &n, nullptr);
  std::vector<float> buf((size_t)std::max(n, 1) * 4);
  const int rc = liodom_map_get(map_, buf.data(), n, &n);
  if (rc != LIODOM_OK) { LIODOM_ERROR("liodom_map_get failed (%d): %s", rc, liodom_map_last_error(map_)); return out; }
  cloud_from_xyzi(buf.data(), n, out.get());
  return out;
}
PointCloud::Ptr Map::getLocalMap(const Isometry3d& pose, int cells_xy, int cells_z) {
  PointCloud::Ptr out(new PointCloud);
  if (!map_) return out;
  int n = 0;
  double T16[16];
  detail::pose_to16(pose, T16);
  int rc = liodom_map_get_local(map_, T16, cells_xy, cells_z, nullptr, 0, &n);
  std::vector<float> buf((size_t)std::max(n, 1) * 4);
  if (rc == LIODOM_OK && n > 0) rc = liodom_map_get_local(map_, T16, cells_xy, cells_z, buf.data(), n, &n);
  if (rc != LIODOM_OK) { LIODOM_ERROR("liodom_map_get_local failed (%d): %s", rc, liodom_map_last_error(map_)); return out; }
  cloud_from_xyzi(buf.data(), n, out.get());
  return out;
}
double Map::getMapEntropy() {
  if (!map_) return 0.0;
  int n = 0, nc = 0;
  liodom_map_size(map_, &n, &nc);
  if (nc == 0 || n == 0) return 0.0;
  std::vector<int32_t> keys((size_t)nc * 3), counts((size_t)nc);
  if (liodom_map_cells(map_, keys.data(), counts.data(), nc, &nc) != LIODOM_OK) return 0.0;
  double h = 0.0;
  for (int i = 0; i < nc; ++i)
    if (counts[(size_t)i] > 0) { const double p = counts[(size_t)i] / (double)n; h += p * std::log(p); }
  return -h;
}

}  // namespace liodom

#ifndef LIODOM_FACADE_USE_PCL   // with ROS present the reference's own liodom_node.cc is the harness
// ---------------------------------------------------------------------------------------------
// Array-driven harness standing in for liodom_node / liodom_mapping_node (src/liodom_node.cc:72-121,
// src/liodom_mapping_node.cc:45-90): it feeds clouds through SharedData exactly like lidarClb does
// and runs the two worker functors on their own threads.  C linkage so that tests can call it.
// ---------------------------------------------------------------------------------------------
#include <liodom/harness.h>   // liodom_host_options and the harness entry points defined below

static std::vector<double> g_run_push_ms, g_run_pose_ms;   // liodom_host_last_run_times()

static int run_sequence_impl(const liodom_host_options* opt, int nframes, const std::function<liodom::PointCloud::Ptr(int)>& make_cloud,
                             double* poses_out, int* nfeats_out, const char* results_dir) {
  using namespace liodom;
  NodeHandle nh;
  nh.setParam("min_range", opt->min_range); nh.setParam("max_range", opt->max_range);
  nh.setParam("lidar_type", opt->lidar_type); nh.setParam("scan_lines", opt->scan_lines);
  nh.setParam("scan_regions", opt->scan_regions); nh.setParam("edges_per_region", opt->edges_per_region);
  nh.setParam("prev_frames", opt->prev_frames); nh.setParam("mapping", opt->mapping != 0);
  nh.setParam("filter_local_map", opt->filter_local_map != 0);
  Params::getInstance()->readParams(nh);
  Stats* stats = Stats::getInstance();
  stats->clear();
  SharedData* sdata = SharedData::getInstance();
  FeatureExtractor fext(nh);
  LaserOdometer lodom(nh);
  std::atomic<int> produced(0);
  std::vector<int> nf((size_t)nframes, 0);
  // Wall-clock marks of the run for liodom_host_last_run_times(): the reference's Stats keeps whole milliseconds
  // (src/stats.cc:42-67), too coarse for sub-millisecond frames.
  const auto run_t0 = Clock::now();
  auto since_ms = [&]() { return std::chrono::duration<double, std::milli>(Clock::now() - run_t0).count(); };
  g_run_push_ms.assign((size_t)nframes, 0.0);
  g_run_pose_ms.assign((size_t)nframes, 0.0);
  fext.setEdgesCallback([&](const Header& h, const PointCloud::Ptr& e) { if ((int)h.seq < nframes) nf[h.seq] = (int)e->size(); });
  lodom.setOdomCallback([&](const Header& h, const Isometry3d& pose) {
    if ((int)h.seq < nframes && poses_out) detail::pose_to16(pose, poses_out + 16 * (size_t)h.seq);
    if ((int)h.seq < nframes) g_run_pose_ms[h.seq] = since_ms();
    produced++;
  });
  std::atomic<bool> running(true);
  std::thread fext_thread(fext, std::ref(running));
  std::thread lodom_thread(lodom, std::ref(running));
  for (int f = 0; f < nframes; ++f) {   // lidarClb (src/liodom_node.cc:40-55)
    PointCloud::Ptr pc = make_cloud(f);
    Header h; h.seq = (uint32_t)f; detail::set_stamp(h, 0.1 * f); h.frame_id = "laser";
    stats->startFrame(Clock::now());
    g_run_push_ms[(size_t)f] = since_ms();
    sdata->pushPointCloud(pc, h);
    if (opt->lockstep) {
      const auto t0 = Clock::now();
      while (produced.load() <= f && std::chrono::duration_cast<std::chrono::seconds>(Clock::now() - t0).count() < 30)
        std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
  }
  const auto t0 = Clock::now();
  while (produced.load() < nframes && std::chrono::duration_cast<std::chrono::seconds>(Clock::now() - t0).count() < 60)
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  running = false;
  fext_thread.join();
  lodom_thread.join();
  if (nfeats_out) std::memcpy(nfeats_out, nf.data(), sizeof(int) * (size_t)nframes);
  if (results_dir && results_dir[0]) stats->writeResults(results_dir);
  return produced.load();
}

extern "C" {

/* Per-frame wall-clock marks of the last liodom_host_run_sequence[_msgs] call, in milliseconds since its start: when
 * the cloud was pushed and when its pose came out.  Returns the number of frames written (<= cap).  Harness only. */
int liodom_host_last_run_times(double* push_ms, double* pose_ms, int cap) {
  const int n = (int)std::min(g_run_push_ms.size(), (size_t)std::max(cap, 0));
  for (int i = 0; i < n; ++i) { if (push_ms) push_ms[i] = g_run_push_ms[(size_t)i]; if (pose_ms) pose_ms[i] = g_run_pose_ms[(size_t)i]; }
  return n;
}

int liodom_host_run_sequence(const liodom_host_options* opt, const float* scans, const int* npts, int nframes,
                             double* poses_out, int* nfeats_out, const char* results_dir) {
  using namespace liodom;
  std::vector<size_t> start((size_t)nframes + 1, 0);
  for (int f = 0; f < nframes; ++f) start[f + 1] = start[f] + (size_t)npts[f];
  return run_sequence_impl(opt, nframes, [&](int f) {
    PointCloud::Ptr pc(new PointCloud);
    pc->points.resize((size_t)npts[f]);
    for (int i = 0; i < npts[f]; ++i) {
      Point& q = pc->points[(size_t)i];
      const float* s = scans + (start[f] + (size_t)i) * 4;
      q.x = s[0]; q.y = s[1]; q.z = s[2]; q.intensity = s[3];
    }
    pc->width = opt->lidar_type == 1 ? (uint32_t)opt->width : (uint32_t)npts[f];
    pc->height = opt->lidar_type == 1 ? (uint32_t)opt->height : 1;
    return pc;
  }, poses_out, nfeats_out, results_dir);
}

// publishOdom's arithmetic through the façade (test hook): out13 = orientation x,y,z,w, position,
// twist linear, twist angular.
void liodom_host_make_odometry(const double* pose16, const double* prev_odom16, const double* l2b16, double stamp, double prev_stamp, double* out13) {
  using namespace liodom;
  const Isometry3d pose = detail::pose_from16(pose16), prev = detail::pose_from16(prev_odom16), l2b = detail::pose_from16(l2b16);
  Header h; detail::set_stamp(h, stamp);
  const Odometry m = LaserOdometer::makeOdometry(h, pose, prev, l2b, prev_stamp, "odom", "base_link");
  out13[0] = m.orientation.x(); out13[1] = m.orientation.y(); out13[2] = m.orientation.z(); out13[3] = m.orientation.w();
  for (int k = 0; k < 3; ++k) { out13[4 + k] = m.position[k]; out13[7 + k] = m.twist_linear[k]; out13[10 + k] = m.twist_angular[k]; }
}

// Stats::writeResults through the façade (test hook; compared with the reference's own Stats, src/stats.cc:73-132):
// poses16 [n x 16], nfeats [n] (int64), times_ms [n x 2] = (feature extraction, laser odometry) spans.
void liodom_host_stats_write(const double* poses16, const long long* nfeats, const double* times_ms, int n, const char* dir) {
  using namespace liodom;
  Stats* stats = Stats::getInstance();
  stats->clear();
  const Clock::time_point t0 = Clock::now();
  for (int k = 0; k < n; ++k) {
    Matrix4d M;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) M(i, j) = poses16[(size_t)k * 16 + i * 4 + j];
    stats->addPose(M);
    stats->addNumOfFeats((size_t)nfeats[k]);
    const auto fe = std::chrono::microseconds((long long)(times_ms[2 * k] * 1000.0)), lo = std::chrono::microseconds((long long)(times_ms[2 * k + 1] * 1000.0));
    stats->addFeatureExtractionTime(t0, t0 + fe);
    stats->addLaserOdometryTime(t0, t0 + lo);
    stats->startFrame(t0);
    stats->stopFrame(t0 + fe + lo);
  }
  stats->writeResults(dir);
  stats->clear();
}

// The same with sensor_msgs/PointCloud2 messages, as lidarClb receives them: `data` holds the frames'
// blobs back to back (frame f: height[f] * row_step bytes), `field_names` is a comma-separated list
// matching field_offsets / field_datatypes.  The bytes are never decoded on the host.
int liodom_host_run_sequence_msgs(const liodom_host_options* opt, const unsigned char* data, const int* widths, const int* heights,
                                  int nframes, int point_step, int row_pad, const char* field_names, const int* field_offsets,
                                  const int* field_datatypes, int nfields, double* poses_out, int* nfeats_out, const char* results_dir) {
  using namespace liodom;
  std::vector<PointField> fields;
  std::string names(field_names ? field_names : "");
  size_t p0 = 0;
  for (int k = 0; k < nfields; ++k) {
    const size_t p1 = names.find(',', p0);
    PointField f;
    f.name = names.substr(p0, p1 == std::string::npos ? std::string::npos : p1 - p0);
    f.offset = (uint32_t)field_offsets[k]; f.datatype = (uint8_t)field_datatypes[k]; f.count = 1;
    fields.push_back(f);
    p0 = p1 == std::string::npos ? names.size() : p1 + 1;
  }
  std::vector<size_t> start((size_t)nframes + 1, 0);
  for (int f = 0; f < nframes; ++f) start[f + 1] = start[f] + (size_t)heights[f] * ((size_t)widths[f] * point_step + row_pad);
  return run_sequence_impl(opt, nframes, [&](int f) {
    PointCloud2::Ptr msg(new PointCloud2);
    msg->width = (uint32_t)widths[f]; msg->height = (uint32_t)heights[f]; msg->fields = fields;
    msg->point_step = (uint32_t)point_step; msg->row_step = (uint32_t)(widths[f] * point_step + row_pad);
    msg->data.assign(data + start[f], data + start[f + 1]);
    PointCloud::Ptr pc(new PointCloud);
    fromROSMsgDeferred(msg, *pc);
    return pc;
  }, poses_out, nfeats_out, results_dir);
}

}  // extern "C"
#endif  // !LIODOM_FACADE_USE_PCL
