/* liodom::FeatureExtractor — worker functor of thread A (include/liodom/feature_extractor.h:62-85,
 * src/feature_extractor.cc:24-82).  Same constructor / operator() as the reference; the private
 * helpers isValidPoint / splitPointCloud / extractFeatures / extractFeaturesFromRegion are ONE
 * call into the CUDA library (liodom_extract), there is no CPU implementation behind it. */
#ifndef INCLUDE_LIODOM_FEATURE_EXTRACTOR_H
#define INCLUDE_LIODOM_FEATURE_EXTRACTOR_H

#include <atomic>
#include <functional>
#include <thread>

#ifdef LIODOM_FACADE_USE_PCL
// the third-party headers the reference's feature_extractor.h pulls in (include/liodom/feature_extractor.h:24-34);
// src/liodom_node.cc relies on them transitively (pcl::fromROSMsg, sensor_msgs::PointCloud2ConstPtr)
#include <omp.h>
#include <pcl_conversions/pcl_conversions.h>
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>
#endif

#include <liodom/params.h>
#include <liodom/shared_data.h>
#include <liodom/stats.h>

struct liodom_ctx;
struct liodom_cloud_layout;

namespace liodom {

/* pcl::fromROSMsg's field matching (pcl/conversions.h createMapping: same name, datatype FLOAT32,
 * count 1) for liodom::Point = PointXYZI.  False when x, y or z has no match or the message is
 * big-endian; a missing intensity gives off_intensity = -1 (the field stays 0, PCL only warns). */
bool cloudLayoutFromFields(const PointCloud2& msg, liodom_cloud_layout* layout);
#ifndef LIODOM_FACADE_USE_PCL   // with PCL present pcl::fromROSMsg is used as the reference does
/* Host-side pcl::fromROSMsg (used for the small received local map, mapClb src/liodom_node.cc:57-64). */
bool fromROSMsg(const PointCloud2& msg, PointCloud& cloud);
/* lidarClb's conversion (src/liodom_node.cc:40-44) without touching the points: keeps the message. */
bool fromROSMsgDeferred(const PointCloud2::ConstPtr& msg, PointCloud& cloud);
#endif

class FeatureExtractor {
 public:
  explicit FeatureExtractor(const NodeHandle& nh);
  FeatureExtractor(const FeatureExtractor& o);   // the reference copies the functor into std::thread
  virtual ~FeatureExtractor();

  void operator()(std::atomic<bool>& running);

  /* Body of one loop iteration (src/feature_extractor.cc:52-59): split + extract on the GPU.
   * Public so that it can be driven without the queues; returns false on a library error
   * (logged, processing continues — the reference's convention). */
  bool extract(const PointCloud::Ptr& pc_curr, PointCloud::Ptr& pc_edges);
  /* stands in for the ROS publisher of "edges" (src/feature_extractor.cc:71-74) */
  void setEdgesCallback(std::function<void(const Header&, const PointCloud::Ptr&)> cb) { edges_cb_ = cb; }

 private:
  NodeHandle nh_;
#ifdef LIODOM_FACADE_USE_PCL
  ros::Publisher pc_edges_pub_;   // "edges", published from inside the worker as the reference does (src/feature_extractor.cc:36, :71-74)
#endif
  SharedData* sdata;
  Stats* stats;
  Params* params;
  std::shared_ptr<liodom_ctx> ctx_;
  std::function<void(const Header&, const PointCloud::Ptr&)> edges_cb_;
  bool ensureContext(size_t npoints);
  size_t ctx_points_ = 0;
};

}  // namespace liodom
#endif  // INCLUDE_LIODOM_FEATURE_EXTRACTOR_H
