/* liodom::FeatureExtractor — worker functor of thread A (include/liodom/feature_extractor.h:62-85,
 * src/feature_extractor.cc:24-82).  Same constructor / operator() as the reference; the private
 * helpers isValidPoint / splitPointCloud / extractFeatures / extractFeaturesFromRegion are ONE
 * call into the CUDA library (liodom_extract), there is no CPU implementation behind it. */
#ifndef INCLUDE_LIODOM_FEATURE_EXTRACTOR_H
#define INCLUDE_LIODOM_FEATURE_EXTRACTOR_H

#include <atomic>
#include <functional>

#include <liodom/params.h>
#include <liodom/shared_data.h>
#include <liodom/stats.h>

struct liodom_ctx;

namespace liodom {

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
