/* liodom::Params — same public fields, defaults and derived values as the reference
 * (include/liodom/params.h:33-56, src/params.cc:37-110). Host-only. */
#ifndef INCLUDE_LIODOM_PARAMS_H
#define INCLUDE_LIODOM_PARAMS_H

#include <mutex>
#include <string>

#include <liodom/defs.h>

namespace liodom {

class Params {
 public:
  double min_range_;
  double max_range_;
  int lidar_type_;
  int scan_lines_;
  int scan_regions_;
  int edges_per_region_;
  size_t min_points_per_scan_;
  size_t local_map_size_;
  bool save_results_;
  std::string results_dir_;
  std::string fixed_frame_;
  std::string base_frame_;
  std::string laser_frame_;
  bool use_imu_;
  bool filter_local_map_;
  bool mapping_;
  bool publish_tf_;

  static Params* getInstance();
  Params(Params const&) = delete;
  void operator=(Params const&) = delete;

  void readParams(const NodeHandle& nh);

 private:
  static Params* pinstance_;
  static std::mutex params_mutex_;

 protected:
  Params() { readParams(NodeHandle()); }
  ~Params() {}
};

}  // namespace liodom
#endif  // INCLUDE_LIODOM_PARAMS_H
