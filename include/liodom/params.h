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
  // --- process-wide instance (the reference's singleton access, include/liodom/params.h:54-56) ---
  static Params* getInstance();
  void readParams(const NodeHandle& nh);

  // --- sensor and extraction (src/params.cc:40-66) ---
  double min_range_, max_range_;          // XY range gate of isValidPoint [m]
  int lidar_type_;                        // 0 Velodyne (ring from elevation), 1 Ouster (ring = row)
  int scan_lines_, scan_regions_, edges_per_region_;
  size_t min_points_per_scan_;            // derived: scan_regions * edges_per_region + 10

  // --- registration (src/params.cc:68-72, :96-104) ---
  size_t local_map_size_;                 // "prev_frames"
  bool use_imu_, filter_local_map_, mapping_;

  // --- outputs (src/params.cc:74-94, :106-108) ---
  bool save_results_, publish_tf_;
  std::string results_dir_;
  std::string fixed_frame_, base_frame_, laser_frame_;

  Params(Params const&) = delete;
  void operator=(Params const&) = delete;

 protected:
  Params();   // the reference's defaults (src/params.cc:40-108); readParams overrides them
  ~Params() {}

 private:
  static Params* instance_;
  static std::mutex instance_mutex_;
};

}  // namespace liodom
#endif  // INCLUDE_LIODOM_PARAMS_H
