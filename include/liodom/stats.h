/* liodom::Stats — pose / timing collection and the five result files of the reference
 * (include/liodom/stats.h:48-54, src/stats.cc:36-132): poses.txt (KITTI 3x4 rows, default
 * ostream precision), feat_ext_times.txt, laser_odom_times.txt, nfeats.txt, frame_times.txt,
 * times truncated to integer milliseconds. Host-only. */
#ifndef INCLUDE_LIODOM_STATS_H
#define INCLUDE_LIODOM_STATS_H

// the same standard headers as the reference's stats.h (include/liodom/stats.h:24-30): src/liodom_node.cc uses std::cout through them
#include <fstream>
#include <iostream>
#include <mutex>
#include <queue>
#include <sstream>
#include <string>
#include <vector>

#include <liodom/defs.h>

namespace liodom {

class Stats {
 public:
  static Stats* getInstance();

  // per-frame records (src/stats.cc:36-71)
  void addPose(const Matrix4d& pose);
  void addNumOfFeats(const size_t& nfeats);
  void addFeatureExtractionTime(const Clock::time_point& start, const Clock::time_point& end);
  void addLaserOdometryTime(const Clock::time_point& start, const Clock::time_point& end);
  // whole-frame latency: startFrame in lidarClb, stopFrame when the pose is out (FIFO of start times)
  void startFrame(const Clock::time_point& start);
  void stopFrame(const Clock::time_point& stop);
  // the five text files (src/stats.cc:73-132)
  void writeResults(const std::string& dir);

  /* additions for the array-driven harness (not in the reference) */
  void clear();
  const std::vector<Matrix4d>& poses() const { return poses_; }

  Stats(Stats const&) = delete;
  void operator=(Stats const&) = delete;

 protected:
  Stats() {}
  ~Stats() {}

 private:
  static Stats* instance_;
  static std::mutex instance_mutex_;
  std::mutex frame_mutex_;                       // guards pending_starts_ / frame_ms_
  std::queue<Clock::time_point> pending_starts_;
  std::vector<Matrix4d> poses_;
  std::vector<size_t> nfeats_;
  std::vector<double> extract_ms_, odom_ms_, frame_ms_;   // whole milliseconds, as the reference truncates them
};

}  // namespace liodom
#endif  // INCLUDE_LIODOM_STATS_H
