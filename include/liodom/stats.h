/* liodom::Stats — pose / timing collection and the five result files of the reference
 * (include/liodom/stats.h:48-54, src/stats.cc:36-132): poses.txt (KITTI 3x4 rows, default
 * ostream precision), feat_ext_times.txt, laser_odom_times.txt, nfeats.txt, frame_times.txt,
 * times truncated to integer milliseconds. Host-only. */
#ifndef INCLUDE_LIODOM_STATS_H
#define INCLUDE_LIODOM_STATS_H

#include <mutex>
#include <queue>
#include <string>
#include <vector>

#include <liodom/defs.h>

namespace liodom {

class Stats {
 public:
  static Stats* getInstance();
  Stats(Stats const&) = delete;
  void operator=(Stats const&) = delete;

  void addPose(const Matrix4d& pose);
  void addFeatureExtractionTime(const Clock::time_point& start, const Clock::time_point& end);
  void addLaserOdometryTime(const Clock::time_point& start, const Clock::time_point& end);
  void addNumOfFeats(const size_t& nfeats);
  void startFrame(const Clock::time_point& start);
  void stopFrame(const Clock::time_point& stop);
  void writeResults(const std::string& dir);

  /* additions for the array-driven harness (not in the reference) */
  void clear();
  const std::vector<Matrix4d>& poses() const { return poses_; }

 private:
  static Stats* pinstance_;
  static std::mutex sdata_mutex_;
  std::vector<Matrix4d> poses_;
  std::vector<double> feat_extr_;
  std::vector<double> laser_odom_;
  std::vector<size_t> num_of_features_;
  std::mutex frame_mutex_;
  std::queue<Clock::time_point> start_times_;
  std::vector<double> frame_times_;

 protected:
  Stats() {}
  ~Stats() {}
};

}  // namespace liodom
#endif  // INCLUDE_LIODOM_STATS_H
