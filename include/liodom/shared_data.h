/* liodom::SharedData — the reference's mutex-guarded hand-off singleton
 * (include/liodom/shared_data.h:36-85, src/shared_data.cc:37-117). Host-only: clouds are
 * shared by pointer, the latest local map is deep-copied under its mutex. */
#ifndef INCLUDE_LIODOM_SHARED_DATA_H
#define INCLUDE_LIODOM_SHARED_DATA_H

#include <condition_variable>
#include <mutex>
#include <queue>

#include <liodom/defs.h>

namespace liodom {

class SharedData {
 public:
  static SharedData* getInstance();

  // ROS callback thread -> extractor thread: scans by pointer, FIFO (src/shared_data.cc:37-62)
  void pushPointCloud(const PointCloud::Ptr& pc_in, const Header& header);
  bool popPointCloud(PointCloud::Ptr& pc_out, Header& header);
  // extractor thread -> odometry thread: edge clouds by pointer, FIFO (:64-89)
  void pushFeatures(const PointCloud::Ptr& feat_in, Header& header);
  bool popFeatures(PointCloud::Ptr& feat_out, Header& header);
  // Not in the reference: its workers sleep 2 ms after every loop turn (src/feature_extractor.cc:80,
  // src/laser_odometry.cc:270), which caps a stream at < 500 scans/s and adds up to 4 ms of latency per scan.
  // The workers here wait on the queue instead — at most `ms` milliseconds, returning at once when an item is
  // there — so the same loop runs at the speed of the GPU stages.  Results are unaffected.
  void waitPointCloud(int ms);
  void waitFeatures(int ms);
  // mapping process -> odometry: latest local map, deep-copied both ways (:91-105)
  void setLocalMap(const PointCloud::Ptr& map_in);
  void getLocalMap(PointCloud::Ptr& map_out);
  // IMU callback -> odometry: latest orientation (:107-117)
  void setLastIMUOri(Quaterniond& imu_ori);
  void getLastIMUOri(Quaterniond& imu_ori);

  SharedData(SharedData const&) = delete;
  void operator=(SharedData const&) = delete;

 protected:
  SharedData() : local_map_(new PointCloud) {}
  ~SharedData() {}

 private:
  template <typename T> struct Fifo { std::mutex m; std::condition_variable cv; std::queue<T> items; std::queue<Header> headers; };
  static SharedData* instance_;
  static std::mutex instance_mutex_;
  Fifo<PointCloud::Ptr> scans_, feats_;
  std::mutex map_mutex_, imu_mutex_;
  PointCloud::Ptr local_map_;
  Quaterniond last_imu_;
};

}  // namespace liodom
#endif  // INCLUDE_LIODOM_SHARED_DATA_H
