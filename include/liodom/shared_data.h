/* liodom::SharedData — the reference's mutex-guarded hand-off singleton
 * (include/liodom/shared_data.h:36-85, src/shared_data.cc:37-117). Host-only: clouds are
 * shared by pointer, the latest local map is deep-copied under its mutex. */
#ifndef INCLUDE_LIODOM_SHARED_DATA_H
#define INCLUDE_LIODOM_SHARED_DATA_H

#include <mutex>
#include <queue>

#include <liodom/defs.h>

namespace liodom {

class SharedData {
 public:
  static SharedData* getInstance();
  SharedData(SharedData const&) = delete;
  void operator=(SharedData const&) = delete;

  void pushPointCloud(const PointCloud::Ptr& pc_in, const Header& header);
  bool popPointCloud(PointCloud::Ptr& pc_out, Header& header);

  void pushFeatures(const PointCloud::Ptr& feat_in, Header& header);
  bool popFeatures(PointCloud::Ptr& feat_out, Header& header);

  void setLocalMap(const PointCloud::Ptr& map_in);
  void getLocalMap(PointCloud::Ptr& map_out);

  void setLastIMUOri(Quaterniond& imu_ori);
  void getLastIMUOri(Quaterniond& imu_ori);

 private:
  static SharedData* pinstance_;
  static std::mutex sdata_mutex_;
  std::mutex pc_mutex_;
  std::queue<PointCloud::Ptr> pc_buf_;
  std::queue<Header> pc_header_;
  std::mutex feat_mutex_;
  std::queue<PointCloud::Ptr> feat_buf_;
  std::queue<Header> feat_header_;
  std::mutex map_mutex_;
  PointCloud::Ptr local_map_;
  std::mutex imu_mutex_;
  Quaterniond last_IMU_ori_;

 protected:
  SharedData() : local_map_(new PointCloud) {}
  ~SharedData() {}
};

}  // namespace liodom
#endif  // INCLUDE_LIODOM_SHARED_DATA_H
