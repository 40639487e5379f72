/* liodom::LocalMapManager and liodom::LaserOdometer — worker functor of thread B
 * (include/liodom/laser_odometry.h:59-123, src/laser_odometry.cc:24-272).  The sliding window,
 * the voxel-hash association and the Levenberg-Marquardt solve all live on the GPU behind
 * liodom_lmap_* / liodom_register; Ceres / PCL / FLANN are not used. */
#ifndef INCLUDE_LIODOM_LASER_ODOMETRY_H
#define INCLUDE_LIODOM_LASER_ODOMETRY_H

#include <atomic>
#include <functional>
#include <thread>

#ifdef LIODOM_FACADE_USE_PCL
// ROS / tf / message headers of the reference's laser_odometry.h (include/liodom/laser_odometry.h:33-50)
#include <ros/ros.h>
#include <nav_msgs/Odometry.h>
#include <geometry_msgs/TwistStamped.h>
#include <tf/transform_datatypes.h>
#include <tf/transform_listener.h>
#include <tf/transform_broadcaster.h>
#include <pcl_conversions/pcl_conversions.h>
#endif

#include <liodom/params.h>
#include <liodom/shared_data.h>
#include <liodom/stats.h>

struct liodom_ctx;

namespace liodom {

class LocalMapManager {
 public:
  explicit LocalMapManager(const size_t max_frames);
  virtual ~LocalMapManager();

  void addPointCloud(const PointCloud::Ptr& pc);
  size_t getLocalMap(PointCloud::Ptr& map);
  void setMaxFrames(const size_t max_nframes);

 private:
  std::shared_ptr<liodom_ctx> ctx_;
  size_t max_nframes_;
  size_t frame_cap_ = 0;
  bool ensureContext(size_t frame_points);
};

class LaserOdometer {
 public:
  explicit LaserOdometer(const NodeHandle& nh);
  LaserOdometer(const LaserOdometer& o);
  virtual ~LaserOdometer();

  void operator()(std::atomic<bool>& running);

  /* One popFeatures() iteration (src/laser_odometry.cc:108-266) without the queues. */
  bool process(const PointCloud::Ptr& feats, const Header& header, Isometry3d* pose_out);
  /* stands in for publishOdom (src/laser_odometry.cc:395-446): header + odom_ (laser frame) */
  void setOdomCallback(std::function<void(const Header&, const Isometry3d&)> cb) { odom_cb_ = cb; }
  /* the messages publishOdom assembles (odometry in base_link + twist); called once per frame */
  void setOdometryMsgCallback(std::function<void(const Odometry&)> cb) { odom_msg_cb_ = cb; }
  /* stands in for getBaseToLaserTf (src/laser_odometry.cc:368-393): the static laser->base transform */
  void setLaserToBase(const Isometry3d& laser_to_base);
  /* publishOdom's arithmetic (src/laser_odometry.cc:395-432) */
  static Odometry makeOdometry(const Header& header, const Isometry3d& pose, const Isometry3d& prev_odom,
                               const Isometry3d& laser_to_base, double prev_stamp, const std::string& fixed_frame,
                               const std::string& base_frame);

 private:
  NodeHandle nh_;
#ifdef LIODOM_FACADE_USE_PCL
  ros::Publisher odom_pub_, twist_pub_;            // published from inside the worker (src/laser_odometry.cc:93-94, :395-446)
  tf::TransformBroadcaster tf_broadcaster_;
  bool getBaseToLaserTf(const std::string& frame_id);   // src/laser_odometry.cc:368-393
#endif
  bool init_;
  Isometry3d odom_;
  SharedData* sdata;
  Stats* stats;
  Params* params;
  std::shared_ptr<liodom_ctx> ctx_;
  std::function<void(const Header&, const Isometry3d&)> odom_cb_;
  std::function<void(const Odometry&)> odom_msg_cb_;
  Isometry3d prev_odom_, laser_to_base_;
  double prev_stamp_ = 0.0;
  // output-rate watchdog (src/laser_odometry.cc:239-256): 5-sample moving means of the input / output frequency
  double in_freqs_[5], out_freqs_[5], mean_in_freq_ = 100.0, mean_out_freq_ = 100.0;
  int num_freqs_ = 0;
  double last_in_time_secs_ = 0.0, last_out_time_secs_ = 0.0;
  long rate_warnings_ = 0;
  void watchdog(double stamp_secs);
 public:
  long rateWarnings() const { return rate_warnings_; }   // how often "Output frequency too low" fired (the reference only logs it)
 private:
  bool ensureContext();
  void publishOdom(const Header& header, const Isometry3d& pose);
};

}  // namespace liodom
#endif  // INCLUDE_LIODOM_LASER_ODOMETRY_H
