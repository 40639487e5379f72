/* Array-driven harness of the C++ façade (libliodom_host.so): what `main()` + `lidarClb` of the reference's
 * src/liodom_node.cc:40-55,85-91 do for a ROS bag, for arrays of scans.  It builds the reference's objects
 * (Params, SharedData, Stats, FeatureExtractor and LaserOdometer on their own threads), pushes the clouds through the
 * SharedData queues and collects the poses.  Test / bench infrastructure around the façade: the product interface is
 * include/liodom_b200.h (C ABI) and the classes of include/liodom/ (drop-in).  Plain C. */
#ifndef INCLUDE_LIODOM_HARNESS_H
#define INCLUDE_LIODOM_HARNESS_H

#ifdef __cplusplus
extern "C" {
#endif

/* Parameters under their ROS names (src/params.cc:35-63). */
typedef struct liodom_host_options {
  double min_range, max_range;
  int lidar_type, scan_lines, scan_regions, edges_per_region, prev_frames, mapping;
  int width, height;      /* organised clouds (lidar_type 1) */
  int lockstep;           /* 1: wait for each frame's pose before pushing the next cloud (deterministic) */
  int filter_local_map;
} liodom_host_options;

/* scans: the frames' float32 x, y, z, intensity records back to back; npts[nframes].  poses_out: nframes x 16
 * (row-major 4x4), nfeats_out: edges per frame; either may be NULL.  results_dir ("" or NULL: none): Stats::writeResults
 * target.  Returns the number of poses produced, negative on a set-up failure. */
int liodom_host_run_sequence(const liodom_host_options* opt, const float* scans, const int* npts, int nframes,
                             double* poses_out, int* nfeats_out, const char* results_dir);

/* The same with sensor_msgs/PointCloud2 messages as lidarClb receives them: `data` holds the frames' blobs back to back
 * (frame f: heights[f] rows of widths[f] * point_step + row_pad bytes), `field_names` is a comma-separated list matching
 * field_offsets / field_datatypes (sensor_msgs/PointField datatype codes).  The bytes are never decoded on the host. */
int liodom_host_run_sequence_msgs(const liodom_host_options* opt, const unsigned char* data, const int* widths, const int* heights,
                                  int nframes, int point_step, int row_pad, const char* field_names, const int* field_offsets,
                                  const int* field_datatypes, int nfields, double* poses_out, int* nfeats_out, const char* results_dir);

/* Wall-clock marks of the last run, milliseconds since its start: when frame f's cloud was pushed and when its pose came
 * out (the reference's Stats keeps whole milliseconds, src/stats.cc:42-67).  Returns the frames written (<= cap). */
int liodom_host_last_run_times(double* push_ms, double* pose_ms, int cap);

/* Test hooks: publishOdom's arithmetic (src/laser_odometry.cc:395-446; out13 = orientation x, y, z, w, position, twist
 * linear, twist angular) and Stats::writeResults (src/stats.cc:73-132; times_ms [n x 2] = extraction, odometry). */
void liodom_host_make_odometry(const double* pose16, const double* prev_odom16, const double* l2b16, double stamp, double prev_stamp,
                               double* out13);
void liodom_host_stats_write(const double* poses16, const long long* nfeats, const double* times_ms, int n, const char* dir);

#ifdef __cplusplus
}
#endif
#endif  /* INCLUDE_LIODOM_HARNESS_H */
