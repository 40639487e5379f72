/* liodom oracle — TEST INFRASTRUCTURE ONLY.
 *
 * A dependency-free CPU restatement of the LiODOM per-scan hot path
 * (emiliofidalgo/liodom).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product (liodom_b200/) never links, imports or calls it.
 *
 * PARITY STATUS
 *   - PINNED by the reference's own object code: oracle/_ref/libliodom_ref.so is
 *     /root/reference/src/{feature_extractor,laser_odometry,map,params,shared_data,stats}.cc
 *     compiled UNMODIFIED (make -C oracle ref) against the test-only API shim oracle/refshim/.
 *     tests/test_ref_pins_oracle.py: ring split and edge lists bit-exact (A1-A4), LocalMapManager
 *     (A5), Point2LineFactor residual/Jacobian from the reference's own functor through Jets (A9),
 *     the whole LaserOdometer::operator() control flow (A6-A11: poses equal to the last bit on
 *     the synthetic sequences), Map keys / creation order / counts (A12-A13), Stats files.
 *   - STILL UNPINNED ("parity unpinned" for these only): what the reference delegates to PCL /
 *     FLANN / Eigen / Ceres / tf.  None is vendored in /root/reference or installed here; inferred
 *     versions (Ubuntu 20.04 / ROS Noetic): PCL 1.10.0, FLANN 1.9.1, Eigen 3.3.7, Ceres 1.14.0.
 *     Their published algorithms are restated twice (here and, independently, in refshim/):
 *     kNN tie order among exactly equal distances, VoxelGrid's in-voxel accumulation order
 *     (PCL sorts unstably; here: input order), the eigen gate within 1e-15 of equality and the
 *     Ceres trust-region loop rest on SURVEY.md App. A plus the NumPy/SciPy cross-checks in tests/.
 */
#ifndef LIODOM_ORACLE_H
#define LIODOM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors liodom::Params (include/liodom/params.h:33-52, defaults src/params.cc:40-108). */
typedef struct {
  double min_range;        /* 3.0  */
  double max_range;        /* 75.0 */
  int lidar_type;          /* 0 Velodyne, 1 Ouster */
  int scan_lines;          /* 64 */
  int scan_regions;        /* 8 */
  int edges_per_region;    /* 10 */
  int prev_frames;         /* 5 (launch: 15) -> local_map_size_ */
  int filter_local_map;    /* false */
  int mapping;             /* false */
  int omp_threads;         /* 0: reference rules (extractor max(2, omp_get_max_threads()-5), solver nproc);
                              >0: that many threads for both (several sequences sharing one host) */
} OrcParams;

void orc_default_params(OrcParams* p);

/* A1+A2  FeatureExtractor::isValidPoint / splitPointCloud (src/feature_extractor.cc:84-179).
 * pts: n points, stride_f floats apart (x,y,z,intensity first).  For lidar_type 1 the
 * cloud is organised width x height (n == width*height).
 * ring_of_point[n]: ring id or -1.  rings_xyzi: ring-major compacted copy (4 floats /
 * point, order within a ring = input order); ring_offsets[scan_lines+1].
 * Returns number of valid points, or <0 on bad scan_lines / lidar_type
 * (the reference logs ROS_ERROR_ONCE and emits nothing). */
int orc_split(const OrcParams* p, const float* pts, int n, int stride_f, int width, int height,
              int32_t* ring_of_point, float* rings_xyzi, int32_t* ring_offsets, int32_t* src_index);

/* A3+A4  extractFeatures / extractFeaturesFromRegion (src/feature_extractor.cc:181-313).
 * keys (optional, one double per ring-major point, NaN where not computed).
 * sort_mode 0: literal std::sort with the reference comparator
 *              (include/liodom/feature_extractor.h:57-59);
 *           1: total order (smoothness desc, index asc).
 * Emits edges in (ring, region, pick) order.  Returns edge count. */
int orc_extract(const OrcParams* p, const float* rings_xyzi, const int32_t* ring_offsets,
                float* edges_xyzi, int32_t* edge_ring, int32_t* edge_idx, double* keys,
                int sort_mode, int cap);

/* Convenience: split + extract. times_us[2] (optional) = {split, extract} wall us. */
int orc_extract_scan(const OrcParams* p, const float* pts, int n, int stride_f, int width, int height,
                     float* edges_xyzi, int cap, double* times_us);

/* A.1 pcl::transformPointCloud with a double 4x4 (row-major T). */
void orc_transform(const float* in_xyzi, int n, const double* T, float* out_xyzi);

/* A.2 exact 5-NN w.r.t. FLANN L2_Simple<float>; order (d2 asc, idx asc).
 * method 0: brute force, 1: kd-tree (leaf 15).  tie[E] (optional): 1 if the 5-set is
 * tie-ambiguous (equal d2 inside the set or d2[4]==d2 of the 6th). If M<5 missing
 * slots are idx=-1, d2=+inf. */
void orc_knn5(const float* map_xyzi, int M, const float* q_xyzi, int E, int method,
              int32_t* idx, float* d2, uint8_t* tie);

/* A8 addEdgeConstraints (src/laser_odometry.cc:300-366) up to the residual-block list.
 * T: row-major pose.  Outputs per edge: knn_idx[5], knn_d2[5], gate (bit0: d2[4]<1.0,
 * bit1: lambda2 > 3*lambda1), eig[3] ascending, q_world (transformed edge, 4 floats). */
void orc_associate(const float* edges_xyzi, int E, const double* T, const float* map_xyzi, int M,
                   int knn_method, int32_t* knn_idx, float* knn_d2, uint8_t* gate, double* eig,
                   float* q_world, uint8_t* tie);

/* A9 Point2LineFactor (include/liodom/factors.hpp:64-121): residual (3) and the 3x6
 * tangent Jacobian (row-major, cols = 3 quaternion-local then 3 translation) that
 * Ceres' autodiff + EigenQuaternionParameterization yields. q = (x,y,z,w). */
/* Eigenvalues (ascending) of a symmetric 3x3 (row-major 9 doubles): the cyclic Jacobi that stands in for
 * Eigen::SelfAdjointEigenSolver in the line gate (src/laser_odometry.cc:341-344). */
void orc_sym3_eigenvalues(const double* A9, double* w3);

void orc_factor(const double* c, const double* a, const double* b, double min_range, double max_range,
                const double* q, const double* t, double* r3, double* J18);

typedef struct {
  int iterations;          /* trust-region iterations executed (<=4) */
  int successful_steps;
  int termination;         /* 0 max-iter, 1 gradient tol, 2 parameter tol, 3 function tol, 4 no residuals, 5 failure */
  double initial_cost;
  double final_cost;
  int num_residual_blocks;
  int cost_evals, jac_evals;
} OrcSolveSummary;

/* A10 one ceres::Solve (src/laser_odometry.cc:201-218): HuberLoss(0.2),
 * EigenQuaternionParameterization, LM trust region, DENSE_QR, max 4 iterations,
 * Ceres 1.14 defaults.  cab: nblocks x 9 doubles (c, a, b).  q (x,y,z,w), t in/out.
 * linear_solver 0: Householder QR on [J;D] (as Ceres), 1: Cholesky on normal equations. */
void orc_solve(const double* cab, int nblocks, double min_range, double max_range,
               double* q, double* t, int linear_solver, OrcSolveSummary* sum);

/* LaserOdometer state machine (src/laser_odometry.cc:100-272 minus ROS/IMU). */
typedef struct OrcOdom OrcOdom;
OrcOdom* orc_odom_create(const OrcParams* p);
void orc_odom_destroy(OrcOdom* o);
/* Teacher forcing: overwrite odom_/prev_odom_ (row-major 4x4) and/or the window. */
void orc_odom_set_pose(OrcOdom* o, const double* odom, const double* prev_odom);
void orc_odom_get_pose(const OrcOdom* o, double* odom, double* prev_odom);
int orc_odom_window_size(const OrcOdom* o);
int orc_odom_window_frames(const OrcOdom* o);
void orc_odom_get_window(const OrcOdom* o, float* xyzi);
void orc_odom_set_window(OrcOdom* o, const float* xyzi, const int32_t* frame_sizes, int nframes);
void orc_odom_set_received_map(OrcOdom* o, const float* xyzi, int n);  /* SharedData::setLocalMap */
/* use_imu (src/laser_odometry.cc:152-183): roll/pitch of the predicted pose, expressed in base_link
 * through laser_to_base (row-major 4x4), replaced by the IMU orientation's. */
void orc_odom_set_imu(OrcOdom* o, int use_imu, const double* q_xyzw, const double* laser_to_base16);
/* tf round trip used by both: rpy3 = getRPY(Matrix3x3(q)), q_back = getRotation(setRPY(rpy3)). */
void orc_tf_rpy(const double* q_xyzw, double* rpy3, double* q_back_xyzw);
void orc_imu_override(const double* odom16, const double* imu_q_xyzw, const double* l2b16, double* out16);
/* publishOdom (src/laser_odometry.cc:395-446): out13 = orientation (x,y,z,w), position, twist linear, twist angular. */
void orc_publish_odom(const double* pose16, const double* prev_odom16, const double* l2b16, double delta_time, double* out13);
typedef struct {
  int n_edges;
  int n_map[2];
  int n_matches[2];
  OrcSolveSummary solve[2];
  double pred_pose[16];
  double times_us[4];      /* local map, associate(2x), solve(2x), window update */
} OrcFrameDiag;
/* One popFeatures() iteration. pose_out row-major 4x4 (odom_ after the frame). */
void orc_odom_process(OrcOdom* o, const float* edges_xyzi, int E, double* pose_out, OrcFrameDiag* diag);

/* LocalMapManager (src/laser_odometry.cc:24-69). */
typedef struct OrcLmap OrcLmap;
OrcLmap* orc_lmap_create(int max_frames);
void orc_lmap_destroy(OrcLmap* m);
void orc_lmap_add(OrcLmap* m, const float* xyzi, int n);
int orc_lmap_size(const OrcLmap* m);
int orc_lmap_frames(const OrcLmap* m);
void orc_lmap_get(const OrcLmap* m, float* xyzi);
void orc_lmap_set_max_frames(OrcLmap* m, int max_frames);

/* A.3 pcl::VoxelGrid<PointXYZI> (leaf cubic). Returns output count, -1 if the index
 * space would overflow int32 (PCL then returns the input unchanged). In-voxel
 * accumulation order = input order (PCL: unstable std::sort order). */
int orc_voxelgrid(const float* in_xyzi, int n, float leaf, float* out_xyzi);

/* Map (src/map.cc:70-189, include/liodom/map.h:58-116). */
typedef struct OrcMap OrcMap;
OrcMap* orc_map_create(double xy_size, double z_size, double resolution);
void orc_map_destroy(OrcMap* m);
void orc_map_update(OrcMap* m, const float* pts_xyzi, int n, const double* T);
int orc_map_size(const OrcMap* m);
int orc_map_num_cells(const OrcMap* m);
void orc_map_get(const OrcMap* m, float* xyzi);
/* cell i (creation order): key[3], count. */
void orc_map_cell_info(const OrcMap* m, int i, int32_t* key3, int32_t* count);
int orc_map_get_local(const OrcMap* m, const double* T, int cells_xy, int cells_z, float* xyzi, int cap);

/* pcl::fromROSMsg<PointXYZI> (call sites src/liodom_node.cc:43-44, :62-63; PCL 1.10
 * pcl/conversions.h fromPCLPointCloud2, third-party): per point, each matched FLOAT32 field is
 * memcpy'd from data + row * row_step + col * point_step + offset; an unmatched intensity
 * stays 0 (off_i < 0).  out_xyzi: width * height x 4 floats. */
void orc_decode_cloud2(const uint8_t* data, int width, int height, int point_step, int row_step,
                       int off_x, int off_y, int off_z, int off_i, float* out_xyzi);

/* Whole-path CPU baseline: for each of nframes scans (concatenated, counts in npts):
 * split -> extract -> odometry.  poses_out nframes x 16.  stage_us[5] accumulates
 * {split, extract, associate, solve, window}. Returns total edges. */
long orc_run_sequence(const OrcParams* p, const float* pts, const int32_t* npts, int nframes,
                      int stride_f, int width, int height, double* poses_out, double* stage_us);

#ifdef __cplusplus
}
#endif
#endif
