/* liodom::Map — GPU hash-grid map behind the reference's class interface
 * (include/liodom/map.h:94-116, src/map.cc:70-211). Cell / HashKey are internal to the CUDA
 * library (coarse-cell hash + radix-sorted voxel centroids), so they are not re-declared. */
#ifndef INCLUDE_LIODOM_MAP_H
#define INCLUDE_LIODOM_MAP_H

#include <liodom/defs.h>

struct liodom_map;

namespace liodom {

class Map {
 public:
  explicit Map(const double xy_size, const double z_size, const double res);
  virtual ~Map();
  Map(const Map&) = delete;
  void operator=(const Map&) = delete;

  void updateMap(const PointCloud::Ptr& pc_in, const Isometry3d& pose);
  PointCloud::Ptr getMap();
  PointCloud::Ptr getLocalMap(const Isometry3d& pose, int cells_xy = 2, int cells_z = 1);
  /* The reference computes -sum p log p over std::unordered_map bucket occupancies
   * (src/map.cc:191-211) — implementation-defined, and its only caller is commented out
   * (src/liodom_mapping_node.cc:103-105).  Here p = points of a cell / total points. */
  double getMapEntropy();

 private:
  liodom_map* map_;
};

}  // namespace liodom
#endif  // INCLUDE_LIODOM_MAP_H
