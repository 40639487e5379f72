// refshim forwarding header (TEST INFRASTRUCTURE ONLY): tf::transformTFToEigen for src/liodom_mapping_node.cc:65-66
#pragma once
#include "../refshim_ros.h"
#include "../refshim_eigen.h"
namespace tf {
inline void transformTFToEigen(const Transform& t, Eigen::Isometry3d& e) {
  const Quaternion q = t.getRotation();
  Matrix3x3 m(q);
  Quaternion back;   // Matrix3x3(q) row access through a second conversion is not exposed: rebuild from the quaternion
  (void)back;
  const double d = q.length2(), s = 2.0 / d;
  const double xs = q.x() * s, ys = q.y() * s, zs = q.z() * s, wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
  const double xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs, yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
  e = Eigen::Isometry3d::Identity();
  e.matrix()(0, 0) = 1.0 - (yy + zz); e.matrix()(0, 1) = xy - wz; e.matrix()(0, 2) = xz + wy;
  e.matrix()(1, 0) = xy + wz; e.matrix()(1, 1) = 1.0 - (xx + zz); e.matrix()(1, 2) = yz - wx;
  e.matrix()(2, 0) = xz - wy; e.matrix()(2, 1) = yz + wx; e.matrix()(2, 2) = 1.0 - (xx + yy);
  e.matrix()(0, 3) = t.getOrigin().x(); e.matrix()(1, 3) = t.getOrigin().y(); e.matrix()(2, 3) = t.getOrigin().z();
}
}  // namespace tf
