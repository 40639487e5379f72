// Synthetic LiDAR workload generator (test/bench harness, NOT part of the hot path).
//
// Produces the inputs SURVEY.md §8(d) specifies for configs C1..C5: a procedural,
// unbounded street-grid city (ground plane, building boxes, vertical poles, parked
// cars) ray-cast from HDL-64-shaped or OS1-128-shaped sensors along a smooth
// ground-truth trajectory.  Everything is a pure function of (seed, frame, beam,
// azimuth) through a counter-based hash, so results do not depend on threading.
//
// Beam tables are the inverse of the ring formulas the reference applies in
// src/feature_extractor.cc:130-134 (HDL-64) so that beam k lands in ring k.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kBlock = 40.0;      // city block pitch [m]
constexpr double kGroundZ = -1.73;   // ground plane in world z (sensor height 1.73 m)

inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t hash4(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
  return mix64(mix64(mix64(mix64(a) ^ b) ^ c) ^ d);
}
inline double u01(uint64_t h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

struct Box { double lo[3], hi[3]; int mat; };
struct Pole { double x, y, r, z0, z1; int mat; };
struct BlockPrims {
  double lo[3], hi[3];
  std::vector<Box> boxes;
  std::vector<Pole> poles;
};

// Contents of city block (bi, bj): deterministic in (seed, bi, bj).
void make_block(uint64_t seed, int bi, int bj, BlockPrims& out) {
  out.boxes.clear(); out.poles.clear();
  const double x0 = bi * kBlock, y0 = bj * kBlock;
  uint64_t h = hash4(seed, (uint64_t)(int64_t)bi, (uint64_t)(int64_t)bj, 0x51ED);
  auto rnd = [&](int k) { return u01(hash4(h, k, 0xA11CE, 7)); };
  // Buildings: footprint inset 7 m from the block edge (14 m wide streets). Either one
  // building or two separated by an alley, so that corners and depth steps exist.
  const double in = 7.0;
  int layout = (int)(rnd(0) * 3.0);
  if (layout == 0) {
    Box b; b.lo[0] = x0 + in + rnd(1) * 2; b.hi[0] = x0 + kBlock - in - rnd(2) * 2;
    b.lo[1] = y0 + in + rnd(3) * 2; b.hi[1] = y0 + kBlock - in - rnd(4) * 2;
    b.lo[2] = kGroundZ; b.hi[2] = kGroundZ + 6 + rnd(5) * 18; b.mat = 80 + (int)(rnd(6) * 40);
    out.boxes.push_back(b);
  } else if (layout == 1) {  // split along x
    double cut = 14 + rnd(1) * 12, gap = 2.0 + rnd(2) * 2.5;
    Box a; a.lo[0] = x0 + in; a.hi[0] = x0 + cut; a.lo[1] = y0 + in + rnd(3) * 3; a.hi[1] = y0 + kBlock - in;
    a.lo[2] = kGroundZ; a.hi[2] = kGroundZ + 5 + rnd(4) * 12; a.mat = 80 + (int)(rnd(5) * 40);
    Box b; b.lo[0] = x0 + cut + gap; b.hi[0] = x0 + kBlock - in; b.lo[1] = y0 + in; b.hi[1] = y0 + kBlock - in - rnd(6) * 3;
    b.lo[2] = kGroundZ; b.hi[2] = kGroundZ + 8 + rnd(7) * 16; b.mat = 80 + (int)(rnd(8) * 40);
    out.boxes.push_back(a); out.boxes.push_back(b);
  } else {  // split along y
    double cut = 14 + rnd(1) * 12, gap = 2.0 + rnd(2) * 2.5;
    Box a; a.lo[1] = y0 + in; a.hi[1] = y0 + cut; a.lo[0] = x0 + in + rnd(3) * 3; a.hi[0] = x0 + kBlock - in;
    a.lo[2] = kGroundZ; a.hi[2] = kGroundZ + 5 + rnd(4) * 12; a.mat = 80 + (int)(rnd(5) * 40);
    Box b; b.lo[1] = y0 + cut + gap; b.hi[1] = y0 + kBlock - in; b.lo[0] = x0 + in; b.hi[0] = x0 + kBlock - in - rnd(6) * 3;
    b.lo[2] = kGroundZ; b.hi[2] = kGroundZ + 8 + rnd(7) * 16; b.mat = 80 + (int)(rnd(8) * 40);
    out.boxes.push_back(a); out.boxes.push_back(b);
  }
  // Poles on the kerb line, 5.5 m in from the block edge.
  const double kerb = 5.5;
  const double along[3] = {6.0, 20.0, 34.0};
  int pk = 20;
  for (int side = 0; side < 4; ++side) {
    for (int a = 0; a < 3; ++a) {
      if (rnd(pk++) < 0.25) { pk += 3; continue; }
      Pole p; double s = along[a] + (rnd(pk++) - 0.5) * 3.0;
      if (side == 0) { p.x = x0 + s; p.y = y0 + kerb; }
      else if (side == 1) { p.x = x0 + s; p.y = y0 + kBlock - kerb; }
      else if (side == 2) { p.x = x0 + kerb; p.y = y0 + s; }
      else { p.x = x0 + kBlock - kerb; p.y = y0 + s; }
      p.r = 0.15; p.z0 = kGroundZ; p.z1 = kGroundZ + 5.0 + rnd(pk++) * 4.0; p.mat = 200;
      pk++;
      out.poles.push_back(p);
    }
  }
  // Parked cars, 4.2 m in from the block edge (outside the +-2.5 m driving corridor).
  for (int c = 0; c < 3; ++c) {
    if (rnd(60 + c * 5) < 0.45) continue;
    int side = (int)(rnd(61 + c * 5) * 4.0);
    double s = 9.0 + rnd(62 + c * 5) * 22.0;
    double L = 4.4, W = 1.8, H = 1.5, off = 3.6;
    Box b; b.lo[2] = kGroundZ; b.hi[2] = kGroundZ + H; b.mat = 150;
    if (side == 0) { b.lo[0] = x0 + s - L / 2; b.hi[0] = x0 + s + L / 2; b.lo[1] = y0 + off; b.hi[1] = y0 + off + W; }
    else if (side == 1) { b.lo[0] = x0 + s - L / 2; b.hi[0] = x0 + s + L / 2; b.hi[1] = y0 + kBlock - off; b.lo[1] = b.hi[1] - W; }
    else if (side == 2) { b.lo[1] = y0 + s - L / 2; b.hi[1] = y0 + s + L / 2; b.lo[0] = x0 + off; b.hi[0] = x0 + off + W; }
    else { b.lo[1] = y0 + s - L / 2; b.hi[1] = y0 + s + L / 2; b.hi[0] = x0 + kBlock - off; b.lo[0] = b.hi[0] - W; }
    out.boxes.push_back(b);
  }
  out.lo[0] = x0; out.lo[1] = y0; out.lo[2] = kGroundZ;
  out.hi[0] = x0 + kBlock; out.hi[1] = y0 + kBlock; out.hi[2] = kGroundZ + 30.0;
}

inline bool ray_box(const double o[3], const double inv[3], const double lo[3], const double hi[3],
                    double tmax, double& tnear) {
  double t0 = 0.0, t1 = tmax;
  for (int k = 0; k < 3; ++k) {
    double a = (lo[k] - o[k]) * inv[k], b = (hi[k] - o[k]) * inv[k];
    if (a > b) std::swap(a, b);
    if (a > t0) t0 = a;
    if (b < t1) t1 = b;
    if (t0 > t1) return false;
  }
  tnear = t0;
  return true;
}

inline bool ray_pole(const double o[3], const double d[3], const Pole& p, double tmax, double& t) {
  double ox = o[0] - p.x, oy = o[1] - p.y;
  double a = d[0] * d[0] + d[1] * d[1];
  if (a < 1e-18) return false;
  double b = ox * d[0] + oy * d[1];
  double c = ox * ox + oy * oy - p.r * p.r;
  double disc = b * b - a * c;
  if (disc < 0) return false;
  double tt = (-b - std::sqrt(disc)) / a;
  if (tt <= 0 || tt >= tmax) return false;
  double z = o[2] + tt * d[2];
  if (z < p.z0 || z > p.z1) return false;
  t = tt;
  return true;
}

void rot_rpy(double roll, double pitch, double yaw, double R[9]) {
  double cr = std::cos(roll), sr = std::sin(roll), cp = std::cos(pitch), sp = std::sin(pitch);
  double cy = std::cos(yaw), sy = std::sin(yaw);
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

}  // namespace

extern "C" {

struct SynthSensor {
  int model;           // 0: HDL-64 beam table, 1: uniform +-22.5 deg (OS1), rows = beams
  int beams;           // 64 or 128
  int az_steps;        // 1875 / 2048 / 15625
  int order;           // 0 ring-major, 1 azimuth-major (firing order), 2 organised row-major with (0,0,0) holes
  double noise_sigma;  // range noise [m]
  double max_ray;      // rays longer than this give no return
};

int synth_num_rays(const SynthSensor* s) { return s->beams * s->az_steps; }

// Ground-truth pose of `frame` (row-major 4x4, world_from_sensor).
//  traj 0: ~1 m/frame along the y=0 street with a +-2 m weave (C1/C2/C3/C5)
//  traj 1: closed rounded-square circuit, 1 km sides, 1 m/frame (C4: 4000 frames)
void synth_pose(uint64_t seed, int traj, int frame, double T[16]) {
  double ph = u01(hash4(seed, 0x7A11, 3, 9)) * 2 * kPi;
  double s = (double)frame;
  // traj 0 starts from rest (KITTI-like): speed ramps 0 -> 1 m/frame with a 5-frame time constant,
  // so the reference's constant-velocity prediction (src/laser_odometry.cc:148-150) is meaningful.
  if (traj == 0) s = s - 5.0 * (1.0 - std::exp(-s / 5.0));
  double x, y, yaw;
  if (traj == 0) {
    double w = 2 * kPi / 60.0;
    x = s; y = 2.0 * std::sin(w * s + ph);
    yaw = std::atan(2.0 * w * std::cos(w * s + ph));
  } else {
    // Rounded square, side S, corner radius r; perimeter = 4*(S-2r) + 2*pi*r.
    const double S = 1000.0, r = 20.0;
    const double straight = S - 2 * r, arc = 0.5 * kPi * r, per = 4 * (straight + arc);
    double u = std::fmod(s, per);
    int leg = (int)(u / (straight + arc));
    double v = u - leg * (straight + arc);
    double lx, ly, lyaw;  // local leg frame: start at (r,0) heading +x
    if (v < straight) { lx = r + v; ly = 0; lyaw = 0; }
    else { double a = (v - straight) / r; lx = S - r + r * std::sin(a); ly = r - r * std::cos(a); lyaw = a; }
    double c = std::cos(leg * 0.5 * kPi), sn = std::sin(leg * 0.5 * kPi);
    // leg origins: (0,0),(S,0),(S,S),(0,S) rotated by leg*90deg
    double ox = (leg == 1 || leg == 2) ? S : 0, oy = (leg >= 2) ? S : 0;
    x = ox + c * lx - sn * ly; y = oy + sn * lx + c * ly; yaw = lyaw + leg * 0.5 * kPi;
  }
  double z = 0.05 * std::sin(2 * kPi * s / 25.0 + ph);
  double pitch = (0.4 * kPi / 180) * std::sin(2 * kPi * s / 18.0 + 0.5 * ph);
  double roll = (0.3 * kPi / 180) * std::sin(2 * kPi * s / 23.0 + 1.0 + ph);
  double R[9]; rot_rpy(roll, pitch, yaw, R);
  T[0] = R[0]; T[1] = R[1]; T[2] = R[2]; T[3] = x;
  T[4] = R[3]; T[5] = R[4]; T[6] = R[5]; T[7] = y;
  T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = z;
  T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

// Ray-cast one scan. `xyzi` must hold 4*synth_num_rays floats. Returns the number of
// points written (== num_rays for order 2, where missing returns are (0,0,0,0)).
int synth_scan(const SynthSensor* sen, uint64_t seed, int frame, const double T[16], float* xyzi) {
  const int B = sen->beams, A = sen->az_steps;
  std::vector<double> ce(B), se(B);
  for (int k = 0; k < B; ++k) {
    double el;
    if (sen->model == 0) el = (k < 32) ? (2.0 - k / 3.0) : (-8.83 - (k - 32) / 2.0);
    else el = 22.5 - 45.0 * k / (double)(B - 1);
    ce[k] = std::cos(el * kPi / 180); se[k] = std::sin(el * kPi / 180);
  }
  const double o[3] = {T[3], T[7], T[11]};
  // Blocks that can be hit.
  std::vector<BlockPrims> blocks;
  {
    int bi0 = (int)std::floor((o[0] - sen->max_ray) / kBlock), bi1 = (int)std::floor((o[0] + sen->max_ray) / kBlock);
    int bj0 = (int)std::floor((o[1] - sen->max_ray) / kBlock), bj1 = (int)std::floor((o[1] + sen->max_ray) / kBlock);
    for (int bi = bi0; bi <= bi1; ++bi)
      for (int bj = bj0; bj <= bj1; ++bj) {
        double cx = std::max(bi * kBlock, std::min(o[0], bi * kBlock + kBlock));
        double cy = std::max(bj * kBlock, std::min(o[1], bj * kBlock + kBlock));
        if ((cx - o[0]) * (cx - o[0]) + (cy - o[1]) * (cy - o[1]) > sen->max_ray * sen->max_ray) continue;
        blocks.emplace_back();
        make_block(seed, bi, bj, blocks.back());
      }
  }
  const long nr = (long)B * A;
  std::vector<float> tmp((size_t)nr * 4);
  std::vector<uint8_t> hit((size_t)nr);
#pragma omp parallel for schedule(static)
  for (long r = 0; r < nr; ++r) {
    int k, a;
    if (sen->order == 1) { a = (int)(r / B); k = (int)(r % B); }
    else { k = (int)(r / A); a = (int)(r % A); }
    double az = 2 * kPi * a / (double)A;
    double ds[3] = {ce[k] * std::cos(az), ce[k] * std::sin(az), se[k]};
    double d[3] = {T[0] * ds[0] + T[1] * ds[1] + T[2] * ds[2],
                   T[4] * ds[0] + T[5] * ds[1] + T[6] * ds[2],
                   T[8] * ds[0] + T[9] * ds[1] + T[10] * ds[2]};
    double inv[3];
    for (int q = 0; q < 3; ++q) inv[q] = 1.0 / (std::fabs(d[q]) > 1e-12 ? d[q] : (d[q] < 0 ? -1e-12 : 1e-12));
    double best = sen->max_ray; int mat = -1;
    if (d[2] < -1e-9) { double t = (kGroundZ - o[2]) / d[2]; if (t > 0 && t < best) { best = t; mat = 10; } }
    for (const BlockPrims& bp : blocks) {
      double tb;
      if (!ray_box(o, inv, bp.lo, bp.hi, best, tb)) continue;
      for (const Box& bx : bp.boxes) { double t; if (ray_box(o, inv, bx.lo, bx.hi, best, t) && t > 0 && t < best) { best = t; mat = bx.mat; } }
      for (const Pole& pl : bp.poles) { double t; if (ray_pole(o, d, pl, best, t)) { best = t; mat = pl.mat; } }
    }
    float* p = &tmp[(size_t)r * 4];
    if (mat < 0) { hit[r] = 0; p[0] = p[1] = p[2] = p[3] = 0.f; continue; }
    uint64_t h = hash4(seed ^ 0xD1CEull, (uint64_t)frame, (uint64_t)k, (uint64_t)a);
    double u1 = u01(h), u2 = u01(mix64(h));
    double g = std::sqrt(-2.0 * std::log(u1)) * std::cos(2 * kPi * u2);
    double rng = best + sen->noise_sigma * g;
    p[0] = (float)(ds[0] * rng); p[1] = (float)(ds[1] * rng); p[2] = (float)(ds[2] * rng);
    p[3] = (float)(mat / 255.0);
    hit[r] = 1;
  }
  if (sen->order == 2) { std::memcpy(xyzi, tmp.data(), (size_t)nr * 16); return (int)nr; }
  int n = 0;
  for (long r = 0; r < nr; ++r)
    if (hit[r]) { std::memcpy(xyzi + (size_t)n * 4, &tmp[(size_t)r * 4], 16); ++n; }
  return n;
}

}  // extern "C"
