// oracle/ref_shim/distance_field/distance_field.h — stand-in for ROS distance_field::PropagationDistanceField as far as
// StompCollisionSpace::getDistanceGradient uses it (include/stomp_motion_planner/stomp_collision_space.h:187-191).
// Third-party semantics restated from the published package, TEST INFRASTRUCTURE:
//   cell = int(round((x - origin) / resolution)) per axis;  a cell on or outside the outermost layer -> distance 0,
//   gradient 0;  distance = sqrt(squared cell distance) * resolution;  gradient = central difference / (2 resolution).
// The voxel array is supplied by the test (x-major [nx][ny][nz], the same array the engine and the oracle receive).
#ifndef STOMP_REF_SHIM_DISTANCE_FIELD
#define STOMP_REF_SHIM_DISTANCE_FIELD
#include <cmath>
#include <cstdint>
#include <vector>
#include <LinearMath/bullet_shim.h>
namespace distance_field {
class PropagationDistanceField {
 public:
  enum VoxelType { F32 = 0, U8_SQ = 1, U16_SQ = 2 };  // include/stomp_b200.h stomp_voxel_dtype
  PropagationDistanceField(const void* voxels, int nx, int ny, int nz, const double origin[3], double resolution, int voxel_type)
      : vox_(voxels), nx_(nx), ny_(ny), nz_(nz), res_(resolution), type_(voxel_type) {
    for (int i = 0; i < 3; ++i) origin_[i] = origin[i];
  }
  // the constructor StompCollisionSpace::init uses (src/stomp_collision_space.cpp:83): an empty field of
  // int(size / resolution) cells per axis.  Only the OCCUPANCY side of the field is restated for this form (which cells
  // addPointsToField marks); the distance propagation itself is third-party and is not run by the harness.
  PropagationDistanceField(double size_x, double size_y, double size_z, double resolution, double origin_x, double origin_y,
                           double origin_z, double max_distance)
      : vox_(0), nx_(int(size_x / resolution)), ny_(int(size_y / resolution)), nz_(int(size_z / resolution)), res_(resolution),
        type_(U8_SQ), max_distance_(max_distance) {
    origin_[0] = origin_x; origin_[1] = origin_y; origin_[2] = origin_z;
    reset();
  }
  void reset() { occupied_.assign(size_t(nx_) * ny_ * nz_, 0); points_.clear(); }
  void addPointsToField(const std::vector<btVector3>& points) {
    for (size_t i = 0; i < points.size(); ++i) {
      points_.push_back(points[i]);
      int x = getCellFromLocation(0, points[i].x()), y = getCellFromLocation(1, points[i].y()), z = getCellFromLocation(2, points[i].z());
      if (x < 0 || y < 0 || z < 0 || x >= nx_ || y >= ny_ || z >= nz_) continue;   // VoxelGrid::isCellValid
      occupied_[(size_t(x) * ny_ + y) * nz_ + z] = 1;
    }
  }
  template <typename A, typename B> void visualize(double, double, const A&, const B&) const {}
  const std::vector<uint8_t>& occupied() const { return occupied_; }
  const std::vector<btVector3>& points() const { return points_; }
  int dim(int a) const { return a == 0 ? nx_ : (a == 1 ? ny_ : nz_); }
  int getCellFromLocation(int dim, double loc) const { return int(round((loc - origin_[dim]) / res_)); }
  double getDistanceFromCell(int x, int y, int z) const {
    size_t i = (size_t(x) * ny_ + y) * nz_ + z;
    switch (type_) {
      case U8_SQ: return std::sqrt(double(static_cast<const uint8_t*>(vox_)[i])) * res_;
      case U16_SQ: return std::sqrt(double(static_cast<const uint16_t*>(vox_)[i])) * res_;
      default: return double(static_cast<const float*>(vox_)[i]);
    }
  }
  double getDistanceGradient(double x, double y, double z, double& gx, double& gy, double& gz) const {
    int cx = getCellFromLocation(0, x), cy = getCellFromLocation(1, y), cz = getCellFromLocation(2, z);
    last_cell_[0] = cx; last_cell_[1] = cy; last_cell_[2] = cz;
    if (cx < 1 || cy < 1 || cz < 1 || cx >= nx_ - 1 || cy >= ny_ - 1 || cz >= nz_ - 1) {
      gx = gy = gz = 0.0;
      return 0.0;
    }
    double inv_twice_resolution = 1.0 / (2.0 * res_);
    gx = (getDistanceFromCell(cx + 1, cy, cz) - getDistanceFromCell(cx - 1, cy, cz)) * inv_twice_resolution;
    gy = (getDistanceFromCell(cx, cy + 1, cz) - getDistanceFromCell(cx, cy - 1, cz)) * inv_twice_resolution;
    gz = (getDistanceFromCell(cx, cy, cz + 1) - getDistanceFromCell(cx, cy, cz - 1)) * inv_twice_resolution;
    return getDistanceFromCell(cx, cy, cz);
  }
  mutable int last_cell_[3];  // parity tap for the integer work (read by oracle/ref_driver.cpp)
 private:
  const void* vox_;
  int nx_, ny_, nz_;
  double origin_[3], res_;
  int type_;
  double max_distance_ = 0.0;
  std::vector<uint8_t> occupied_;
  std::vector<btVector3> points_;
};
}
#endif
