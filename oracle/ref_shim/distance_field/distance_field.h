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
namespace distance_field {
class PropagationDistanceField {
 public:
  enum VoxelType { F32 = 0, U8_SQ = 1, U16_SQ = 2 };  // include/stomp_b200.h stomp_voxel_dtype
  PropagationDistanceField(const void* voxels, int nx, int ny, int nz, const double origin[3], double resolution, int voxel_type)
      : vox_(voxels), nx_(nx), ny_(ny), nz_(nz), res_(resolution), type_(voxel_type) {
    for (int i = 0; i < 3; ++i) origin_[i] = origin[i];
  }
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
};
}
#endif
