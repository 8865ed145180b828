// gicp_settings.hpp — enum surface of the rot_gicp API (drop-in for the reference header of the same path,
// include/rot_gicp/gicp/gicp_settings.hpp:6-12). Enumerator ORDER is part of the ABI: the values are passed
// straight through the C boundary (include/rolo_gicp.h).
#ifndef ROLO_B200_GICP_SETTINGS_HPP
#define ROLO_B200_GICP_SETTINGS_HPP

namespace fast_gicp {

enum class RegularizationMethod { NONE, MIN_EIG, NORMALIZED_MIN_EIG, PLANE, FROBENIUS, PLANE_S };
enum class NeighborSearchMethod { DIRECT27, DIRECT7, DIRECT1, DIRECT_RADIUS };
enum class VoxelAccumulationMode { ADDITIVE, ADDITIVE_WEIGHTED, MULTIPLICATIVE };
enum class VoxelType { POLAR, UNIFORM };
enum class LSQ_OPTIMIZER_TYPE { GaussNewton, LevenbergMarquardt, SO3_LevenbergMarquardt };

}  // namespace fast_gicp

#endif
