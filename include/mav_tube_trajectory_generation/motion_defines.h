// Derivative-order constants of the reference (motion_defines.h:28-47).
#ifndef MTG_SHIM_MOTION_DEFINES_H_
#define MTG_SHIM_MOTION_DEFINES_H_

#include <string>

namespace mav_trajectory_generation {

namespace derivative_order {
static constexpr int POSITION = 0;
static constexpr int VELOCITY = 1;
static constexpr int ACCELERATION = 2;
static constexpr int JERK = 3;
static constexpr int SNAP = 4;

static constexpr int ORIENTATION = 0;
static constexpr int ANGULAR_VELOCITY = 1;
static constexpr int ANGULAR_ACCELERATION = 2;

static constexpr int INVALID = -1;
static constexpr int kINVALID = -1;
}  // namespace derivative_order

inline std::string positionDerivativeToString(int derivative) {
  static const char* names[] = {"position", "velocity", "acceleration", "jerk", "snap"};
  return (derivative >= 0 && derivative <= 4) ? names[derivative] : "invalid";
}
inline int positionDerivativeToInt(const std::string& s) {
  for (int i = 0; i <= 4; ++i)
    if (s == positionDerivativeToString(i)) return i;
  return derivative_order::INVALID;
}
inline std::string orintationDerivativeToString(int derivative) {  // (sic) the reference's spelling
  static const char* names[] = {"orientation", "angular_velocity", "angular_acceleration"};
  return (derivative >= 0 && derivative <= 2) ? names[derivative] : "invalid";
}
inline int orientationDerivativeToInt(const std::string& s) {
  for (int i = 0; i <= 2; ++i)
    if (s == orintationDerivativeToString(i)) return i;
  return derivative_order::INVALID;
}

}  // namespace mav_trajectory_generation
#endif
