// Extremum of a derivative magnitude (extremum.h:30-44): ordered by value only, time relative to
// the start of the segment it lies in.
#ifndef MTG_SHIM_EXTREMUM_H_
#define MTG_SHIM_EXTREMUM_H_

#include <ostream>

namespace mav_trajectory_generation {

struct Extremum {
  Extremum() : time(0.0), value(0.0), segment_idx(0) {}
  Extremum(double t, double v, int idx) : time(t), value(v), segment_idx(idx) {}
  bool operator<(const Extremum& rhs) const { return value < rhs.value; }
  bool operator>(const Extremum& rhs) const { return value > rhs.value; }

  double time;
  double value;
  int segment_idx;
};

inline std::ostream& operator<<(std::ostream& os, const Extremum& e) {
  return os << "time: " << e.time << ", value: " << e.value << ", segment idx: " << e.segment_idx << std::endl;
}

}  // namespace mav_trajectory_generation
#endif
