// Wall-clock timers with the reference's names (reference include/timer.h:8-54), so CalsReport / AlsReport CSV
// columns keep their meaning.  On the B200 path the per-phase numbers come from CUDA events (cals_b200_report).
#ifndef CALS_B200_TIMER_H
#define CALS_B200_TIMER_H

#include <chrono>
#include <string>

namespace cals {

class Timer {
  using clock = std::chrono::steady_clock;
  clock::time_point begin_{};
  double seconds_{-1.0}; // < 0: never stopped

public:
  void start() { begin_ = clock::now(); }
  void stop() { seconds_ = std::chrono::duration<double>(clock::now() - begin_).count(); }
  void reset() { seconds_ = 0.0; }
  void set(double seconds) { seconds_ = seconds; } // extension: load a device-measured time
  [[nodiscard]] double get_time() const { return seconds_ < 0.0 ? 0.0 : seconds_; }
};

namespace detail {
template <int N> struct TimerSet {
  Timer timers[N];
  Timer &operator[](int i) { return timers[i]; }
  const Timer &operator[](int i) const { return timers[i]; }
};
} // namespace detail

struct MttkrpTimers : detail::TimerSet<4> {
  enum TIMERS { MT_KRP = 0, MT_GEMM, TS_GEMM, TS_GEMV, LENGTH };
  std::string names[LENGTH] = {"MT_KRP", "MT_GEMM", "TS_GEMM", "TS_GEMV"};
};

struct ModeTimers : detail::TimerSet<2> {
  enum TIMERS { MTTKRP = 0, UPDATE, LENGTH };
  std::string names[LENGTH] = {"TOTAL_MTTKRP", "UPDATE"};
};

struct AlsTimers : detail::TimerSet<5> {
  enum TIMERS { ITERATION = 0, DEFRAGMENTATION, ERROR, LINE_SEARCH, G_COPY, LENGTH };
  std::string names[LENGTH] = {"ITERATION", "DEFRAGMENTATION", "ERROR", "LINESEARCH", "G_COPY"};
};

} // namespace cals
#endif
