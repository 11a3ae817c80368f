// Wall-clock timers and the three timer groups whose names end up as CSV columns of CalsReport / AlsReport
// (counterpart of the reference's include/timer.h; the group and member names are part of the API).  On the B200 path
// the per-phase numbers come from CUDA events (cals_b200_report) and are loaded into these objects with Timer::set.
#ifndef CALS_B200_TIMER_H
#define CALS_B200_TIMER_H

#include <array>
#include <chrono>
#include <string>

namespace cals {

class Timer {
  using clock = std::chrono::steady_clock;
  clock::time_point begin_{};
  double seconds_{-1.0}; // negative: never stopped, reads as 0

public:
  void start() { begin_ = clock::now(); }
  void stop() { seconds_ = std::chrono::duration<double>(clock::now() - begin_).count(); }
  void reset() { seconds_ = 0.0; }
  void set(double seconds) { seconds_ = seconds; } // extension: a time measured on the device
  [[nodiscard]] double get_time() const { return seconds_ < 0.0 ? 0.0 : seconds_; }
};

namespace detail {
// N timers addressed by an enum, plus the column names that go with them.
template <int N> struct TimerGroup {
  std::array<Timer, N> timers{};
  std::string names[N];
  Timer &operator[](int which) { return timers[static_cast<size_t>(which)]; }
  const Timer &operator[](int which) const { return timers[static_cast<size_t>(which)]; }

protected:
  TimerGroup(std::initializer_list<const char *> labels) {
    int i = 0;
    for (const char *l : labels)
      names[i++] = l;
  }
};
} // namespace detail

// Inside one MTTKRP (the reference's KRP + GEMM / two-step split; one fused kernel here, so these stay 0).
struct MttkrpTimers : detail::TimerGroup<4> {
  enum TIMERS { MT_KRP = 0, MT_GEMM, TS_GEMM, TS_GEMV, LENGTH };
  MttkrpTimers() : TimerGroup({"MT_KRP", "MT_GEMM", "TS_GEMM", "TS_GEMV"}) {}
};

// Per mode of one ALS iteration.
struct ModeTimers : detail::TimerGroup<2> {
  enum TIMERS { MTTKRP = 0, UPDATE, LENGTH };
  ModeTimers() : TimerGroup({"TOTAL_MTTKRP", "UPDATE"}) {}
};

// Per ALS / CALS iteration.
struct AlsTimers : detail::TimerGroup<5> {
  enum TIMERS { ITERATION = 0, DEFRAGMENTATION, ERROR, LINE_SEARCH, G_COPY, LENGTH };
  AlsTimers() : TimerGroup({"ITERATION", "DEFRAGMENTATION", "ERROR", "LINESEARCH", "G_COPY"}) {}
};

} // namespace cals
#endif
