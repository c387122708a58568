// stand-in for the reference's timer collection: timing is not needed by the encoder check
#ifndef SPARSEX_INTERNALS_TIMER_COLLECTION_HPP
#define SPARSEX_INTERNALS_TIMER_COLLECTION_HPP
#include <ostream>
namespace sparsex { namespace timing {
class TimerCollection {
 public:
  void CreateTimer(const char *, const char *) {}
  void StartTimer(const char *) {}
  void PauseTimer(const char *) {}
  void PrintAllTimers(std::ostream &) const {}
};
} }
#endif
