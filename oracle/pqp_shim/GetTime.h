// TEST INFRASTRUCTURE (oracle) -- PQP's GetTime.h: wall clock in seconds
// (used only by off-path discrete queries, C2A/src/C2A_PQP.cpp:916,964).
#ifndef PQP_SHIM_GETTIME_H
#define PQP_SHIM_GETTIME_H
#include <time.h>
inline double GetTime()
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
#endif
