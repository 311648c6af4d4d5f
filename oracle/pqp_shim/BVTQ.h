// TEST INFRASTRUCTURE (oracle) -- clean-room stand-in for PQP's BVTQ.h (bounded
// min-priority queue of BV tests).  Only dead / off-path code in the reference
// uses it (C2A/src/C2A.cpp:438, C2A/src/C2A_PQP.cpp:629,1611); the CCD path
// ignores qsize (C2A/src/C2A.cpp:1987-1995).
#ifndef PQP_SHIM_BVTQ_H
#define PQP_SHIM_BVTQ_H
#include <vector>
#include "PQP_Compile.h"

struct BVT
{
  PQP_REAL d;       // distance between the bvs
  int b1, b2;       // bv indices
  PQP_REAL R[3][3]; // relative rotation
  PQP_REAL T[3];    // relative translation
  int pindex;
};

class BVTQ
{
  std::vector<BVT> q_;
  int size_;
public:
  BVTQ(int sz) : size_(sz > 2 ? sz : 2) { q_.reserve(size_); }
  int Empty() { return q_.empty(); }
  int GetNumTests() { return (int)q_.size(); }
  int GetSize() { return size_; }
  int Full() { return (int)q_.size() >= size_; }
  PQP_REAL MinTest()
  {
    PQP_REAL m = q_[0].d;
    for (size_t i = 1; i < q_.size(); i++) if (q_[i].d < m) m = q_[i].d;
    return m;
  }
  BVT ExtractMinTest()
  {
    size_t k = 0;
    for (size_t i = 1; i < q_.size(); i++) if (q_[i].d < q_[k].d) k = i;
    BVT t = q_[k];
    q_.erase(q_.begin() + k);
    return t;
  }
  void AddTest(BVT &t) { q_.push_back(t); }
};
#endif
