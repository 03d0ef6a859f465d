// ForceLJ (ref/force_lj.h, ref/force_lj.cpp:41-113)
#pragma once
#include "force.h"

class ForceLJ : public Force {
 public:
  explicit ForceLJ(int ntypes_);
  virtual ~ForceLJ();
  int setup(Atom& atom) override;
  void compute(Atom& atom, Neighbor& neighbor, Comm& comm, int me) override;
};
