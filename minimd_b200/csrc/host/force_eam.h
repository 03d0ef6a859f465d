// ForceEAM (ref/force_eam.h): funcfl reader, re-gridding and spline construction stay on the host
// (ref/force_eam.cpp:505-793); the density / embedding / pair passes and the fp halo run on the device.
#pragma once
#include <string>
#include <vector>

#include "force.h"

class ForceEAM : public Force {
 public:
  MMD_float cutmax;
  MMD_int nrho, nr;
  MMD_int nrho_tot, nr_tot;
  MMD_float dr, rdr, drho, rdrho;
  MMD_float *rhor_spline, *frho_spline, *z2r_spline;  // 7 coefficients per knot, replicated per type pair

  std::string potential_file;  // default "Cu_u6.eam" in cwd (ref/force_eam.cpp:77)

  explicit ForceEAM(int ntypes_);
  virtual ~ForceEAM();
  int setup(Atom& atom) override;
  void compute(Atom& atom, Neighbor& neighbor, Comm& comm, int me) override;

 private:
  struct Funcfl {
    int nrho = 0, nr = 0;
    double drho = 0, dr = 0, cut = 0, mass = 0;
    std::vector<MMD_float> frho, rhor, zr;  // 1-based after the shift (ref/force_eam.cpp:575-579)
  } funcfl;
  std::vector<MMD_float> frho_, rhor_, z2r_;
  std::vector<MMD_float> rhor_sp_, frho_sp_, z2r_sp_;

  int read_file(const char* filename, int me);
  void file2array();
  void array2spline();
  void interpolate(MMD_int n, MMD_float delta, const MMD_float* f, MMD_float* spline);
};
