// Simulation: the body of the reference's main() (ref/ljs.cpp:61-504) as an object, so that the
// stand-alone driver (ljs.cpp) and embedders (include/minimd_host.h, used by bench.py and the
// tests) run exactly the same sequence.
#pragma once
#include <string>
#include <vector>

#include "atom.h"
#include "comm.h"
#include "force.h"
#include "integrate.h"
#include "ljs.h"
#include "neighbor.h"
#include "thermo.h"
#include "timer.h"
#include "world.h"

struct Options {  // argv of ref/ljs.cpp:112-261
  std::string input_file;  // empty -> in.lj.miniMD
  int num_threads = 1;
  int teams = 1;
  int num_steps = -1;
  int system_size = -1;
  int nx = -1, ny = -1, nz = -1;
  int ntypes = 4;
  int neighbor_size = -1;
  int halfneigh = 1;
  int use_sse = 0;
  int check_safeexchange = 0;
  int sort = -1;
  int yaml_output = 0;
  int screen_yaml = 0;
  std::string datafile;
  int units = -1;      // -1: from the input file
  int forcetype = -1;  // -1: from the input file
  int ghost_newton = 1;
  int device = -1;     // -d / --device; -1: LOCAL_RANK
  int stepwise = 0;    // --stepwise: per-call host loop instead of the fused mmd_run
  int quiet = 0;       // --quiet: no stdout (embedding)
  int help = 0;
  std::string eam_file;  // --eam_file (default Cu_u6.eam in cwd, as the reference)
  std::vector<std::string> unknown;
};

int parse_options(int argc, const char* const* argv, Options& opt);
void print_help();

class Simulation {
 public:
  Options opt;
  In in;
  World world;
  Atom* atom = nullptr;
  Neighbor* neighbor = nullptr;
  Integrate integrate;
  Thermo thermo;
  Comm comm;
  Timer timer;
  Force* force = nullptr;
  mmd_ctx* ctx = nullptr;
  std::string error;
  double device_ms_total = 0;  // CUDA-event time of all run() calls

  Simulation();
  ~Simulation();
  // everything up to and including the step-0 thermo record (ref/ljs.cpp:263-468); 0 = ok
  // host_only: stop after create_velocity without touching a GPU (planning / CPU tests of the host logic)
  int init(const Options& o, const World& w, const unsigned char* nccl_id128, bool host_only = false);
  // the timed region: Integrate::run for nsteps (<0: all remaining)
  int run(int nsteps = -1);
  // final force + thermo record + PERF_SUMMARY (ref/ljs.cpp:474-498)
  int finish();
  void print_header() const;

 private:
  bool out() const { return world.me == 0 && !opt.quiet; }
};
