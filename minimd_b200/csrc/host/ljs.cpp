// miniMD_b200: drop-in driver with the reference's command line, input file format and screen
// output (ref/ljs.cpp:61-504).  One process per GPU; ranks come from the launcher's environment
// (RANK / WORLD_SIZE / LOCAL_RANK, e.g. torchrun or mpirun), the NCCL id travels over a TCP
// rendezvous on MASTER_ADDR:MASTER_PORT+17.
#include <stdio.h>
#include <stdlib.h>

#include "sim.h"

int main(int argc, char** argv) {
  Options opt;
  parse_options(argc, argv, opt);
  if (opt.help) {
    print_help();
    return 0;
  }
  World world;
  world.from_env();
  for (const std::string& u : opt.unknown)
    if (world.me == 0) fprintf(stderr, "# warning: unknown option '%s' ignored\n", u.c_str());

  unsigned char id[128];
  std::string err;
  if (world.bootstrap_nccl_id(id, 17, &err)) {
    fprintf(stderr, "ERROR: %s\n", err.c_str());
    return 1;
  }

  Simulation sim;
  if (sim.init(opt, world, world.nprocs > 1 ? id : nullptr)) {
    if (world.me == 0) fprintf(stderr, "ERROR: %s\n", sim.error.c_str());
    return 1;
  }
  sim.run();
  sim.finish();
  sim.world.barrier();
  return 0;
}
