#include "sim.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "force_eam.h"
#include "force_lj.h"
#include "output.h"
#include "setup.h"

int parse_options(int argc, const char* const* argv, Options& o) {
  auto is = [&](int i, const char* a, const char* b = nullptr) { return !strcmp(argv[i], a) || (b && !strcmp(argv[i], b)); };
  for (int i = 1; i < argc; i++) {
    auto next_i = [&]() { return (i + 1 < argc) ? atoi(argv[++i]) : 0; };
    auto next_s = [&]() { return std::string((i + 1 < argc) ? argv[++i] : ""); };
    if (is(i, "-i", "--input_file")) o.input_file = next_s();
    else if (is(i, "-t", "--num_threads")) o.num_threads = next_i();
    else if (is(i, "--teams")) o.teams = next_i();
    else if (is(i, "-n", "--nsteps")) o.num_steps = next_i();
    else if (is(i, "-s", "--size")) o.system_size = next_i();
    else if (is(i, "-nx")) o.nx = next_i();
    else if (is(i, "-ny")) o.ny = next_i();
    else if (is(i, "-nz")) o.nz = next_i();
    else if (is(i, "--ntypes")) o.ntypes = next_i();
    else if (is(i, "-b", "--neigh_bins")) o.neighbor_size = next_i();
    else if (is(i, "--half_neigh")) o.halfneigh = next_i();
    else if (is(i, "-sse")) o.use_sse = next_i();
    else if (is(i, "--check_exchange")) o.check_safeexchange = 1;
    else if (is(i, "--sort")) o.sort = next_i();
    else if (is(i, "-o", "--yaml_output")) o.yaml_output = next_i();
    else if (is(i, "--yaml_screen")) o.screen_yaml = 1;
    else if (is(i, "-f", "--data_file")) o.datafile = next_s();
    else if (is(i, "-u", "--units")) o.units = next_s() == "metal" ? METAL : LJ;
    else if (is(i, "-p", "--force")) o.forcetype = next_s() == "eam" ? FORCEEAM : FORCELJ;
    else if (is(i, "-gn", "--ghost_newton")) o.ghost_newton = next_i();
    else if (is(i, "-d", "--device")) o.device = next_i();
    else if (is(i, "-ng", "--num_gpus") || is(i, "--skip_gpu")) next_i();  // accepted, unused (rank->GPU comes from LOCAL_RANK)
    else if (is(i, "-dm", "--device_map") || is(i, "--safe_exchange")) {}   // accepted, unused
    else if (is(i, "--stepwise")) o.stepwise = 1;
    else if (is(i, "--quiet")) o.quiet = 1;
    else if (is(i, "--eam_file")) o.eam_file = next_s();
    else if (is(i, "-h", "--help")) o.help = 1;
    else o.unknown.push_back(argv[i]);
  }
  return 0;
}

void print_help() {
  printf("\n-----------------------------------------------------------------------------------------------------------\n");
  printf("-------------" VARIANT_STRING "--------------------\n");
  printf("-------------------------------------------------------------------------------------------------------------\n\n");
  printf("miniMD hot path on NVIDIA B200: same input files, options and output as the Mantevo miniMD\n"
         "reference variant; force, neighbor-list and integration kernels run on the GPU.\n\n");
  printf("Commandline Options:\n");
  printf("\n  Execution configuration:\n");
  printf("\t--teams <nteams>:             accepted for compatibility (no effect)\n");
  printf("\t-t / --num_threads <threads>: accepted for compatibility; echoed in the report (the GPU host loop is single-threaded)\n");
  printf("\t--half_neigh <int>:           use half neighborlists (default 1)\n"
         "\t                                0: full neighborlist\n"
         "\t                                1: half neighborlist\n"
         "\t                               -1: treated as 1\n");
  printf("\t-d / --device <int>:          CUDA device to use (default: LOCAL_RANK, one rank per GPU)\n");
  printf("\t-dm / -ng <int> / --skip_gpu <int>: accepted for compatibility (no effect)\n");
  printf("\t-gn / --ghost_newton <int>:   set usage of newtons third law for ghost atoms\n"
         "\t                                (only applicable with half neighborlists)\n");
  printf("\t--stepwise:                   run the time loop call by call through the host classes instead of the fused device loop\n");
  printf("\n  Simulation setup:\n");
  printf("\t-i / --input_file <string>:   set input file to be used (default: in.lj.miniMD)\n");
  printf("\t--ntypes <int>:               set number of atom types for simulation (default: 4)\n");
  printf("\t-n / --nsteps <int>:          set number of timesteps for simulation\n");
  printf("\t-s / --size <int>:            set linear dimension of systembox\n");
  printf("\t-nx/-ny/-nz <int>:            set linear dimension of systembox in x/y/z direction\n");
  printf("\t-b / --neigh_bins <int>:      set linear dimension of neighbor bin grid\n");
  printf("\t-u / --units <string>:        set units (lj or metal), see LAMMPS documentation\n");
  printf("\t-p / --force <string>:        set interaction model (lj or eam)\n");
  printf("\t--eam_file <string>:          EAM funcfl table (default: Cu_u6.eam in the working directory)\n");
  printf("\n  Miscelaneous:\n");
  printf("\t--sort <n>:                   resort atoms (simple bins) every <n> steps (default: use reneigh frequency; never=0)\n");
  printf("\t-o / --yaml_output <int>:     level of yaml output (default 0)\n");
  printf("\t--yaml_screen:                write yaml output also to screen\n");
  printf("\t--quiet:                      no screen output\n");
  printf("\t-h / --help:                  display this help message\n\n");
  printf("---------------------------------------------------------\n\n");
}

Simulation::Simulation() {}

Simulation::~Simulation() {
  delete force;
  delete neighbor;
  delete atom;
  if (ctx) mmd_ctx_destroy(ctx);
}

int Simulation::init(const Options& o, const World& w, const unsigned char* nccl_id128, bool host_only) {
  opt = o;
  world = w;
  const int me = world.me;
  const char* infile = opt.input_file.empty() ? "in.lj.miniMD" : opt.input_file.c_str();
  if (input(in, infile, opt.quiet ? 1 : me)) {
    error = std::string("cannot read input file ") + infile;
    return 1;
  }
  srand(5413);  // ref/ljs.cpp:110: the atom-type stream
  if (!opt.datafile.empty()) in.datafile = opt.datafile;
  if (opt.units >= 0) in.units = opt.units;
  if (opt.forcetype >= 0) in.forcetype = (ForceStyle)opt.forcetype;

  // device context of this rank
  const int device = opt.device >= 0 ? opt.device : world.device;
  if (!host_only && mmd_ctx_create(device, (int)sizeof(MMD_float), opt.ntypes, nullptr, &ctx)) {
    error = mmd_last_error();
    if (out()) printf("ERROR: %s\n", error.c_str());
    return 1;
  }
  world.ctx = ctx;
  if (!host_only && world.nprocs > 1) {
    if (!nccl_id128 || mmd_comm_nccl_init(ctx, nccl_id128, world.me, world.nprocs)) {
      error = nccl_id128 ? mmd_last_error() : "multi-rank run without an NCCL id";
      if (out()) printf("ERROR: %s\n", error.c_str());
      return 1;
    }
  }

  atom = new Atom(opt.ntypes);
  atom->ctx = ctx;
  neighbor = new Neighbor(opt.ntypes);
  thermo.world = &world;
  thermo.quiet = opt.quiet != 0;
  comm.world = &world;

  int ghost_newton = opt.ghost_newton;
  if (in.forcetype == FORCEEAM) {
    ForceEAM* eam = new ForceEAM(opt.ntypes);
    if (!opt.eam_file.empty()) eam->potential_file = opt.eam_file;
    force = eam;
    if (ghost_newton == 1) {
      if (out()) printf("# EAM currently requires '--ghost_newton 0'; Changing setting now.\n");
      ghost_newton = 0;
    }
  } else {
    force = new ForceLJ(opt.ntypes);
    for (int i = 0; i < opt.ntypes * opt.ntypes; i++) {
      force->epsilon[i] = in.epsilon;
      force->sigma[i] = in.sigma;
      force->sigma6[i] = in.sigma * in.sigma * in.sigma * in.sigma * in.sigma * in.sigma;
    }
  }
  opt.ghost_newton = ghost_newton;
  neighbor->ghost_newton = ghost_newton;
  comm.check_safeexchange = opt.check_safeexchange;
  comm.do_safeexchange = 0;
  force->use_sse = 0;
  neighbor->halfneigh = opt.halfneigh;
  if (opt.halfneigh < 0) force->use_oldcompute = 1;

  // size overrides (ref/ljs.cpp:330-349)
  if (opt.num_steps > 0) in.ntimes = opt.num_steps;
  if (opt.system_size > 0) in.nx = in.ny = in.nz = opt.system_size;
  if (opt.nx > 0) {
    in.nx = opt.nx;
    if (opt.ny > 0) in.ny = opt.ny;
    else if (opt.system_size < 0) in.ny = opt.nx;
    if (opt.nz > 0) in.nz = opt.nz;
    else if (opt.system_size < 0) in.nz = opt.nx;
  }
  // 5 bins per 6 lattice cells unless -b (ref/ljs.cpp:351-368); with a data file the bin count comes from the density
  if (opt.neighbor_size > 0) {
    neighbor->nbinx = neighbor->nbiny = neighbor->nbinz = opt.neighbor_size;
  } else if (!in.datafile.empty()) {
    neighbor->nbinx = neighbor->nbiny = neighbor->nbinz = -1;
  } else {
    MMD_float neighscale = 5.0 / 6.0;
    neighbor->nbinx = neighscale * in.nx;
    neighbor->nbiny = neighscale * in.ny;
    neighbor->nbinz = neighscale * in.nz;
  }
  if (neighbor->nbinx == 0) neighbor->nbinx = 1;
  if (neighbor->nbiny == 0) neighbor->nbiny = 1;
  if (neighbor->nbinz == 0) neighbor->nbinz = 1;
  // (with a data file and no -b, nbinx is -1 here and nbiny/nbinz follow in read_lammps_data)

  integrate.ntimes = in.ntimes;
  integrate.dt = in.dt;
  integrate.sort_every = opt.sort > 0 ? opt.sort : (opt.sort < 0 ? in.neigh_every : 0);
  integrate.stepwise = opt.stepwise;
  neighbor->every = in.neigh_every;
  neighbor->cutneigh = in.neigh_cut;
  force->cutforce = in.force_cut;
  thermo.nstat = in.thermo_nstat;

  if (out()) printf("# Create System:\n");
  if (!in.datafile.empty()) {  // ref/ljs.cpp:382-390
    if (read_lammps_data(*atom, comm, *neighbor, integrate, thermo, in.datafile.c_str(), in.units, world)) {
      error = "read_lammps_data failed for " + in.datafile;
      return 1;
    }
    const MMD_float volume = atom->box.xprd * atom->box.yprd * atom->box.zprd;
    in.rho = 1.0 * atom->natoms / volume;
    if (force->setup(*atom)) { error = "Force::setup failed"; return 1; }
    if (in.forcetype == FORCEEAM) atom->mass = force->mass;
  } else {
    create_box(*atom, in.nx, in.ny, in.nz, in.rho);
    if (comm.setup(neighbor->cutneigh, *atom)) { error = "Comm::setup failed"; return 1; }
    if (neighbor->setup(*atom)) { error = "Neighbor::setup failed"; return 1; }
    integrate.setup();
    if (force->setup(*atom)) { error = "Force::setup failed"; return 1; }
    if (in.forcetype == FORCEEAM) atom->mass = force->mass;
    if (create_atoms(*atom, in.nx, in.ny, in.nz, in.rho, world)) { error = "create_atoms failed"; return 1; }
    thermo.setup(in.rho, integrate, *atom, in.units);
    create_velocity(in.t_request, *atom, thermo, world);
  }
  if (host_only) return 0;
  if (atom->upload()) { error = mmd_last_error(); return 1; }
  if (out()) printf("# Done .... \n");
  if (out()) print_header();

  comm.exchange(*atom);
  if (opt.sort > 0) atom->sort(*neighbor);
  comm.borders(*atom);
  force->evflag = 1;
  neighbor->build(*atom);
  force->compute(*atom, *neighbor, comm, me);
  if (neighbor->halfneigh && neighbor->ghost_newton) comm.reverse_communicate(*atom);

  if (out()) printf("# Starting dynamics ...\n");
  if (out()) printf("# Timestep T U P Time\n");
  timer.clear();
  thermo.compute(0, *atom, *neighbor, force, timer, comm);
  mmd_set_option(ctx, "phase_timing", 1);
  return 0;
}

void Simulation::print_header() const {
  fprintf(stdout, "# " VARIANT_STRING " output ...\n");
  fprintf(stdout, "# Run Settings: \n");
  fprintf(stdout, "\t# MPI processes: %i\n", world.nprocs);
  fprintf(stdout, "\t# OpenMP threads: %i\n", opt.num_threads);
  fprintf(stdout, "\t# Inputfile: %s\n", opt.input_file.empty() ? "in.lj.miniMD" : opt.input_file.c_str());
  fprintf(stdout, "\t# Datafile: %s\n", in.datafile.empty() ? "None" : in.datafile.c_str());
  fprintf(stdout, "# Physics Settings: \n");
  fprintf(stdout, "\t# ForceStyle: %s\n", in.forcetype == FORCELJ ? "LJ" : "EAM");
  fprintf(stdout, "\t# Force Parameters: %2.2lf %2.2lf\n", (double)in.epsilon, (double)in.sigma);
  fprintf(stdout, "\t# Units: %s\n", in.units == 0 ? "LJ" : "METAL");
  fprintf(stdout, "\t# Atoms: %i\n", atom->natoms);
  fprintf(stdout, "\t# Atom types: %i\n", atom->ntypes);
  fprintf(stdout, "\t# System size: %2.2lf %2.2lf %2.2lf (unit cells: %i %i %i)\n", (double)atom->box.xprd,
          (double)atom->box.yprd, (double)atom->box.zprd, in.nx, in.ny, in.nz);
  fprintf(stdout, "\t# Density: %lf\n", (double)in.rho);
  fprintf(stdout, "\t# Force cutoff: %lf\n", (double)force->cutforce);
  fprintf(stdout, "\t# Timestep size: %lf\n", (double)integrate.dt);
  fprintf(stdout, "# Technical Settings: \n");
  fprintf(stdout, "\t# Neigh cutoff: %lf\n", (double)neighbor->cutneigh);
  fprintf(stdout, "\t# Half neighborlists: %i\n", neighbor->halfneigh);
  fprintf(stdout, "\t# Neighbor bins: %i %i %i\n", neighbor->nbinx, neighbor->nbiny, neighbor->nbinz);
  fprintf(stdout, "\t# Neighbor frequency: %i\n", neighbor->every);
  fprintf(stdout, "\t# Sorting frequency: %i\n", integrate.sort_every);
  fprintf(stdout, "\t# Thermo frequency: %i\n", thermo.nstat);
  fprintf(stdout, "\t# Ghost Newton: %i\n", opt.ghost_newton);
  fprintf(stdout, "\t# Use intrinsics: %i\n", force->use_sse);
  fprintf(stdout, "\t# Do safe exchange: %i\n", comm.do_safeexchange);
  fprintf(stdout, "\t# Size of float: %i\n\n", (int)sizeof(MMD_float));
}

int Simulation::run(int nsteps) {
  timer.barrier_start(TIME_TOTAL);
  integrate.run(*atom, force, *neighbor, comm, thermo, timer, nsteps);
  timer.barrier_stop(TIME_TOTAL);
  device_ms_total += integrate.device_ms;
  double ms[MMD_NPHASE];
  long long calls[MMD_NPHASE];
  if (!mmd_run_phase_times(ctx, ms, calls, 0)) {
    timer.array[TIME_COMM] = 1e-3 * ms[MMD_PHASE_COMM];
    timer.array[TIME_FORCE] = 1e-3 * ms[MMD_PHASE_FORCE];
    timer.array[TIME_NEIGH] = 1e-3 * ms[MMD_PHASE_NEIGH];
  }
  return 0;
}

int Simulation::finish() {
  const long long natoms = world.sum_ll(atom->nlocal);
  force->evflag = 1;
  force->compute(*atom, *neighbor, comm, world.me);
  if (neighbor->halfneigh && neighbor->ghost_newton) comm.reverse_communicate(*atom);
  thermo.compute(-1, *atom, *neighbor, force, timer, comm);

  if (out()) {
    const double t_total = timer.array[TIME_TOTAL];
    const double time_other = t_total - timer.array[TIME_FORCE] - timer.array[TIME_NEIGH] - timer.array[TIME_COMM];
    const int nsteps = integrate.steps_done;
    printf("\n\n");
    printf("# Performance Summary:\n");
    printf("# MPI_proc OMP_threads nsteps natoms t_total t_force t_neigh t_comm t_other performance perf/thread grep_string t_extra\n");
    printf("%i %i %i %i %lf %lf %lf %lf %lf %lf %lf PERF_SUMMARY %lf\n\n\n", world.nprocs, opt.num_threads, nsteps, (int)natoms,
           t_total, timer.array[TIME_FORCE], timer.array[TIME_NEIGH], timer.array[TIME_COMM], time_other,
           1.0 * natoms * nsteps / t_total, 1.0 * natoms * nsteps / t_total / world.nprocs / opt.num_threads,
           timer.array[TIME_TEST]);
    printf("# device time of the time loop (CUDA events): %lf s => %lf atom-steps/s\n", 1e-3 * device_ms_total,
           1.0 * natoms * nsteps / (1e-3 * device_ms_total));
  }
  if (opt.yaml_output) output(*this);
  return 0;
}
