// The process group of a run: one rank per GPU.  Replaces the reference's use of MPI_COMM_WORLD
// (rank/size queries, MPI_Allreduce of a few scalars, MPI_Barrier); all data-path messages are
// NCCL point-to-point inside the device library (mmd_comm_*).
#pragma once
#include <string>

#include "minimd_b200.h"

struct World {
  int me = 0;
  int nprocs = 1;
  int device = 0;
  mmd_ctx* ctx = nullptr;  // set once the context exists; collectives need it when nprocs > 1
  // optional embedder-supplied reduction (op 0 sum, 1 max) used instead of NCCL when set: lets the
  // host-only planning path (mmd_sim_plan) run multi-rank on a box without GPUs (e.g. over gloo)
  void (*reduce_cb)(double* values, int n, int op, void* user) = nullptr;
  void* reduce_user = nullptr;

  // Rank layout from the launcher's environment (torchrun: RANK / WORLD_SIZE / LOCAL_RANK).
  void from_env();
  // Fetch the 128-byte NCCL id: rank 0 creates it and serves it over TCP on MASTER_ADDR:(MASTER_PORT+port_offset);
  // used by the stand-alone driver.  Embedders (bench.py) pass the id in instead.  0 = ok.
  int bootstrap_nccl_id(unsigned char id[128], int port_offset, std::string* err);

  // MPI_Allreduce(MPI_SUM / MPI_MAX) of a few doubles; identity on one rank.
  void sum(double* v, int n) const;
  void max(double* v, int n) const;
  long long sum_ll(long long v) const;
  void barrier() const;
};
