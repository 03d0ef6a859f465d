#include "comm.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void die(const char* what) {
  fprintf(stderr, "ERROR: %s: %s\n", what, mmd_last_error());
  exit(1);
}

Comm::Comm() {
  world = nullptr;
  me = 0;
  nswap = 0;
  check_safeexchange = 0;
  do_safeexchange = 0;
  memset(&table, 0, sizeof table);
  memset(procneigh, 0, sizeof procneigh);
  procgrid[0] = procgrid[1] = procgrid[2] = 1;
  myloc[0] = myloc[1] = myloc[2] = 0;
  need[0] = need[1] = need[2] = 1;
  memset(sendnum, 0, sizeof sendnum);
  memset(recvnum, 0, sizeof recvnum);
  memset(firstrecv, 0, sizeof firstrecv);
}
Comm::~Comm() {}

// rank of grid point (a,b,c) in a periodic Cartesian grid, last dimension fastest
// (the MPI_Cart_create / MPI_Cart_shift convention the reference relies on, ref/comm.cpp:133-137)
static int cart_rank(const int grid[3], int a, int b, int c) {
  a = ((a % grid[0]) + grid[0]) % grid[0];
  b = ((b % grid[1]) + grid[1]) % grid[1];
  c = ((c % grid[2]) + grid[2]) % grid[2];
  return (a * grid[1] + b) * grid[2] + c;
}

int Comm::setup(MMD_float cutneigh, Atom& atom) {
  const int nprocs = world ? world->nprocs : 1;
  me = world ? world->me : 0;
  MMD_float prd[3] = {atom.box.xprd, atom.box.yprd, atom.box.zprd};

  // factorisation of nprocs with the smallest sub-domain surface (ref/comm.cpp:80-120)
  MMD_float area[3] = {prd[0] * prd[1], prd[0] * prd[2], prd[1] * prd[2]};
  MMD_float bestsurf = 2.0 * (area[0] + area[1] + area[2]);
  procgrid[0] = procgrid[1] = procgrid[2] = 0;
  for (int ipx = 1; ipx <= nprocs; ipx++) {
    if (nprocs % ipx) continue;
    const int nremain = nprocs / ipx;
    for (int ipy = 1; ipy <= nremain; ipy++) {
      if (nremain % ipy) continue;
      const int ipz = nremain / ipy;
      MMD_float surf = area[0] / ipx / ipy + area[1] / ipx / ipz + area[2] / ipy / ipz;
      if (surf < bestsurf) {
        bestsurf = surf;
        procgrid[0] = ipx;
        procgrid[1] = ipy;
        procgrid[2] = ipz;
      }
    }
  }
  if (procgrid[0] * procgrid[1] * procgrid[2] != nprocs) {
    if (me == 0) printf("ERROR: Bad grid of processors\n");
    return 1;
  }

  myloc[2] = me % procgrid[2];
  myloc[1] = (me / procgrid[2]) % procgrid[1];
  myloc[0] = me / (procgrid[2] * procgrid[1]);
  for (int d = 0; d < 3; d++) {
    int lo[3] = {myloc[0], myloc[1], myloc[2]}, hi[3] = {myloc[0], myloc[1], myloc[2]};
    lo[d] -= 1;
    hi[d] += 1;
    procneigh[d][0] = cart_rank(procgrid, lo[0], lo[1], lo[2]);
    procneigh[d][1] = cart_rank(procgrid, hi[0], hi[1], hi[2]);
  }

  // my sub-box
  atom.box.xlo = myloc[0] * prd[0] / procgrid[0];
  atom.box.xhi = (myloc[0] + 1) * prd[0] / procgrid[0];
  atom.box.ylo = myloc[1] * prd[1] / procgrid[1];
  atom.box.yhi = (myloc[1] + 1) * prd[1] / procgrid[1];
  atom.box.zlo = myloc[2] * prd[2] / procgrid[2];
  atom.box.zhi = (myloc[2] + 1) * prd[2] / procgrid[2];
  const MMD_float sublo[3] = {atom.box.xlo, atom.box.ylo, atom.box.zlo};
  const MMD_float subhi[3] = {atom.box.xhi, atom.box.yhi, atom.box.zhi};

  for (int d = 0; d < 3; d++) need[d] = static_cast<int>(cutneigh * procgrid[d] / prd[d] + 1);
  if (2 * (need[0] + need[1] + need[2]) > MMD_MAX_SWAPS) {
    if (me == 0) printf("ERROR: box too small for the neighbor cutoff (more than %d swaps needed)\n", MMD_MAX_SWAPS);
    return 1;
  }

  // swap table (ref/comm.cpp:208-269): per dimension `need` layers of two swaps; the even swap
  // sends my low slab down, the odd one my high slab up; a swap that crosses the periodic
  // boundary shifts positions by +-prd when packing
  memset(&table, 0, sizeof table);
  table.me = me;
  table.nprocs = nprocs;
  nswap = 0;
  for (int idim = 0; idim < 3; idim++) {
    table.need[idim] = need[idim];
    table.procgrid[idim] = procgrid[idim];
    table.procneigh[idim][0] = procneigh[idim][0];
    table.procneigh[idim][1] = procneigh[idim][1];
    for (int ineed = 0; ineed < 2 * need[idim]; ineed++) {
      int* flag = idim == 0 ? table.pbc_flagx : (idim == 1 ? table.pbc_flagy : table.pbc_flagz);
      MMD_float lo, hi;
      if (ineed % 2 == 0) {
        table.sendproc[nswap] = procneigh[idim][0];
        table.recvproc[nswap] = procneigh[idim][1];
        const int nbox = myloc[idim] + ineed / 2;
        lo = nbox * prd[idim] / procgrid[idim];
        hi = sublo[idim] + cutneigh;
        const MMD_float cap = (nbox + 1) * prd[idim] / procgrid[idim];
        if (cap < hi) hi = cap;
        if (myloc[idim] == 0) {
          table.pbc_any[nswap] = 1;
          flag[nswap] = 1;
        }
      } else {
        table.sendproc[nswap] = procneigh[idim][1];
        table.recvproc[nswap] = procneigh[idim][0];
        const int nbox = myloc[idim] - ineed / 2;
        hi = (nbox + 1) * prd[idim] / procgrid[idim];
        lo = subhi[idim] - cutneigh;
        const MMD_float floor_ = nbox * prd[idim] / procgrid[idim];
        if (floor_ > lo) lo = floor_;
        if (myloc[idim] == procgrid[idim] - 1) {
          table.pbc_any[nswap] = 1;
          flag[nswap] = -1;
        }
      }
      table.slablo[nswap] = (double)lo;
      table.slabhi[nswap] = (double)hi;
      nswap++;
    }
  }
  table.nswap = nswap;

  if (atom.ctx) {
    const double p[3] = {(double)prd[0], (double)prd[1], (double)prd[2]};
    const double l[3] = {(double)sublo[0], (double)sublo[1], (double)sublo[2]};
    const double h[3] = {(double)subhi[0], (double)subhi[1], (double)subhi[2]};
    if (mmd_atom_set_box(atom.ctx, p, l, h) || mmd_comm_setup(atom.ctx, &table)) {
      fprintf(stderr, "ERROR: mmd_comm_setup: %s\n", mmd_last_error());
      return 1;
    }
  }
  return 0;
}

void Comm::communicate(Atom& atom) {
  if (mmd_comm_communicate(atom.ctx)) die("mmd_comm_communicate");
}

void Comm::reverse_communicate(Atom& atom) {
  if (mmd_comm_reverse_communicate(atom.ctx)) die("mmd_comm_reverse_communicate");
}

void Comm::exchange(Atom& atom) {
  if (mmd_comm_exchange(atom.ctx)) die("mmd_comm_exchange");
  atom.refresh_counts();
}

void Comm::borders(Atom& atom) {
  if (mmd_comm_borders(atom.ctx)) die("mmd_comm_borders");
  atom.refresh_counts();
  refresh_counts(atom);
}

void Comm::refresh_counts(Atom& atom) {
  if (mmd_comm_swap_counts(atom.ctx, sendnum, recvnum, firstrecv)) die("mmd_comm_swap_counts");
}
