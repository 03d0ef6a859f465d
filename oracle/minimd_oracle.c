/* =====================================================================================
 * minimd_oracle.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * A plain-C, single-rank, single-thread CPU restatement of the miniMD `ref/` hot path:
 * pair force (LJ half/full, EAM half/full), bin-and-stencil neighbor rebuild, atom sort,
 * single-rank ghost self-swaps, velocity-Verlet, and the thermo reductions, plus the
 * deterministic FCC/Park-Miller setup that defines the synthetic inputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this.  It is the CHECKER for the CUDA path in minimd_b200/csrc, nothing else.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file's T/U/P (a) at every
 * step against a live run of the unmodified reference binary (oracle/_ref/, built by
 * oracle/build_ref.sh from /root/reference), (b) against 10-digit outputs of that binary on
 * BASELINE.json's configurations (tests/golden/reference_runs.json, made by
 * tests/golden/make_golden.py) and (c) against the reference's shipped logs
 * tests/reference_output/{4k,16k,32k}.{lj,eam} (committed as tests/golden/reference_logs.json).
 *
 * Arithmetic is kept in the reference's evaluation order (serial, no FMA contraction: build
 * with -ffp-contract=off, matching the reference's g++ -O3 -mavx build) so that single-thread
 * results are bit-identical to the reference binary.  `real` is the reference's MMD_float
 * (ref/types.h:61-74), chosen at compile time with -DORC_PRECISION=1|2 exactly like the
 * reference's -DPRECISION.  Arrays are AoS with stride PAD=3 (ref/types.h:77-81).
 *
 * All `ref/...` citations are relative to /root/reference.
 * ===================================================================================== */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef ORC_PRECISION
#define ORC_PRECISION 2
#endif
#if ORC_PRECISION == 1
typedef float real;
#define REAL_SQRT(v) sqrtf(v) /* C++ overload resolution picks sqrt(float) in the reference's FP32 build */
#else
#define REAL_SQRT(v) sqrt(v)
typedef double real;
#endif
#define PAD 3

#define ORC_MAX_SWAP 32

typedef struct {
  real xprd, yprd, zprd;
  real xlo, xhi, ylo, yhi, zlo, zhi;
} orc_box;

typedef struct orc_sim {
  /* ---- Atom (ref/atom.h:47-106) ---- */
  int natoms, nlocal, nghost, nmax, ntypes;
  real *x, *v, *f;
  int *type;
  real *x_alt, *v_alt; /* sort double buffers (ref/atom.cpp:355-421) */
  int *type_alt;
  real mass;
  orc_box box;
  /* ---- Neighbor (ref/neighbor.h) ---- */
  int every, nbinx, nbiny, nbinz;
  real cutneigh;
  real *cutneighsq;
  int *numneigh, *neighbors;
  int maxneighs, neigh_rows;
  int halfneigh, ghost_newton;
  int *bincount, *bins;
  int mbins, atoms_per_bin;
  int nstencil, *stencil;
  int mbinx, mbiny, mbinz, mbinxlo, mbinylo, mbinzlo;
  real binsizex, binsizey, binsizez, bininvx, bininvy, bininvz;
  int ncalls;
  /* ---- Comm, single rank (ref/comm.cpp:60-272) ---- */
  int nswap, need[3];
  real slablo[ORC_MAX_SWAP], slabhi[ORC_MAX_SWAP];
  int pbc_any[ORC_MAX_SWAP], pbc_flagx[ORC_MAX_SWAP], pbc_flagy[ORC_MAX_SWAP], pbc_flagz[ORC_MAX_SWAP];
  int sendnum[ORC_MAX_SWAP], recvnum[ORC_MAX_SWAP], firstrecv[ORC_MAX_SWAP];
  int *sendlist[ORC_MAX_SWAP];
  int maxsendlist[ORC_MAX_SWAP];
  /* ---- Force ---- */
  int forcetype; /* 0 LJ, 1 EAM */
  real cutforce;
  real *cutforcesq, *epsilon, *sigma6;
  real eng_vdwl, virial;
  int evflag;
  /* EAM (ref/force_eam.h) */
  int nr, nrho, nr_tot, nrho_tot;
  real dr, rdr, drho, rdrho;
  real *rhor_spline, *z2r_spline, *frho_spline;
  real *rho, *fp;
  int eam_nmax;
  real eam_mass, eam_cut;
  /* ---- Integrate / Thermo ---- */
  real dt, dtforce;
  int ntimes, sort_every, nstat, units;
  real rho_in;
  real t_scale, e_scale, p_scale, mvv2e, dof_boltz;
  /* thermo log */
  int nlog, log_cap;
  int *log_step;
  double *log_t, *log_e, *log_p;
} orc_sim;

/* ------------------------------------------------------------------------------------- */
/* storage                                                                               */
/* ------------------------------------------------------------------------------------- */

static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}

/* Atom::growarray (ref/atom.cpp:71-84): capacity grows in steps of 20000 atoms. */
static void atoms_reserve(orc_sim *s, int n) {
  if (n <= s->nmax) return;
  int cap = s->nmax;
  while (cap < n) cap += 20000;
  s->x = (real *)realloc(s->x, sizeof(real) * PAD * (size_t)cap);
  s->v = (real *)realloc(s->v, sizeof(real) * PAD * (size_t)cap);
  s->f = (real *)realloc(s->f, sizeof(real) * PAD * (size_t)cap);
  s->type = (int *)realloc(s->type, sizeof(int) * (size_t)cap);
  s->x_alt = (real *)realloc(s->x_alt, sizeof(real) * PAD * (size_t)cap);
  s->v_alt = (real *)realloc(s->v_alt, sizeof(real) * PAD * (size_t)cap);
  s->type_alt = (int *)realloc(s->type_alt, sizeof(int) * (size_t)cap);
  s->nmax = cap;
}

orc_sim *orc_create(int ntypes) {
  orc_sim *s = (orc_sim *)calloc(1, sizeof(orc_sim));
  s->ntypes = ntypes;
  s->mass = 1;                 /* ref/atom.cpp:55 */
  s->maxneighs = 100;          /* ref/neighbor.cpp:48 */
  s->atoms_per_bin = 8;        /* ref/neighbor.cpp:52 */
  s->ghost_newton = 1;         /* ref/neighbor.cpp:56 */
  s->halfneigh = 1;
  s->every = 20;
  s->sort_every = 20;
  int nn = ntypes * ntypes;
  s->cutneighsq = (real *)xmalloc(sizeof(real) * nn);
  s->cutforcesq = (real *)xmalloc(sizeof(real) * nn);
  s->epsilon = (real *)xmalloc(sizeof(real) * nn);
  s->sigma6 = (real *)xmalloc(sizeof(real) * nn);
  for (int i = 0; i < nn; i++) { s->cutforcesq[i] = 0; s->epsilon[i] = 1; s->sigma6[i] = 1; s->cutneighsq[i] = 0; }
  for (int i = 0; i < ORC_MAX_SWAP; i++) { s->maxsendlist[i] = 0; s->sendlist[i] = NULL; }
  return s;
}

void orc_destroy(orc_sim *s) {
  if (!s) return;
  free(s->x); free(s->v); free(s->f); free(s->type); free(s->x_alt); free(s->v_alt); free(s->type_alt);
  free(s->cutneighsq); free(s->cutforcesq); free(s->epsilon); free(s->sigma6);
  free(s->numneigh); free(s->neighbors); free(s->bincount); free(s->bins); free(s->stencil);
  for (int i = 0; i < ORC_MAX_SWAP; i++) free(s->sendlist[i]);
  free(s->rhor_spline); free(s->z2r_spline); free(s->frho_spline); free(s->rho); free(s->fp);
  free(s->log_step); free(s->log_t); free(s->log_e); free(s->log_p);
  free(s);
}

/* ------------------------------------------------------------------------------------- */
/* setup: box, lattice, velocities  (ref/setup.cpp:305-517)                              */
/* ------------------------------------------------------------------------------------- */

/* create_box (ref/setup.cpp:305-311) + the single-rank part of Comm::setup that sets the
   sub-box bounds (ref/comm.cpp:139-146 with procgrid = 1x1x1, myloc = 0). */
void orc_create_box(orc_sim *s, int nx, int ny, int nz, double rho) {
  double lattice = pow((4.0 / rho), (1.0 / 3.0));
  s->box.xprd = nx * lattice;
  s->box.yprd = ny * lattice;
  s->box.zprd = nz * lattice;
  real prd[3] = {s->box.xprd, s->box.yprd, s->box.zprd};
  s->box.xlo = 0 * prd[0] / 1;  s->box.xhi = (0 + 1) * prd[0] / 1;
  s->box.ylo = 0 * prd[1] / 1;  s->box.yhi = (0 + 1) * prd[1] / 1;
  s->box.zlo = 0 * prd[2] / 1;  s->box.zhi = (0 + 1) * prd[2] / 1;
  s->rho_in = rho;
}

/* Park-Miller minimal standard generator, no masking (ref/setup.cpp:498-517). */
static double park_miller(int *state) {
  const int a = 16807, m = 2147483647, q = 127773, r = 2836;
  int hi = *state / q;
  *state = a * (*state - hi * q) - r * hi;
  if (*state < 0) *state += m;
  return (1.0 / m) * (*state);
}

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* create_atoms (ref/setup.cpp:315-450).  The reference walks the half-lattice points in
   8x8x8 sub-blocks (block index slowest: z, y, x; inside a block z, y, x with x fastest);
   that order defines the initial atom numbering.  Velocity of the atom at lattice point
   (i,j,k) comes from a Park-Miller stream seeded with its global lattice id (5 discards
   before each of the 3 draws).  type = rand()%ntypes (ref/atom.cpp:97, libc rand seeded with
   5413 in ref/ljs.cpp:110 -- the caller seeds). */
int orc_create_atoms(orc_sim *s, int nx, int ny, int nz, double rho) {
  s->natoms = 4 * nx * ny * nz;
  s->nlocal = 0;
  double alat = pow((4.0 / rho), (1.0 / 3.0));
  int ilo = (int)(s->box.xlo / (0.5 * alat) - 1), ihi = (int)(s->box.xhi / (0.5 * alat) + 1);
  int jlo = (int)(s->box.ylo / (0.5 * alat) - 1), jhi = (int)(s->box.yhi / (0.5 * alat) + 1);
  int klo = (int)(s->box.zlo / (0.5 * alat) - 1), khi = (int)(s->box.zhi / (0.5 * alat) + 1);
  ilo = imax(ilo, 0); ihi = imin(ihi, 2 * nx - 1);
  jlo = imax(jlo, 0); jhi = imin(jhi, 2 * ny - 1);
  klo = imax(klo, 0); khi = imin(khi, 2 * nz - 1);
  const int B = 8;
  for (int oz = 0; oz * B <= khi; oz++)
    for (int oy = 0; oy * B <= jhi; oy++)
      for (int ox = 0; ox * B <= ihi; ox++)
        for (int sz = 0; sz < B; sz++)
          for (int sy = 0; sy < B; sy++)
            for (int sx = 0; sx < B; sx++) {
              int i = ox * B + sx, j = oy * B + sy, k = oz * B + sz;
              if ((i + j + k) % 2 != 0) continue;
              if (i < ilo || i > ihi || j < jlo || j > jhi || k < klo || k > khi) continue;
              double xt = 0.5 * alat * i, yt = 0.5 * alat * j, zt = 0.5 * alat * k;
              if (!(xt >= s->box.xlo && xt < s->box.xhi && yt >= s->box.ylo && yt < s->box.yhi &&
                    zt >= s->box.zlo && zt < s->box.zhi)) continue;
              int seed = k * (2 * ny) * (2 * nx) + j * (2 * nx) + i + 1;
              double vel[3];
              for (int c = 0; c < 3; c++) {
                for (int d = 0; d < 5; d++) park_miller(&seed);
                vel[c] = park_miller(&seed);
              }
              atoms_reserve(s, s->nlocal + 1);
              int n = s->nlocal;
              s->x[n * PAD + 0] = xt; s->x[n * PAD + 1] = yt; s->x[n * PAD + 2] = zt;
              s->v[n * PAD + 0] = vel[0]; s->v[n * PAD + 1] = vel[1]; s->v[n * PAD + 2] = vel[2];
              s->type[n] = rand() % s->ntypes;
              s->nlocal++;
            }
  return s->nlocal == s->natoms ? 0 : 1;
}

/* Thermo::setup (ref/thermo.cpp:42-72).  NOTE: for METAL it also rescales
   integrate.dtforce by 1/mvv2e (ref/thermo.cpp:69) -- so orc_integrate_setup must run first,
   as in ref/ljs.cpp:397-405. */
void orc_thermo_setup(orc_sim *s, int units) {
  s->units = units;
  if (units == 0) {
    s->mvv2e = 1.0;
    s->dof_boltz = (s->natoms * 3 - 3);
    s->t_scale = s->mvv2e / s->dof_boltz;
    s->p_scale = 1.0 / 3 / s->box.xprd / s->box.yprd / s->box.zprd;
    s->e_scale = 0.5;
  } else {
    s->mvv2e = 1.036427e-04;
    s->dof_boltz = (s->natoms * 3 - 3) * 8.617343e-05;
    s->t_scale = s->mvv2e / s->dof_boltz;
    s->p_scale = 1.602176e+06 / 3 / s->box.xprd / s->box.yprd / s->box.zprd;
    s->e_scale = 524287.985533;
    s->dtforce /= s->mvv2e;
  }
}

/* Thermo::temperature (ref/thermo.cpp:140-174): sum_i (v.v)*mass, then * t_scale. */
real orc_temperature(orc_sim *s) {
  real t = 0.0;
  const real *v = s->v;
  for (int i = 0; i < s->nlocal; i++) {
    real vx = v[i * PAD + 0], vy = v[i * PAD + 1], vz = v[i * PAD + 2];
    t += (vx * vx + vy * vy + vz * vz) * s->mass;
  }
  return t * s->t_scale;
}

/* create_velocity (ref/setup.cpp:454-494): remove centre-of-mass drift, rescale to T. */
void orc_create_velocity(orc_sim *s, double t_request) {
  double tot[3] = {0.0, 0.0, 0.0};
  for (int i = 0; i < s->nlocal; i++)
    for (int c = 0; c < 3; c++) tot[c] += s->v[i * PAD + c];
  for (int c = 0; c < 3; c++) tot[c] = tot[c] / s->natoms;
  for (int i = 0; i < s->nlocal; i++)
    for (int c = 0; c < 3; c++) s->v[i * PAD + c] -= tot[c];
  double t = orc_temperature(s);
  double factor = sqrt(t_request / t);
  for (int i = 0; i < s->nlocal; i++)
    for (int c = 0; c < 3; c++) s->v[i * PAD + c] *= factor;
}

/* Integrate::setup (ref/integrate.cpp:41-44). */
void orc_integrate_setup(orc_sim *s, double dt) {
  s->dt = dt;
  s->dtforce = 0.5 * s->dt;
}

/* ------------------------------------------------------------------------------------- */
/* neighbor geometry (ref/neighbor.cpp:274-482)                                          */
/* ------------------------------------------------------------------------------------- */

/* Neighbor::coord2bin (ref/neighbor.cpp:274-300): truncating casts, reciprocal multiply,
   three branches per axis, and the "+1" on the linear index are all part of the contract. */
static int axis_bin(real c, real prd, real bininv, int nbin, int mbinlo) {
  if (c >= prd) return (int)((c - prd) * bininv) + nbin - mbinlo;
  if (c >= 0.0) return (int)(c * bininv) - mbinlo;
  return (int)(c * bininv) - mbinlo - 1;
}
int orc_coord2bin(const orc_sim *s, real x, real y, real z) {
  int ix = axis_bin(x, s->box.xprd, s->bininvx, s->nbinx, s->mbinxlo);
  int iy = axis_bin(y, s->box.yprd, s->bininvy, s->nbiny, s->mbinylo);
  int iz = axis_bin(z, s->box.zprd, s->bininvz, s->nbinz, s->mbinzlo);
  return (iz * s->mbiny * s->mbinx + iy * s->mbinx + ix + 1);
}

/* Neighbor::bindist (ref/neighbor.cpp:456-482): closest approach of bin (i,j,k) to bin 0. */
static real bin_gap_sq(const orc_sim *s, int i, int j, int k) {
  real dx = (i > 0) ? (i - 1) * s->binsizex : ((i == 0) ? 0.0 : (i + 1) * s->binsizex);
  real dy = (j > 0) ? (j - 1) * s->binsizey : ((j == 0) ? 0.0 : (j + 1) * s->binsizey);
  real dz = (k > 0) ? (k - 1) * s->binsizez : ((k == 0) ? 0.0 : (k + 1) * s->binsizez);
  return (dx * dx + dy * dy + dz * dz);
}

static void axis_extent(real lo, real hi, real cut, real prd, real bininv, int *mlo, int *mhi) {
  real c = lo - cut - 1.0e-6 * prd;
  *mlo = (int)(c * bininv);
  if (c < 0.0) *mlo = *mlo - 1;
  c = hi + cut + 1.0e-6 * prd;
  *mhi = (int)(c * bininv);
  *mlo -= 1; /* one extra layer for stencil coverage (ref/neighbor.cpp:383-393) */
  *mhi += 1;
}

/* Neighbor::setup (ref/neighbor.cpp:318-452). nbin{x,y,z}, cutneigh, halfneigh, ghost_newton
   must be set by the caller beforehand (ref/ljs.cpp:357-378). */
int orc_neighbor_setup(orc_sim *s) {
  for (int i = 0; i < s->ntypes * s->ntypes; i++) s->cutneighsq[i] = s->cutneigh * s->cutneigh;
  s->binsizex = s->box.xprd / s->nbinx;
  s->binsizey = s->box.yprd / s->nbiny;
  s->binsizez = s->box.zprd / s->nbinz;
  s->bininvx = 1.0 / s->binsizex;
  s->bininvy = 1.0 / s->binsizey;
  s->bininvz = 1.0 / s->binsizez;
  int hix, hiy, hiz;
  axis_extent(s->box.xlo, s->box.xhi, s->cutneigh, s->box.xprd, s->bininvx, &s->mbinxlo, &hix);
  axis_extent(s->box.ylo, s->box.yhi, s->cutneigh, s->box.yprd, s->bininvy, &s->mbinylo, &hiy);
  axis_extent(s->box.zlo, s->box.zhi, s->cutneigh, s->box.zprd, s->bininvz, &s->mbinzlo, &hiz);
  s->mbinx = hix - s->mbinxlo + 1;
  s->mbiny = hiy - s->mbinylo + 1;
  s->mbinz = hiz - s->mbinzlo + 1;

  int nextx = (int)(s->cutneigh * s->bininvx); if (nextx * s->binsizex < 0.999 * s->cutneigh) nextx++;
  int nexty = (int)(s->cutneigh * s->bininvy); if (nexty * s->binsizey < 0.999 * s->cutneigh) nexty++;
  int nextz = (int)(s->cutneigh * s->bininvz); if (nextz * s->binsizez < 0.999 * s->cutneigh) nextz++;

  free(s->stencil);
  s->stencil = (int *)xmalloc(sizeof(int) * (2 * nextz + 1) * (2 * nexty + 1) * (2 * nextx + 1));
  s->nstencil = 0;
  int newton_half = s->halfneigh && s->ghost_newton;
  int kstart = -nextz;
  if (newton_half) { kstart = 0; s->stencil[s->nstencil++] = 0; }
  for (int k = kstart; k <= nextz; k++)
    for (int j = -nexty; j <= nexty; j++)
      for (int i = -nextx; i <= nextx; i++) {
        if (newton_half && !(k > 0 || j > 0 || (j == 0 && i > 0))) continue;
        if (bin_gap_sq(s, i, j, k) < s->cutneighsq[0])
          s->stencil[s->nstencil++] = k * s->mbiny * s->mbinx + j * s->mbinx + i;
      }
  s->mbins = s->mbinx * s->mbiny * s->mbinz;
  free(s->bincount); free(s->bins);
  s->bincount = (int *)xmalloc(sizeof(int) * s->mbins);
  s->bins = (int *)xmalloc(sizeof(int) * (size_t)s->mbins * s->atoms_per_bin);
  return 0;
}

/* ------------------------------------------------------------------------------------- */
/* binning, neighbor build, sort                                                          */
/* ------------------------------------------------------------------------------------- */

/* Neighbor::binatoms (ref/neighbor.cpp:215-268): append atom ids to fixed-width bin rows in
   index order; on overflow double the row width and start over. count<0 => nlocal+nghost. */
void orc_binatoms(orc_sim *s, int count) {
  int nall = count < 0 ? s->nlocal + s->nghost : count;
  int again = 1;
  while (again) {
    again = 0;
    for (int b = 0; b < s->mbins; b++) s->bincount[b] = 0;
    for (int i = 0; i < nall; i++) {
      int b = orc_coord2bin(s, s->x[i * PAD + 0], s->x[i * PAD + 1], s->x[i * PAD + 2]);
      if (s->bincount[b] < s->atoms_per_bin) {
        int slot = s->bincount[b]++;
        s->bins[(size_t)b * s->atoms_per_bin + slot] = i;
      } else again = 1;
    }
    if (again) {
      free(s->bins);
      s->atoms_per_bin *= 2;
      s->bins = (int *)xmalloc(sizeof(int) * (size_t)s->mbins * s->atoms_per_bin);
    }
  }
}

/* Neighbor::build (ref/neighbor.cpp:79-213).  Accept test is rsq <= cutneighsq (note <=).
   Same-bin filter (:154-157): skip self; half lists skip j<i; half+ghost_newton also skips a
   ghost j that is lexicographically (z,y,x) below i, by exact FP comparison.  Other bins
   (:171): only half & !ghost_newton skips j<i.  Resize protocol (:186-208): if any row reaches
   maxneighs the whole build is repeated with maxneighs = max_n*1.2.  Unlike the reference this
   restatement never stores past a row's end (the reference overruns into the next rows before
   it notices; the stored result after the redo is identical). */
void orc_neighbor_build(orc_sim *s) {
  s->ncalls++;
  const int nlocal = s->nlocal, nall = s->nlocal + s->nghost;
  if (nall > s->neigh_rows) {
    s->neigh_rows = nall;
    free(s->numneigh); free(s->neighbors);
    s->numneigh = (int *)xmalloc(sizeof(int) * (size_t)s->neigh_rows);
    s->neighbors = (int *)xmalloc(sizeof(int) * (size_t)s->neigh_rows * s->maxneighs);
  }
  orc_binatoms(s, -1);
  const real *x = s->x;
  const int *type = s->type;
  const int nt = s->ntypes, half = s->halfneigh, gn = s->ghost_newton;
  int again = 1;
  while (again) {
    again = 0;
    int longest = s->maxneighs;
    for (int i = 0; i < nlocal; i++) {
      int *row = &s->neighbors[(size_t)i * s->maxneighs];
      int n = 0;
      const real xi = x[i * PAD + 0], yi = x[i * PAD + 1], zi = x[i * PAD + 2];
      const int ti = type[i];
      const int ibin = orc_coord2bin(s, xi, yi, zi);
      for (int k = 0; k < s->nstencil; k++) {
        const int jbin = ibin + s->stencil[k];
        const int *cell = &s->bins[(size_t)jbin * s->atoms_per_bin];
        const int cnt = s->bincount[jbin];
        for (int m = 0; m < cnt; m++) {
          const int j = cell[m];
          if (ibin == jbin) {
            if (j == i) continue;
            if (half && !gn && j < i) continue;
            if (half && gn) {
              if (j < i) continue;
              if (j >= nlocal) {
                const real xj = x[j * PAD + 0], yj = x[j * PAD + 1], zj = x[j * PAD + 2];
                if ((zj < zi) || (zj == zi && yj < yi) || (zj == zi && yj == yi && xj < xi)) continue;
              }
            }
          } else {
            if (half && !gn && j < i) continue;
          }
          const real delx = xi - x[j * PAD + 0];
          const real dely = yi - x[j * PAD + 1];
          const real delz = zi - x[j * PAD + 2];
          const real rsq = delx * delx + dely * dely + delz * delz;
          if (rsq <= s->cutneighsq[ti * nt + type[j]]) {
            if (n < s->maxneighs) row[n] = j;
            n++;
          }
        }
      }
      s->numneigh[i] = n;
      if (n >= s->maxneighs) {
        again = 1;
        if (n >= longest) longest = n;
      }
    }
    if (again) {
      s->maxneighs = longest * 1.2;
      free(s->neighbors);
      s->neighbors = (int *)xmalloc(sizeof(int) * (size_t)s->neigh_rows * s->maxneighs);
    }
  }
}

/* Atom::sort (ref/atom.cpp:355-421): bin the LOCAL atoms, inclusive-scan the counts in place,
   then gather x, v, type bin by bin into the alternate buffers and swap pointers.
   f is not permuted (it is recomputed before its next use). */
void orc_sort(orc_sim *s) {
  orc_binatoms(s, s->nlocal);
  int *pos = s->bincount;
  for (int b = 1; b < s->mbins; b++) pos[b] += pos[b - 1];
  for (int b = 0; b < s->mbins; b++) {
    const int start = b > 0 ? pos[b - 1] : 0;
    const int cnt = pos[b] - start;
    for (int k = 0; k < cnt; k++) {
      const int dst = start + k, src = s->bins[(size_t)b * s->atoms_per_bin + k];
      for (int c = 0; c < 3; c++) {
        s->x_alt[dst * PAD + c] = s->x[src * PAD + c];
        s->v_alt[dst * PAD + c] = s->v[src * PAD + c];
      }
      s->type_alt[dst] = s->type[src];
    }
  }
  real *t;
  int *ti;
  t = s->x; s->x = s->x_alt; s->x_alt = t;
  t = s->v; s->v = s->v_alt; s->v_alt = t;
  ti = s->type; s->type = s->type_alt; s->type_alt = ti;
}

/* ------------------------------------------------------------------------------------- */
/* single-rank Comm (ref/comm.cpp)                                                        */
/* ------------------------------------------------------------------------------------- */

/* Comm::setup for procgrid 1x1x1 (ref/comm.cpp:148-269): per dimension `need` layers of two
   swaps; even swaps send the low slab "down" (+prd shift), odd swaps the high slab "up". */
int orc_comm_setup(orc_sim *s) {
  real prd[3] = {s->box.xprd, s->box.yprd, s->box.zprd};
  real lo3[3] = {s->box.xlo, s->box.ylo, s->box.zlo};
  real hi3[3] = {s->box.xhi, s->box.yhi, s->box.zhi};
  const int procgrid = 1, myloc = 0;
  for (int d = 0; d < 3; d++) s->need[d] = (int)(s->cutneigh * procgrid / prd[d] + 1);
  if (2 * (s->need[0] + s->need[1] + s->need[2]) > ORC_MAX_SWAP) return 1;
  s->nswap = 0;
  for (int d = 0; d < 3; d++)
    for (int layer = 0; layer < 2 * s->need[d]; layer++) {
      int w = s->nswap;
      s->pbc_any[w] = s->pbc_flagx[w] = s->pbc_flagy[w] = s->pbc_flagz[w] = 0;
      real lo, hi;
      int *flag = d == 0 ? &s->pbc_flagx[w] : (d == 1 ? &s->pbc_flagy[w] : &s->pbc_flagz[w]);
      if (layer % 2 == 0) {
        int nbox = myloc + layer / 2;
        lo = nbox * prd[d] / procgrid;
        hi = lo3[d] + s->cutneigh;
        real cap = (nbox + 1) * prd[d] / procgrid;
        hi = hi < cap ? hi : cap;
        if (myloc == 0) { s->pbc_any[w] = 1; *flag = 1; }
      } else {
        int nbox = myloc - layer / 2;
        hi = (nbox + 1) * prd[d] / procgrid;
        lo = hi3[d] - s->cutneigh;
        real floor_ = nbox * prd[d] / procgrid;
        lo = lo > floor_ ? lo : floor_;
        if (myloc == procgrid - 1) { s->pbc_any[w] = 1; *flag = -1; }
      }
      s->slablo[w] = lo;
      s->slabhi[w] = hi;
      s->nswap++;
    }
  return 0;
}

/* Atom::pbc (ref/atom.cpp:106-122); Comm::exchange on one rank is only this
   (ref/comm.cpp:364-385: every dimension is skipped when procgrid[idim]==1). */
void orc_pbc(orc_sim *s) {
  real prd[3] = {s->box.xprd, s->box.yprd, s->box.zprd};
  for (int i = 0; i < s->nlocal; i++)
    for (int c = 0; c < 3; c++) {
      if (s->x[i * PAD + c] < 0.0) s->x[i * PAD + c] += prd[c];
      if (s->x[i * PAD + c] >= prd[c]) s->x[i * PAD + c] -= prd[c];
    }
}

/* Comm::borders with every swap a self-swap (ref/comm.cpp:700-883, pack/unpack_border
   ref/atom.cpp:197-226).  Slab membership uses >= lo && <= hi (both inclusive, :776); the
   first swap of a layer pair fixes the scanned range [nfirst,nlast) for both (:759-762);
   ghosts are appended as x + flag*prd with the sender's type. */
void orc_borders(orc_sim *s) {
  s->nghost = 0;
  real prd[3] = {s->box.xprd, s->box.yprd, s->box.zprd};
  int w = 0;
  for (int d = 0; d < 3; d++) {
    int nfirst = 0, nlast = 0;
    for (int layer = 0; layer < 2 * s->need[d]; layer++, w++) {
      if (layer % 2 == 0) { nfirst = nlast; nlast = s->nlocal + s->nghost; }
      const real lo = s->slablo[w], hi = s->slabhi[w];
      int nsend = 0;
      for (int i = nfirst; i < nlast; i++)
        if (s->x[i * PAD + d] >= lo && s->x[i * PAD + d] <= hi) {
          if (nsend >= s->maxsendlist[w]) {
            s->maxsendlist[w] = (int)(1.5 * (nsend + 1)) + 1000;
            s->sendlist[w] = (int *)realloc(s->sendlist[w], sizeof(int) * s->maxsendlist[w]);
          }
          s->sendlist[w][nsend++] = i;
        }
      const int first = s->nlocal + s->nghost;
      atoms_reserve(s, first + nsend);
      const int shift[3] = {s->pbc_flagx[w], s->pbc_flagy[w], s->pbc_flagz[w]};
      for (int k = 0; k < nsend; k++) {
        const int src = s->sendlist[w][k], dst = first + k;
        if (s->pbc_any[w] == 0) {
          for (int c = 0; c < 3; c++) s->x[dst * PAD + c] = s->x[src * PAD + c];
        } else {
          for (int c = 0; c < 3; c++) s->x[dst * PAD + c] = s->x[src * PAD + c] + shift[c] * prd[c];
        }
        s->type[dst] = s->type[src];
      }
      s->sendnum[w] = nsend;
      s->recvnum[w] = nsend;
      s->firstrecv[w] = first;
      s->nghost += nsend;
    }
  }
}

/* Comm::communicate, self-swaps (ref/comm.cpp:276-317; Atom::pack_comm/unpack_comm
   ref/atom.cpp:135-170): swaps in order, x[first+k] = x[list[k]] + flag*prd. */
void orc_communicate(orc_sim *s) {
  real prd[3] = {s->box.xprd, s->box.yprd, s->box.zprd};
  for (int w = 0; w < s->nswap; w++) {
    const int shift[3] = {s->pbc_flagx[w], s->pbc_flagy[w], s->pbc_flagz[w]};
    for (int k = 0; k < s->sendnum[w]; k++) {
      const int src = s->sendlist[w][k], dst = s->firstrecv[w] + k;
      if (s->pbc_any[w] == 0) {
        for (int c = 0; c < 3; c++) s->x[dst * PAD + c] = s->x[src * PAD + c];
      } else {
        for (int c = 0; c < 3; c++) s->x[dst * PAD + c] = s->x[src * PAD + c] + shift[c] * prd[c];
      }
    }
  }
}

/* Comm::reverse_communicate, self-swaps (ref/comm.cpp:321-355; pack/unpack_reverse
   ref/atom.cpp:172-195): swaps in REVERSE order, f[list[k]] += f[first+k]. */
void orc_reverse_communicate(orc_sim *s) {
  for (int w = s->nswap - 1; w >= 0; w--)
    for (int k = 0; k < s->sendnum[w]; k++) {
      const int dst = s->sendlist[w][k], src = s->firstrecv[w] + k;
      for (int c = 0; c < 3; c++) s->f[dst * PAD + c] += s->f[src * PAD + c];
    }
}

/* ------------------------------------------------------------------------------------- */
/* Lennard-Jones force (ref/force_lj.cpp)                                                 */
/* ------------------------------------------------------------------------------------- */

/* ForceLJ::setup (ref/force_lj.cpp:65-69). */
void orc_force_lj_setup(orc_sim *s) {
  for (int i = 0; i < s->ntypes * s->ntypes; i++) s->cutforcesq[i] = s->cutforce * s->cutforce;
}

/* ForceLJ::compute_halfneigh<EVFLAG,GHOST_NEWTON> (ref/force_lj.cpp:185-263): clear f over
   local+ghost; per stored pair inside the cutoff apply +F to i and -F to j (j only if
   ghost_newton or j local); energy/virial pairs with an un-updated ghost count half. */
static void lj_half(orc_sim *s) {
  const int nlocal = s->nlocal, nall = s->nlocal + s->nghost, nt = s->ntypes, gn = s->ghost_newton;
  const real *x = s->x;
  real *f = s->f;
  const int *type = s->type;
  for (int i = 0; i < nall; i++) { f[i * PAD + 0] = 0.0; f[i * PAD + 1] = 0.0; f[i * PAD + 2] = 0.0; }
  real t_energy = 0, t_virial = 0;
  for (int i = 0; i < nlocal; i++) {
    const int *row = &s->neighbors[(size_t)i * s->maxneighs];
    const int cnt = s->numneigh[i];
    const real xi = x[i * PAD + 0], yi = x[i * PAD + 1], zi = x[i * PAD + 2];
    const int ti = type[i];
    real fix = 0.0, fiy = 0.0, fiz = 0.0;
    for (int k = 0; k < cnt; k++) {
      const int j = row[k];
      const real delx = xi - x[j * PAD + 0];
      const real dely = yi - x[j * PAD + 1];
      const real delz = zi - x[j * PAD + 2];
      const real rsq = delx * delx + dely * dely + delz * delz;
      const int tij = ti * nt + type[j];
      if (rsq < s->cutforcesq[tij]) {
        const real sr2 = 1.0 / rsq;
        const real sr6 = sr2 * sr2 * sr2 * s->sigma6[tij];
        const real force = 48.0 * sr6 * (sr6 - 0.5) * sr2 * s->epsilon[tij];
        fix += delx * force;
        fiy += dely * force;
        fiz += delz * force;
        if (gn || j < nlocal) {
          f[j * PAD + 0] -= delx * force;
          f[j * PAD + 1] -= dely * force;
          f[j * PAD + 2] -= delz * force;
        }
        if (s->evflag) {
          const real scale = (gn || j < nlocal) ? 1.0 : 0.5;
          t_energy += scale * (4.0 * sr6 * (sr6 - 1.0)) * s->epsilon[tij];
          t_virial += scale * (delx * delx + dely * dely + delz * delz) * force;
        }
      }
    }
    f[i * PAD + 0] += fix;
    f[i * PAD + 1] += fiy;
    f[i * PAD + 2] += fiz;
  }
  s->eng_vdwl += t_energy;
  s->virial += t_virial;
}

/* ForceLJ::compute_fullneigh<EVFLAG> (ref/force_lj.cpp:366-449): clear f over locals only,
   no j update; energy accumulates sr6(sr6-1)eps then *4, virial *0.5 (each pair seen twice). */
static void lj_full(orc_sim *s) {
  const int nlocal = s->nlocal, nt = s->ntypes;
  const real *x = s->x;
  real *f = s->f;
  const int *type = s->type;
  real t_eng = 0, t_vir = 0;
  for (int i = 0; i < nlocal; i++) { f[i * PAD + 0] = 0.0; f[i * PAD + 1] = 0.0; f[i * PAD + 2] = 0.0; }
  for (int i = 0; i < nlocal; i++) {
    const int *row = &s->neighbors[(size_t)i * s->maxneighs];
    const int cnt = s->numneigh[i];
    const real xi = x[i * PAD + 0], yi = x[i * PAD + 1], zi = x[i * PAD + 2];
    const int ti = type[i];
    real fix = 0, fiy = 0, fiz = 0;
    for (int k = 0; k < cnt; k++) {
      const int j = row[k];
      const real delx = xi - x[j * PAD + 0];
      const real dely = yi - x[j * PAD + 1];
      const real delz = zi - x[j * PAD + 2];
      const real rsq = delx * delx + dely * dely + delz * delz;
      const int tij = ti * nt + type[j];
      if (rsq < s->cutforcesq[tij]) {
        const real sr2 = 1.0 / rsq;
        const real sr6 = sr2 * sr2 * sr2 * s->sigma6[tij];
        const real force = 48.0 * sr6 * (sr6 - 0.5) * sr2 * s->epsilon[tij];
        fix += delx * force;
        fiy += dely * force;
        fiz += delz * force;
        if (s->evflag) {
          t_eng += sr6 * (sr6 - 1.0) * s->epsilon[tij];
          t_vir += (delx * delx + dely * dely + delz * delz) * force;
        }
      }
    }
    f[i * PAD + 0] += fix;
    f[i * PAD + 1] += fiy;
    f[i * PAD + 2] += fiz;
  }
  t_eng *= 4.0;
  t_vir *= 0.5;
  s->eng_vdwl += t_eng;
  s->virial += t_vir;
}

/* ------------------------------------------------------------------------------------- */
/* EAM (ref/force_eam.cpp)                                                                */
/* ------------------------------------------------------------------------------------- */

static int read_reals(FILE *fp, int n, real *dst) { /* ForceEAM::grab (ref/force_eam.cpp:800-815) */
  char line[1024];
  int got = 0;
  while (got < n) {
    if (!fgets(line, sizeof line, fp)) return 1;
    for (char *tok = strtok(line, " \t\n\r\f"); tok; tok = strtok(NULL, " \t\n\r\f")) dst[got++] = atof(tok);
  }
  return 0;
}

/* four-point Lagrange re-grid used by ForceEAM::file2array (ref/force_eam.cpp:630-726);
   tables are 1-based like the reference's (shift at :575-579). */
static double lagrange4(const real *tab, int ntab, double dtab, double r) {
  const double sixth = 1.0 / 6.0;
  double p = r / dtab + 1.0;
  int k = (int)p;
  k = imin(k, ntab - 2);
  k = imax(k, 2);
  p -= k;
  p = p < 2.0 ? p : 2.0;
  double c1 = -sixth * p * (p - 1.0) * (p - 2.0);
  double c2 = 0.5 * (p * p - 1.0) * (p - 2.0);
  double c3 = -0.5 * p * (p + 1.0) * (p - 2.0);
  double c4 = sixth * p * (p * p - 1.0);
  return c1 * tab[k - 1] + c2 * tab[k] + c3 * tab[k + 1] + c4 * tab[k + 2];
}

/* ForceEAM::interpolate (ref/force_eam.cpp:765-793): 7 coefficients per knot; [3..6] value
   cubic, [0..2] its derivative; 1-based knots. */
static void spline7(int n, real delta, const real *f, real *sp) {
  for (int m = 1; m <= n; m++) sp[m * 7 + 6] = f[m];
  sp[1 * 7 + 5] = sp[2 * 7 + 6] - sp[1 * 7 + 6];
  sp[2 * 7 + 5] = 0.5 * (sp[3 * 7 + 6] - sp[1 * 7 + 6]);
  sp[(n - 1) * 7 + 5] = 0.5 * (sp[n * 7 + 6] - sp[(n - 2) * 7 + 6]);
  sp[n * 7 + 5] = sp[n * 7 + 6] - sp[(n - 1) * 7 + 6];
  for (int m = 3; m <= n - 2; m++)
    sp[m * 7 + 5] = ((sp[(m - 2) * 7 + 6] - sp[(m + 2) * 7 + 6]) + 8.0 * (sp[(m + 1) * 7 + 6] - sp[(m - 1) * 7 + 6])) / 12.0;
  for (int m = 1; m <= n - 1; m++) {
    sp[m * 7 + 4] = 3.0 * (sp[(m + 1) * 7 + 6] - sp[m * 7 + 6]) - 2.0 * sp[m * 7 + 5] - sp[(m + 1) * 7 + 5];
    sp[m * 7 + 3] = sp[m * 7 + 5] + sp[(m + 1) * 7 + 5] - 2.0 * (sp[(m + 1) * 7 + 6] - sp[m * 7 + 6]);
  }
  sp[n * 7 + 4] = 0.0;
  sp[n * 7 + 3] = 0.0;
  for (int m = 1; m <= n; m++) {
    sp[m * 7 + 2] = sp[m * 7 + 5] / delta;
    sp[m * 7 + 1] = 2.0 * sp[m * 7 + 4] / delta;
    sp[m * 7 + 0] = 3.0 * sp[m * 7 + 3] / delta;
  }
}

/* ForceEAM::setup = coeff + init_style (ref/force_eam.cpp:74-79, 457-487): read the DYNAMO
   funcfl file (:505-582), re-grid (:589-728), spline (:732-761), replicate per type pair. */
int orc_force_eam_setup(orc_sim *s, const char *path) {
  FILE *fp = fopen(path, "r");
  if (!fp) return 1;
  char line[1024];
  int itmp, fnrho, fnr;
  double fmass, fdrho, fdr, fcut;
  if (!fgets(line, sizeof line, fp) || !fgets(line, sizeof line, fp)) { fclose(fp); return 2; }
  sscanf(line, "%d %lg", &itmp, &fmass);
  if (!fgets(line, sizeof line, fp)) { fclose(fp); return 2; }
  sscanf(line, "%d %lg %d %lg %lg", &fnrho, &fdrho, &fnr, &fdr, &fcut);
  real *ffrho = (real *)xmalloc(sizeof(real) * (fnrho + 1));
  real *frhor = (real *)xmalloc(sizeof(real) * (fnr + 1));
  real *fzr = (real *)xmalloc(sizeof(real) * (fnr + 1));
  int bad = read_reals(fp, fnrho, ffrho) || read_reals(fp, fnr, fzr) || read_reals(fp, fnr, frhor);
  fclose(fp);
  if (bad) return 3;
  for (int i = fnrho; i > 0; i--) ffrho[i] = ffrho[i - 1];
  for (int i = fnr; i > 0; i--) frhor[i] = frhor[i - 1];
  for (int i = fnr; i > 0; i--) fzr[i] = fzr[i - 1];
  s->eam_mass = fmass;
  s->eam_cut = fcut;
  s->cutforce = fcut; /* not in the reference (it prints in.force_cut); kept for queries */
  for (int i = 0; i < s->ntypes * s->ntypes; i++) s->cutforcesq[i] = s->eam_cut * s->eam_cut;

  s->dr = fdr;
  s->drho = fdrho;
  double rmax = (fnr - 1) * fdr, rhomax = (fnrho - 1) * fdrho;
  s->nr = (int)(rmax / s->dr + 0.5);
  s->nrho = (int)(rhomax / s->drho + 0.5);
  real *frho = (real *)xmalloc(sizeof(real) * (s->nrho + 1));
  real *rhor = (real *)xmalloc(sizeof(real) * (s->nr + 1));
  real *z2r = (real *)xmalloc(sizeof(real) * (s->nr + 1));
  for (int m = 1; m <= s->nrho; m++) frho[m] = lagrange4(ffrho, fnrho, fdrho, (m - 1) * s->drho);
  for (int m = 1; m <= s->nr; m++) rhor[m] = lagrange4(frhor, fnr, fdr, (m - 1) * s->dr);
  for (int m = 1; m <= s->nr; m++) {
    double r = (m - 1) * s->dr;
    double zri = lagrange4(fzr, fnr, fdr, r);
    double zrj = lagrange4(fzr, fnr, fdr, r);
    z2r[m] = 27.2 * 0.529 * zri * zrj;
  }
  s->rdr = 1.0 / s->dr;
  s->rdrho = 1.0 / s->drho;
  s->nrho_tot = (s->nrho + 1) * 7 + 64;
  s->nr_tot = (s->nr + 1) * 7 + 64;
  s->nrho_tot -= s->nrho_tot % 64;
  s->nr_tot -= s->nr_tot % 64;
  int nn = s->ntypes * s->ntypes;
  free(s->frho_spline); free(s->rhor_spline); free(s->z2r_spline);
  s->frho_spline = (real *)calloc((size_t)nn * s->nrho_tot, sizeof(real));
  s->rhor_spline = (real *)calloc((size_t)nn * s->nr_tot, sizeof(real));
  s->z2r_spline = (real *)calloc((size_t)nn * s->nr_tot, sizeof(real));
  spline7(s->nrho, s->drho, frho, s->frho_spline);
  spline7(s->nr, s->dr, rhor, s->rhor_spline);
  spline7(s->nr, s->dr, z2r, s->z2r_spline);
  for (int t = 1; t < nn; t++) {
    memcpy(s->frho_spline + (size_t)t * s->nrho_tot, s->frho_spline, sizeof(real) * s->nrho_tot);
    memcpy(s->rhor_spline + (size_t)t * s->nr_tot, s->rhor_spline, sizeof(real) * s->nr_tot);
    memcpy(s->z2r_spline + (size_t)t * s->nr_tot, s->z2r_spline, sizeof(real) * s->nr_tot);
  }
  free(ffrho); free(frhor); free(fzr); free(frho); free(rhor); free(z2r);
  s->forcetype = 1;
  return 0;
}

static void eam_reserve(orc_sim *s) {
  if (s->nmax > s->eam_nmax) {
    s->eam_nmax = s->nmax;
    free(s->rho); free(s->fp);
    s->rho = (real *)xmalloc(sizeof(real) * s->eam_nmax);
    s->fp = (real *)xmalloc(sizeof(real) * s->eam_nmax);
  }
}

/* ForceEAM::communicate (ref/force_eam.cpp:851-914): forward halo of fp, one value per ghost. */
static void eam_fp_halo(orc_sim *s) {
  for (int w = 0; w < s->nswap; w++)
    for (int k = 0; k < s->sendnum[w]; k++) s->fp[s->firstrecv[w] + k] = s->fp[s->sendlist[w][k]];
}

#define RHO_VAL(sp, p) ((((sp)[3] * (p) + (sp)[4]) * (p) + (sp)[5]) * (p) + (sp)[6])
#define QUAD_DER(sp, p) (((sp)[0] * (p) + (sp)[1]) * (p) + (sp)[2])

/* embedding pass shared by both list styles (ref/force_eam.cpp:172-185 / :336-347), incl. the
   type_ii = type*type indexing quirk. */
static real eam_embed(orc_sim *s, int i, real rho_i, real *evdwl) {
  const int tii = s->type[i] * s->type[i];
  real p = 1.0 * rho_i * s->rdrho + 1.0;
  int m = (int)p;
  m = imax(1, imin(m, s->nrho - 1));
  p -= m;
  p = p < 1.0 ? p : 1.0;
  const real *sp = &s->frho_spline[(size_t)tii * s->nrho_tot + m * 7];
  if (s->evflag) *evdwl += RHO_VAL(sp, p);
  return QUAD_DER(sp, p);
}

/* ForceEAM::compute_halfneigh (ref/force_eam.cpp:94-270): rho pass with j-scatter for local
   j, embed pass, fp halo, pair pass (energy accumulated regardless of evflag; virial only
   with evflag; ghost pairs count half). */
static void eam_half(orc_sim *s) {
  real evdwl = 0.0;
  s->virial = 0;
  eam_reserve(s);
  const int nlocal = s->nlocal, nall = s->nlocal + s->nghost, nt = s->ntypes;
  const real *x = s->x;
  real *f = s->f, *rho = s->rho, *fp = s->fp;
  const int *type = s->type;
  for (int i = 0; i < nall; i++) { f[i * PAD + 0] = 0; f[i * PAD + 1] = 0; f[i * PAD + 2] = 0; }
  for (int i = 0; i < nlocal; i++) rho[i] = 0.0;
  for (int i = 0; i < nlocal; i++) {
    const int *row = &s->neighbors[(size_t)i * s->maxneighs];
    const int cnt = s->numneigh[i];
    const real xi = x[i * PAD + 0], yi = x[i * PAD + 1], zi = x[i * PAD + 2];
    const int ti = type[i];
    real rhoi = 0.0;
    for (int jj = 0; jj < cnt; jj++) {
      const int j = row[jj];
      const real delx = xi - x[j * PAD + 0];
      const real dely = yi - x[j * PAD + 1];
      const real delz = zi - x[j * PAD + 2];
      const real rsq = delx * delx + dely * dely + delz * delz;
      const int tij = ti * nt + type[j];
      if (rsq < s->cutforcesq[tij]) {
        real p = REAL_SQRT(rsq) * s->rdr + 1.0;
        int m = (int)p;
        m = m < s->nr - 1 ? m : s->nr - 1;
        p -= m;
        p = p < 1.0 ? p : 1.0;
        const real *sp = &s->rhor_spline[(size_t)tij * s->nr_tot + m * 7];
        rhoi += RHO_VAL(sp, p);
        if (j < nlocal) rho[j] += RHO_VAL(sp, p);
      }
    }
    rho[i] += rhoi;
  }
  for (int i = 0; i < nlocal; i++) fp[i] = eam_embed(s, i, rho[i], &evdwl);
  eam_fp_halo(s);
  for (int i = 0; i < nlocal; i++) {
    const int *row = &s->neighbors[(size_t)i * s->maxneighs];
    const int cnt = s->numneigh[i];
    const real xi = x[i * PAD + 0], yi = x[i * PAD + 1], zi = x[i * PAD + 2];
    const int ti = type[i];
    real fx = 0, fy = 0, fz = 0;
    for (int jj = 0; jj < cnt; jj++) {
      const int j = row[jj];
      const real delx = xi - x[j * PAD + 0];
      const real dely = yi - x[j * PAD + 1];
      const real delz = zi - x[j * PAD + 2];
      const real rsq = delx * delx + dely * dely + delz * delz;
      const int tij = ti * nt + type[j];
      if (rsq < s->cutforcesq[tij]) {
        real r = REAL_SQRT(rsq);
        real p = r * s->rdr + 1.0;
        int m = (int)p;
        m = m < s->nr - 1 ? m : s->nr - 1;
        p -= m;
        p = p < 1.0 ? p : 1.0;
        const real *rs = &s->rhor_spline[(size_t)tij * s->nr_tot + m * 7];
        const real *zs = &s->z2r_spline[(size_t)tij * s->nr_tot + m * 7];
        real rhoip = QUAD_DER(rs, p);
        real z2p = QUAD_DER(zs, p);
        real z2 = RHO_VAL(zs, p);
        real recip = 1.0 / r;
        real phi = z2 * recip;
        real phip = z2p * recip - phi * recip;
        real psip = fp[i] * rhoip + fp[j] * rhoip + phip;
        real fpair = -psip * recip;
        fx += delx * fpair;
        fy += dely * fpair;
        fz += delz * fpair;
        if (j < nlocal) {
          f[j * PAD + 0] -= delx * fpair;
          f[j * PAD + 1] -= dely * fpair;
          f[j * PAD + 2] -= delz * fpair;
        } else fpair *= 0.5;
        if (s->evflag) s->virial += delx * delx * fpair + dely * dely * fpair + delz * delz * fpair;
        if (j < nlocal) evdwl += phi;
        else evdwl += 0.5 * phi;
      }
    }
    f[i * PAD + 0] += fx;
    f[i * PAD + 1] += fy;
    f[i * PAD + 2] += fz;
  }
  s->eng_vdwl = evdwl;
}

/* ForceEAM::compute_fullneigh (ref/force_eam.cpp:274-449): rho+embed fused pass, fp halo,
   pair pass with plain store of f[i]; eng_vdwl += 2*evdwl. */
static void eam_full(orc_sim *s) {
  real evdwl = 0.0;
  s->eng_vdwl = 0;
  s->virial = 0;
  eam_reserve(s);
  const int nlocal = s->nlocal, nt = s->ntypes;
  const real *x = s->x;
  real *f = s->f, *fp = s->fp;
  const int *type = s->type;
  for (int i = 0; i < nlocal; i++) {
    const int *row = &s->neighbors[(size_t)i * s->maxneighs];
    const int cnt = s->numneigh[i];
    const real xi = x[i * PAD + 0], yi = x[i * PAD + 1], zi = x[i * PAD + 2];
    const int ti = type[i];
    real rhoi = 0;
    for (int jj = 0; jj < cnt; jj++) {
      const int j = row[jj];
      const real delx = xi - x[j * PAD + 0];
      const real dely = yi - x[j * PAD + 1];
      const real delz = zi - x[j * PAD + 2];
      const real rsq = delx * delx + dely * dely + delz * delz;
      const int tij = ti * nt + type[j];
      if (rsq < s->cutforcesq[tij]) {
        real p = REAL_SQRT(rsq) * s->rdr + 1.0;
        int m = (int)p;
        m = m < s->nr - 1 ? m : s->nr - 1;
        p -= m;
        p = p < 1.0 ? p : 1.0;
        const real *sp = &s->rhor_spline[(size_t)tij * s->nr_tot + m * 7];
        rhoi += RHO_VAL(sp, p);
      }
    }
    fp[i] = eam_embed(s, i, rhoi, &evdwl);
  }
  eam_fp_halo(s);
  real t_virial = 0;
  for (int i = 0; i < nlocal; i++) {
    const int *row = &s->neighbors[(size_t)i * s->maxneighs];
    const int cnt = s->numneigh[i];
    const real xi = x[i * PAD + 0], yi = x[i * PAD + 1], zi = x[i * PAD + 2];
    const int ti = type[i];
    real fx = 0.0, fy = 0.0, fz = 0.0;
    for (int jj = 0; jj < cnt; jj++) {
      const int j = row[jj];
      const real delx = xi - x[j * PAD + 0];
      const real dely = yi - x[j * PAD + 1];
      const real delz = zi - x[j * PAD + 2];
      const real rsq = delx * delx + dely * dely + delz * delz;
      const int tij = ti * nt + type[j];
      if (rsq < s->cutforcesq[tij]) {
        real r = REAL_SQRT(rsq);
        real p = r * s->rdr + 1.0;
        int m = (int)p;
        m = m < s->nr - 1 ? m : s->nr - 1;
        p -= m;
        p = p < 1.0 ? p : 1.0;
        const real *rs = &s->rhor_spline[(size_t)tij * s->nr_tot + m * 7];
        const real *zs = &s->z2r_spline[(size_t)tij * s->nr_tot + m * 7];
        real rhoip = QUAD_DER(rs, p);
        real z2p = QUAD_DER(zs, p);
        real z2 = RHO_VAL(zs, p);
        real recip = 1.0 / r;
        real phi = z2 * recip;
        real phip = z2p * recip - phi * recip;
        real psip = fp[i] * rhoip + fp[j] * rhoip + phip;
        real fpair = -psip * recip;
        fx += delx * fpair;
        fy += dely * fpair;
        fz += delz * fpair;
        fpair *= 0.5;
        if (s->evflag) {
          t_virial += delx * delx * fpair + dely * dely * fpair + delz * delz * fpair;
          evdwl += 0.5 * phi;
        }
      }
    }
    f[i * PAD + 0] = fx;
    f[i * PAD + 1] = fy;
    f[i * PAD + 2] = fz;
  }
  s->virial += t_virial;
  s->eng_vdwl += 2.0 * evdwl;
}

/* Force::compute dispatch (ref/force_lj.cpp:72-113, ref/force_eam.cpp:82-91). */
void orc_force_compute(orc_sim *s) {
  if (s->forcetype == 0) {
    s->eng_vdwl = 0;
    s->virial = 0;
    if (s->halfneigh) lj_half(s); else lj_full(s);
  } else {
    if (s->halfneigh) eam_half(s); else eam_full(s);
  }
}

/* ------------------------------------------------------------------------------------- */
/* velocity Verlet + thermo                                                               */
/* ------------------------------------------------------------------------------------- */

/* Integrate::initialIntegrate (ref/integrate.cpp:46-57). */
void orc_initial_integrate(orc_sim *s) {
  real *x = s->x, *v = s->v;
  const real *f = s->f;
  for (int i = 0; i < s->nlocal; i++) {
    v[i * PAD + 0] += s->dtforce * f[i * PAD + 0];
    v[i * PAD + 1] += s->dtforce * f[i * PAD + 1];
    v[i * PAD + 2] += s->dtforce * f[i * PAD + 2];
    x[i * PAD + 0] += s->dt * v[i * PAD + 0];
    x[i * PAD + 1] += s->dt * v[i * PAD + 1];
    x[i * PAD + 2] += s->dt * v[i * PAD + 2];
  }
}

/* Integrate::finalIntegrate (ref/integrate.cpp:59-68). */
void orc_final_integrate(orc_sim *s) {
  real *v = s->v;
  const real *f = s->f;
  for (int i = 0; i < s->nlocal; i++) {
    v[i * PAD + 0] += s->dtforce * f[i * PAD + 0];
    v[i * PAD + 1] += s->dtforce * f[i * PAD + 1];
    v[i * PAD + 2] += s->dtforce * f[i * PAD + 2];
  }
}

/* Thermo::compute -> energy/pressure (ref/thermo.cpp:74-136, 181-194): values appended to the
   in-memory log instead of printed. */
void orc_thermo_record(orc_sim *s, int step) {
  real t = orc_temperature(s);
  real e = s->eng_vdwl;
  if (s->halfneigh) e *= 2.0;
  e *= s->e_scale;
  real eng = e / s->natoms;
  real p = (t * s->dof_boltz + s->virial) * s->p_scale;
  if (s->nlog == s->log_cap) {
    s->log_cap = s->log_cap ? 2 * s->log_cap : 64;
    s->log_step = (int *)realloc(s->log_step, sizeof(int) * s->log_cap);
    s->log_t = (double *)realloc(s->log_t, sizeof(double) * s->log_cap);
    s->log_e = (double *)realloc(s->log_e, sizeof(double) * s->log_cap);
    s->log_p = (double *)realloc(s->log_p, sizeof(double) * s->log_cap);
  }
  s->log_step[s->nlog] = step;
  s->log_t[s->nlog] = t;
  s->log_e[s->nlog] = eng;
  s->log_p[s->nlog] = p;
  s->nlog++;
}

/* ------------------------------------------------------------------------------------- */
/* whole-run drivers                                                                      */
/* ------------------------------------------------------------------------------------- */

typedef struct {
  int nx, ny, nz;
  int ntimes;
  int forcetype;      /* 0 lj, 1 eam */
  int units;          /* 0 lj, 1 metal */
  int halfneigh;      /* 1 half, 0 full */
  int ghost_newton;
  int neigh_every;
  int sort;           /* the --sort CLI value: -1 => reneigh frequency, 0 never, n>0 */
  int thermo_nstat;
  int nbin_override;  /* -b ; <=0 => 5/6 rule */
  double epsilon, sigma;
  double dt, t_request, rho, force_cut, neigh_cut; /* neigh_cut = force_cut + skin already */
} orc_config;

/* main() up to and including the step-0 thermo line (ref/ljs.cpp:263-468).  The caller seeds
   libc rand with 5413 first (ref/ljs.cpp:110) if reference-identical types are wanted. */
int orc_init(orc_sim *s, const orc_config *c, const char *eam_file) {
  s->forcetype = c->forcetype;
  int gn = c->ghost_newton;
  if (c->forcetype == 1 && gn == 1) gn = 0; /* ref/ljs.cpp:277-282 */
  if (c->forcetype == 0) {
    for (int i = 0; i < s->ntypes * s->ntypes; i++) {
      s->epsilon[i] = c->epsilon;
      real sg = c->sigma;
      s->sigma6[i] = sg * sg * sg * sg * sg * sg;
    }
  }
  s->ghost_newton = gn;
  s->halfneigh = c->halfneigh;
  if (c->nbin_override > 0) {
    s->nbinx = s->nbiny = s->nbinz = c->nbin_override;
  } else {
    real neighscale = 5.0 / 6.0; /* ref/ljs.cpp:357-362 */
    s->nbinx = neighscale * c->nx;
    s->nbiny = neighscale * c->ny;
    s->nbinz = neighscale * c->nz;
  }
  if (s->nbinx == 0) s->nbinx = 1;
  if (s->nbiny == 0) s->nbiny = 1;
  if (s->nbinz == 0) s->nbinz = 1;
  s->ntimes = c->ntimes;
  s->sort_every = c->sort > 0 ? c->sort : (c->sort < 0 ? c->neigh_every : 0);
  s->every = c->neigh_every;
  s->cutneigh = c->neigh_cut;
  s->cutforce = c->force_cut;
  s->nstat = c->thermo_nstat;

  orc_create_box(s, c->nx, c->ny, c->nz, c->rho);
  if (orc_comm_setup(s)) return 10;
  orc_neighbor_setup(s);
  orc_integrate_setup(s, c->dt);
  if (c->forcetype == 0) orc_force_lj_setup(s);
  else {
    real keep = s->cutforce;
    int rc = orc_force_eam_setup(s, eam_file);
    if (rc) return 20 + rc;
    s->cutforce = keep;
    s->mass = s->eam_mass; /* ref/ljs.cpp:403 */
  }
  if (orc_create_atoms(s, c->nx, c->ny, c->nz, c->rho)) return 30;
  orc_thermo_setup(s, c->units);
  orc_create_velocity(s, c->t_request);

  orc_pbc(s);                       /* comm.exchange (ref/ljs.cpp:445) */
  if (c->sort > 0) orc_sort(s);     /* ref/ljs.cpp:446-447 */
  orc_borders(s);
  s->evflag = 1;
  orc_neighbor_build(s);
  orc_force_compute(s);
  if (s->halfneigh && s->ghost_newton) orc_reverse_communicate(s);
  s->nlog = 0;
  orc_thermo_record(s, 0);
  /* Integrate::run prologue (ref/integrate.cpp:80-81) */
  s->dtforce = s->dtforce / s->mass;
  return 0;
}

/* Integrate::run loop body (ref/integrate.cpp:88-205), steps [first, first+nsteps) of an
   ntimes-step run; `first` lets tests interleave inspection with stepping. */
void orc_run(orc_sim *s, int first, int nsteps) {
  int next_sort = s->sort_every > 0 ? s->sort_every : s->ntimes + 1;
  while (next_sort <= first && s->sort_every > 0) next_sort += s->sort_every;
  for (int n = first; n < first + nsteps; n++) {
    orc_initial_integrate(s);
    if ((n + 1) % s->every) {
      orc_communicate(s);
    } else {
      orc_pbc(s);
      if (n + 1 >= next_sort) { orc_sort(s); next_sort += s->sort_every; }
      orc_borders(s);
      orc_neighbor_build(s);
    }
    s->evflag = s->nstat ? ((n + 1) % s->nstat == 0) : 0;
    orc_force_compute(s);
    if (s->halfneigh && s->ghost_newton) orc_reverse_communicate(s);
    orc_final_integrate(s);
    if (s->nstat && (n + 1) % s->nstat == 0) orc_thermo_record(s, n + 1);
  }
}

/* ------------------------------------------------------------------------------------- */
/* accessors for the ctypes wrapper                                                       */
/* ------------------------------------------------------------------------------------- */
int orc_sizeof_real(void) { return (int)sizeof(real); }
int orc_get_int(const orc_sim *s, const char *k) {
#define K(name) if (!strcmp(k, #name)) return s->name;
  K(natoms) K(nlocal) K(nghost) K(nmax) K(ntypes) K(every) K(nbinx) K(nbiny) K(nbinz) K(maxneighs)
  K(halfneigh) K(ghost_newton) K(mbins) K(atoms_per_bin) K(nstencil) K(mbinx) K(mbiny) K(mbinz)
  K(mbinxlo) K(mbinylo) K(mbinzlo) K(nswap) K(nlog) K(nr) K(nrho) K(nr_tot) K(nrho_tot) K(evflag)
  K(sort_every) K(nstat) K(ntimes) K(forcetype) K(ncalls)
#undef K
  fprintf(stderr, "orc_get_int: unknown key %s\n", k);
  return -2147483647;
}
double orc_get_real(const orc_sim *s, const char *k) {
#define K(name) if (!strcmp(k, #name)) return (double)s->name;
  K(mass) K(cutneigh) K(cutforce) K(eng_vdwl) K(virial) K(dt) K(dtforce) K(t_scale) K(e_scale) K(p_scale)
  K(mvv2e) K(dof_boltz) K(binsizex) K(binsizey) K(binsizez) K(bininvx) K(bininvy) K(bininvz)
  K(rdr) K(rdrho) K(dr) K(drho) K(eam_mass) K(eam_cut)
  K(box.xprd) K(box.yprd) K(box.zprd) K(box.xlo) K(box.xhi) K(box.ylo) K(box.yhi) K(box.zlo) K(box.zhi)
#undef K
  fprintf(stderr, "orc_get_real: unknown key %s\n", k);
  return NAN;
}
void orc_set_int(orc_sim *s, const char *k, int v) {
#define K(name) if (!strcmp(k, #name)) { s->name = v; return; }
  K(halfneigh) K(ghost_newton) K(evflag) K(nbinx) K(nbiny) K(nbinz) K(every) K(sort_every) K(nstat)
  K(ntimes) K(forcetype) K(maxneighs) K(atoms_per_bin) K(nlocal) K(nghost) K(natoms)
#undef K
  fprintf(stderr, "orc_set_int: unknown key %s\n", k);
}
void orc_set_real(orc_sim *s, const char *k, double v) {
#define K(name) if (!strcmp(k, #name)) { s->name = v; return; }
  K(cutneigh) K(cutforce) K(mass) K(dt) K(dtforce) K(eng_vdwl) K(virial)
#undef K
  fprintf(stderr, "orc_set_real: unknown key %s\n", k);
}
void *orc_get_ptr(orc_sim *s, const char *k) {
#define K(name) if (!strcmp(k, #name)) return (void *)s->name;
  K(x) K(v) K(f) K(type) K(numneigh) K(neighbors) K(bincount) K(bins) K(stencil) K(cutneighsq) K(cutforcesq)
  K(epsilon) K(sigma6) K(rhor_spline) K(z2r_spline) K(frho_spline) K(rho) K(fp) K(log_step) K(log_t) K(log_e)
  K(log_p) K(sendnum) K(recvnum) K(firstrecv) K(slablo) K(slabhi) K(pbc_any) K(pbc_flagx) K(pbc_flagy)
  K(pbc_flagz) K(need)
#undef K
  fprintf(stderr, "orc_get_ptr: unknown key %s\n", k);
  return NULL;
}
int *orc_get_sendlist(orc_sim *s, int iswap) { return s->sendlist[iswap]; }
/* make room for externally supplied atoms (tests that inject their own configuration) */
void orc_reserve_atoms(orc_sim *s, int n) { atoms_reserve(s, n); }
