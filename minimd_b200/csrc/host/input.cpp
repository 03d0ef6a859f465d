// in.*.miniMD reader: 14 positional lines, two header lines first, trailing comments ignored.
// Same format and error behaviour as ref/input.cpp:48-187 (values are parsed at MMD_float
// precision, as the reference's sscanf("%e"/"%le") does, and neigh_cut becomes cut + skin).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "ljs.h"

namespace {

const int MAXLINE = 256;

bool next_line(FILE* fp, char* line) { return fgets(line, MAXLINE, fp) != nullptr; }

std::string first_token(const char* line) {
  const char* p = line;
  while (*p == ' ' || *p == '\t') p++;
  const char* e = p;
  while (*e && *e != ' ' && *e != '\t' && *e != '\n' && *e != '\r') e++;
  return std::string(p, e);
}

MMD_float to_real(const char* s, char** end) {
#if PRECISION == 1
  return strtof(s, end);
#else
  return strtod(s, end);
#endif
}

// up to `n` reals from one line; missing trailing values leave the destination untouched (sscanf semantics)
void scan_reals(const char* line, MMD_float** dst, int n) {
  const char* p = line;
  for (int k = 0; k < n; k++) {
    char* end = nullptr;
    MMD_float v = to_real(p, &end);
    if (end == p) return;
    *dst[k] = v;
    p = end;
  }
}

}  // namespace

int input(In& in, const char* filename, int me) {
  FILE* fp = fopen(filename, "r");
  if (!fp) {
    if (me == 0) printf("ERROR: Cannot open %s\n", filename);
    return 1;
  }
  char line[MAXLINE];
  bool ok = next_line(fp, line) && next_line(fp, line);  // title + blank line

  ok = ok && next_line(fp, line);
  std::string tok = ok ? first_token(line) : "";
  if (tok == "lj") in.units = LJ;
  else if (tok == "metal") in.units = METAL;
  else {
    if (me == 0) printf("Unknown units option in file at line 3 ('%s'). Expecting either 'lj' or 'metal'.\n", tok.c_str());
    fclose(fp);
    return 1;
  }

  ok = ok && next_line(fp, line);
  tok = ok ? first_token(line) : "none";
  in.datafile = (tok == "none") ? "" : tok;

  ok = ok && next_line(fp, line);
  tok = ok ? first_token(line) : "";
  if (tok == "lj") in.forcetype = FORCELJ;
  else if (tok == "eam") in.forcetype = FORCEEAM;
  else {
    if (me == 0) printf("Unknown forcetype option in file at line 5 ('%s'). Expecting either 'lj' or 'eam'.\n", tok.c_str());
    fclose(fp);
    return 1;
  }

  MMD_float skin = 0;
  in.epsilon = in.sigma = 1;
  in.nx = in.ny = in.nz = 0;
  in.ntimes = 0;
  in.dt = in.t_request = in.rho = in.force_cut = 0;
  in.neigh_every = 1;
  in.thermo_nstat = 0;
  if (ok && next_line(fp, line)) { MMD_float* d[2] = {&in.epsilon, &in.sigma}; scan_reals(line, d, 2); }
  if (ok && next_line(fp, line)) sscanf(line, "%d %d %d", &in.nx, &in.ny, &in.nz);
  if (ok && next_line(fp, line)) sscanf(line, "%d", &in.ntimes);
  if (ok && next_line(fp, line)) { MMD_float* d[1] = {&in.dt}; scan_reals(line, d, 1); }
  if (ok && next_line(fp, line)) { MMD_float* d[1] = {&in.t_request}; scan_reals(line, d, 1); }
  if (ok && next_line(fp, line)) { MMD_float* d[1] = {&in.rho}; scan_reals(line, d, 1); }
  if (ok && next_line(fp, line)) sscanf(line, "%d", &in.neigh_every);
  if (ok && next_line(fp, line)) { MMD_float* d[2] = {&in.force_cut, &skin}; scan_reals(line, d, 2); }
  if (ok && next_line(fp, line)) sscanf(line, "%d", &in.thermo_nstat);
  fclose(fp);

  in.neigh_cut = skin;
  in.neigh_cut += in.force_cut;  // ref/input.cpp:183, evaluated in MMD_float
  return 0;
}
