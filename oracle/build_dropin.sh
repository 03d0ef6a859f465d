#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- the drop-in proof of INTEGRATION.md section B: the reference program with ONE of its
# translation units replaced by a file that calls the C ABI (tests/dropin/*.cpp), everything else -- main(), input, setup,
# Atom, Neighbor, Comm, Thermo, output, timer -- compiled UNMODIFIED from the sources where they lie under
# $MINIMD_REFERENCE (default /root/reference), as oracle/build_ref.sh does.  Outputs into oracle/_ref/ only:
#   miniMD_dropin_run_f64     ref/integrate.cpp replaced: Integrate::run hands the time loop to mmd_run
#   miniMD_dropin_force_f64   ref/force_lj.cpp replaced: ForceLJ::compute is one upload / mmd_force_lj_compute / download
# Both link minimd_b200/lib/libminimd_b200.so (run build.py first) with an rpath relative to oracle/_ref/.
set -euo pipefail
REF="${MINIMD_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
OUT="$HERE/_ref"
if [ ! -d "$REF/ref" ]; then
  echo "build_dropin.sh: reference tree not found at $REF (expected on the build box only)"; exit 0
fi
if [ ! -f "$ROOT/minimd_b200/lib/libminimd_b200.so" ]; then
  echo "build_dropin.sh: libminimd_b200.so missing (python -m minimd_b200.build)"; exit 1
fi
mkdir -p "$OUT"
SRCS="ljs input integrate atom force_lj force_eam neighbor thermo comm timer output setup"
FLAGS="-O3 -fopenmp -DNOCHUNK -mavx -DUSE_SIMD -DPRECISION=2 -I$REF/ref -I$REF/kokkos/MPI-Stubs -I$ROOT/include -w"
TMP="$(mktemp -d /tmp/minimd_dropin_build.XXXXXX)"
gcc -O -w -c "$REF/kokkos/MPI-Stubs/mpi.c" -I"$REF/kokkos/MPI-Stubs" -o "$TMP/mpi_stub.o"
for f in $SRCS; do
  ( g++ $FLAGS -E "$REF/ref/$f.cpp" > "$TMP/$f.2.cpp" && g++ $FLAGS -c "$TMP/$f.2.cpp" -o "$TMP/$f.o" ) &
done
for f in integrate_b200 force_lj_b200; do
  ( g++ $FLAGS -E "$ROOT/tests/dropin/$f.cpp" > "$TMP/$f.2.cpp" && g++ $FLAGS -c "$TMP/$f.2.cpp" -o "$TMP/$f.o" ) &
done
wait
link() {  # name, replaced reference unit, replacement
  OBJS=""
  for f in $SRCS; do
    if [ "$f" = "$2" ]; then OBJS="$OBJS $TMP/$3.o"; else OBJS="$OBJS $TMP/$f.o"; fi
  done
  g++ -O3 -fopenmp $OBJS "$TMP/mpi_stub.o" -L"$ROOT/minimd_b200/lib" -lminimd_b200 \
      -Wl,-rpath,'$ORIGIN/../../minimd_b200/lib' -o "$OUT/$1"
  echo "build_dropin.sh: built $OUT/$1"
}
link miniMD_dropin_run_f64 integrate integrate_b200
link miniMD_dropin_force_f64 force_lj force_lj_b200
rm -rf "$TMP"
