#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (`ref/` variant of
# Mantevo/miniMD) from the sources where they lie under $MINIMD_REFERENCE
# (default /root/reference) into oracle/_ref/ (git-ignored, travels to the GPU box).
#
# Nothing is copied into the repository history: intermediates (the two-pass
# `-E` preprocess the reference's own Makefile.openmpi:72-74 needs because
# OMPFORSCHEDULE expands to a #pragma, ref/types.h:51-59) live in a temp dir
# that is removed afterwards; only the linked binaries + the reference's own
# EAM table (needed in cwd at run time, ref/force_eam.cpp:77) land in oracle/_ref/.
#
# The reference needs MPI; none is installed, so it is linked against the
# reference's OWN serial MPI stub (kokkos/MPI-Stubs/mpi.{c,h}) => 1 rank only.
#
#   oracle/_ref/miniMD_ref_f64   PRECISION=2 (double)
#   oracle/_ref/miniMD_ref_f32   PRECISION=1 (float)
set -euo pipefail
REF="${MINIMD_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/ref" ]; then
  echo "build_ref.sh: reference tree not found at $REF (expected on the build box only)"; exit 0
fi
mkdir -p "$OUT"
SRCS="ljs input integrate atom force_lj force_eam neighbor thermo comm timer output setup"
for PREC in 2 1; do
  if [ "$PREC" = 2 ]; then NAME=miniMD_ref_f64; else NAME=miniMD_ref_f32; fi
  if [ -x "$OUT/$NAME" ] && [ "$OUT/$NAME" -nt "$REF/ref/ljs.cpp" ] && [ -z "${FORCE:-}" ]; then
    echo "build_ref.sh: $NAME up to date"; continue
  fi
  TMP="$(mktemp -d /tmp/minimd_ref_build.XXXXXX)"
  FLAGS="-O3 -fopenmp -DNOCHUNK -mavx -DUSE_SIMD -DPRECISION=$PREC -I$REF/ref -I$REF/kokkos/MPI-Stubs -w"
  gcc -O -w -c "$REF/kokkos/MPI-Stubs/mpi.c" -I"$REF/kokkos/MPI-Stubs" -o "$TMP/mpi_stub.o"
  for f in $SRCS; do
    ( g++ $FLAGS -E "$REF/ref/$f.cpp" > "$TMP/$f.2.cpp" && g++ $FLAGS -c "$TMP/$f.2.cpp" -o "$TMP/$f.o" ) &
  done
  wait
  OBJS=""
  for f in $SRCS; do OBJS="$OBJS $TMP/$f.o"; done
  g++ -O3 -fopenmp $OBJS "$TMP/mpi_stub.o" -o "$OUT/$NAME"
  rm -rf "$TMP"
  echo "build_ref.sh: built $OUT/$NAME"
done
# run-time data the reference binary opens from cwd (not source code)
cp -f "$REF/ref/Cu_u6.eam" "$OUT/Cu_u6.eam"
chmod u+w "$OUT/Cu_u6.eam"
