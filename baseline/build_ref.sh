#!/usr/bin/env bash
# baseline/build_ref.sh — build the TRUE reference (jchristopherson/nonlin, Fortran) for a CPU baseline, on a machine
# that has what this image lacks: a Fortran compiler, CMake >= 3.24, BLAS/LAPACK and the `linalg` package sources.
#
# Not run by the tests, by bench.py or by __graft_entry__: in this image there is no gfortran (so `bench.py --impl
# reference` and `cpu_baseline` time the C++ port under oracle/, and say so).  The script only documents, in
# executable form, how the reference's own CPU path would be produced and timed next to the engine.
#
#   NONLIN_SRC=/path/to/nonlin  LINALG_SRC=/path/to/linalg  baseline/build_ref.sh [build-dir]
#
# Outputs <build-dir>/install/{lib,include}: link a small driver against libnonlin + liblinalg + LAPACK that loops
# `call solver%solve(...)` over the systems of nonlin_b200/workloads.py (one system per OpenMP thread) to obtain
# "converged systems/s" for the reference itself.
set -euo pipefail
NONLIN_SRC="${NONLIN_SRC:-/root/reference}"
LINALG_SRC="${LINALG_SRC:-}"
OUT="${1:-baseline/_ref_build}"

command -v gfortran >/dev/null || { echo "gfortran not found: the reference cannot be built here" >&2; exit 3; }
command -v cmake >/dev/null || { echo "cmake not found" >&2; exit 3; }
[ -d "$NONLIN_SRC/src" ] || { echo "NONLIN_SRC=$NONLIN_SRC has no src/" >&2; exit 3; }
[ -n "$LINALG_SRC" ] && [ -d "$LINALG_SRC" ] || {
    echo "LINALG_SRC must point at a checkout of github.com/jchristopherson/linalg (>= 2.0; the reference does not pin a tag)" >&2
    exit 3
}

mkdir -p "$OUT"
# linalg first (it finds or builds BLAS/LAPACK and ferror itself), then nonlin against it.  Default gfortran flags:
# -O2, no -ffast-math, no FMA contraction on x86-64 - the arithmetic the parity build of the engine models.
cmake -S "$LINALG_SRC" -B "$OUT/linalg" -DCMAKE_BUILD_TYPE=Release -DCMAKE_INSTALL_PREFIX="$PWD/$OUT/install"
cmake --build "$OUT/linalg" -j && cmake --install "$OUT/linalg"
cmake -S "$NONLIN_SRC" -B "$OUT/nonlin" -DCMAKE_BUILD_TYPE=Release -DCMAKE_PREFIX_PATH="$PWD/$OUT/install" \
      -DCMAKE_INSTALL_PREFIX="$PWD/$OUT/install" -DBUILD_TESTING=TRUE
cmake --build "$OUT/nonlin" -j && cmake --install "$OUT/nonlin"
# the reference's own 35 checks, including the KATs the oracle is pinned to
ctest --test-dir "$OUT/nonlin" --output-on-failure
echo "reference installed under $OUT/install"
