#!/usr/bin/env bash
# Builds the UNMODIFIED reference (parthenon-hpc-lab/athenapk at /root/reference) with its own
# cmake build, as SURVEY.md 8(c) records.  Outputs only into baseline/_ref/ (git-ignored, but it
# travels to the GPU box).  Build trees live in /tmp.  Nothing of the reference is copied into the
# repo; only the resulting binaries are kept.
#   baseline/_ref/serial/bin/athenaPK       Kokkos Serial, g++ -O3         (~3 min, -j8)
#   baseline/_ref/cuda_sm100/bin/athenaPK   Kokkos CUDA, nvcc -arch=sm_100 (~25 min, -j6)
# usage: baseline/build_reference.sh [serial|cuda|all] [jobs]
set -euo pipefail
WHAT=${1:-all}
JOBS=${2:-6}
REF=${APK_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
BT=${APK_REF_BUILD:-/tmp/apk_ref_build}
COMMON="-DPARTHENON_DISABLE_MPI=ON -DPARTHENON_DISABLE_HDF5=ON -DPARTHENON_ENABLE_PYTHON_MODULE_CHECK=OFF \
 -DAthenaPK_ENABLE_TESTING=OFF -DCMAKE_BUILD_TYPE=Release"
mkdir -p "$OUT" "$BT"
if [[ $WHAT == serial || $WHAT == all ]]; then
  cmake -S "$REF" -B "$BT/serial" $COMMON -DKokkos_ENABLE_SERIAL=ON > "$BT/serial.configure.log" 2>&1
  make -C "$BT/serial" -j"$JOBS" athenaPK > "$BT/serial.build.log" 2>&1
  mkdir -p "$OUT/serial/bin"
  cp "$BT/serial/bin/athenaPK" "$OUT/serial/bin/athenaPK"
  # generated config headers (KokkosCore_config.h, parthenon config.hpp): what a syntax check of
  # integration/apk_b200_shim.cpp against the reference's headers needs (tests/test_integration_shim.py)
  mkdir -p "$OUT/serial/generated/kokkos/desul/atomics"
  cp -r "$BT/serial/parthenon/src/generated" "$OUT/serial/generated/parthenon"
  cp "$BT"/serial/parthenon/Kokkos/*.h* "$OUT/serial/generated/kokkos/"
  cp "$BT/serial/parthenon/Kokkos/core/src/desul/atomics/Config.hpp" "$OUT/serial/generated/kokkos/desul/atomics/"
  echo "serial: $(ls -la "$OUT/serial/bin/athenaPK")"
fi
if [[ $WHAT == cuda || $WHAT == all ]]; then
  cmake -S "$REF" -B "$BT/cuda_sm100" $COMMON -DCMAKE_CXX_COMPILER="$REF/external/Kokkos/bin/nvcc_wrapper" \
        -DKokkos_ENABLE_CUDA=ON -DKokkos_ARCH_BLACKWELL100=ON > "$BT/cuda.configure.log" 2>&1
  make -C "$BT/cuda_sm100" -j"$JOBS" athenaPK > "$BT/cuda.build.log" 2>&1
  mkdir -p "$OUT/cuda_sm100/bin"
  cp "$BT/cuda_sm100/bin/athenaPK" "$OUT/cuda_sm100/bin/athenaPK"
  strip "$OUT/cuda_sm100/bin/athenaPK" || true
  echo "cuda: $(ls -la "$OUT/cuda_sm100/bin/athenaPK")"
fi
