"""The Parthenon-side shim (integration/apk_b200_shim.cpp) must compile against the reference's own
headers: a syntax-only g++ pass with the include paths of the reference's Kokkos-Serial build.  This
catches pointer / extent / signature mistakes in the documented binding -- e.g. a
ParArray1D<double*> is a rank-2 view, BoundaryFlag has no `outflow`.  Needs /root/reference and the
generated config headers kept by baseline/build_reference.sh (build container only)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GEN = os.path.join(ROOT, "baseline", "_ref", "serial", "generated")


@pytest.mark.skipif(not (os.path.isdir(os.path.join(REF, "src")) and
                         os.path.isfile(os.path.join(GEN, "kokkos", "KokkosCore_config.h")) and
                         shutil.which("g++")),
                    reason="reference tree / generated Kokkos+Parthenon config headers not present")
def test_shim_compiles_against_reference_headers():
    inc = [REF + "/src", REF + "/external/parthenon/src", GEN + "/parthenon", GEN + "/kokkos",
           REF + "/external/Kokkos/core/src", REF + "/external/Kokkos/tpls/desul/include",
           REF + "/external/Kokkos/tpls/mdspan/include", REF + "/external/Kokkos/containers/src",
           REF + "/external/Kokkos/algorithms/src", REF + "/external/Kokkos/simd/src"]
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-DKOKKOS_DEPENDENCE"] + ["-I" + i for i in inc] + \
          [os.path.join(ROOT, "integration", "apk_b200_shim.cpp")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-4000:]
