"""CPU-only checks of the drop-in boundary: the shared library loads and exports every symbol
that include/*.h declare, struct mirrors have the C layout, and the product package does not
reference the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def apk():
    import athenapk_b200 as m
    if not os.path.exists(m.LIB_PATH):
        m.build()
    return m


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(apk_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(apk):
    L = apk.lib()
    names = declared("apk_b200.h") + declared("apk_b200_sim.h")
    assert len(names) > 40
    for n in names:
        assert hasattr(L, n), n
    assert set(declared("apk_b200.h")) == set(apk.ABI_SYMBOLS)
    assert set(declared("apk_b200_sim.h")) == set(apk.SIM_SYMBOLS)
    assert L.apk_abi_version() == 5


def test_struct_layouts_match_c(apk, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "apk_b200_sim.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(apk_block_t),sizeof(apk_pack_desc_t),sizeof(apk_sim_config_t),'
                   'sizeof(apk_halo_item_t));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes[0] == C.sizeof(apk.Block)
    assert sizes[1] == C.sizeof(apk.PackDesc)
    assert sizes[2] == C.sizeof(apk.SimConfig)
    assert sizes[3] == 40


def test_headers_are_plain_c(tmp_path):
    """The boundary must be consumable from C (cgo/JNI/ctypes-style FFI): no C++ in the headers."""
    src = tmp_path / "c.c"
    src.write_text('#include "apk_b200.h"\n#include "apk_b200_sim.h"\nint main(void){return 0;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Werror", "-I",
                           os.path.join(ROOT, "include"), "-c", str(src), "-o",
                           str(tmp_path / "c.o")])


def test_product_never_touches_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "athenapk_b200")):
        if os.sep + "build" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"oracle_lib|apk_oracle|apko_|/oracle/", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_no_gpu_means_loud_failure(apk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        apk.HydroSim(nx=(16, 16, 16))
