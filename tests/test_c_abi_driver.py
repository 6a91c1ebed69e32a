"""The C-ABI driven from plain C (gcc, no ctypes) in the call order of the Fortran ISO_C_BINDING shim
(fortran/rpbmd_iso_c.f90), with static_asserts of the rpb_config layout the shim's bind(C) type implies."""
import os
import re
import subprocess

import pytest

from reactive_pb_nn_md_b200 import _binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_driver.c")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("abi") / "abi_driver")
    subprocess.check_call(["gcc", "-std=gnu11", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-ldl", "-lm"])
    return exe


def test_plain_c_driver_on_the_oracle(driver, oracle_lib):
    out = subprocess.run([driver, oracle_lib.path], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi driver: OK" in out.stdout and "oracle-cpu" in out.stdout


def test_plain_c_driver_resolves_every_symbol_of_the_cuda_library(driver):
    out = subprocess.run([driver, _binding.CUDA_LIB_PATH, "symbols-only"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cuda-sm100a" in out.stdout


@pytest.mark.gpu
def test_plain_c_driver_on_the_cuda_library(driver):
    out = subprocess.run([driver, _binding.CUDA_LIB_PATH], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi driver: OK" in out.stdout


def test_fortran_shim_binds_only_exported_names():
    """every bind(C, name=...) of fortran/rpbmd_iso_c.f90 is declared in include/rpbmd.h, and the bind(C) mirror of
    rpb_config lists the header's fields in the header's order"""
    f90 = open(os.path.join(ROOT, "fortran", "rpbmd_iso_c.f90")).read()
    hdr = open(os.path.join(ROOT, "include", "rpbmd.h")).read()
    names = set(re.findall(r'bind\(C,\s*name="(rpb_[a-z0-9_]+)"\)', f90))
    assert len(names) >= 18
    for n in names:
        assert re.search(r"\b%s\s*\(" % n, hdr), n
    for must in ("rpb_set_molecule_types", "rpb_set_evb", "rpb_peer_export", "rpb_peer_import", "rpb_upload_state", "rpb_download_state"):
        assert must in names
    # field order of rpb_config
    body = hdr[hdr.index("typedef struct {"):hdr.index("} rpb_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = [re.sub(r"\[.*\]", "", x.strip()) for decl in re.findall(r"(?:int|double)\s+([^;]+);", body) for x in decl.split(",")]
    t = f90[f90.index("\n  type, bind(C) :: rpb_config") + 1:f90.index("end type rpb_config")].split("\n")[1:]
    f_fields = [re.sub(r"\(.*\)", "", x.strip()) for line in t if "::" in line
                for x in line.split("!")[0].split("::")[1].split(",")]
    assert c_fields == f_fields, (c_fields, f_fields)
