"""The C-ABI library loads on a CPU-only box and exports every symbol include/pipe_b200.h declares;
the Python binding covers the same set; and the product path fails loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

from pipe_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pipe_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("pb_chain_create", "pb_chain_process", "pb_chain_process_batch_device", "pb_chain_submit",
                 "pb_chain_collect", "pb_chain_set_stage", "pb_chain_reset", "pb_chain_destroy", "pb_mix_sum_device",
                 "pb_meter_device", "pb_source_fill_device", "pb_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(abi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"libpipe_b200.so does not export {name}"


def test_binding_covers_every_declared_symbol():
    assert sorted(abi.SIGNATURES) == declared_symbols()
    assert abi.lib().pb_abi_version() == abi.ABI_VERSION


def test_struct_layouts_match_the_header():
    # pb_stage_desc: 4*int32 + pad int32 (+4 align) + double + 3 double + 2 double + pointer
    assert ctypes.sizeof(abi.StageDesc) == 24 + 8 + 24 + 16 + 8
    assert ctypes.sizeof(abi.ChainDesc) == 16 + 8 + 16 + 8


def test_no_cpu_fallback_without_a_device():
    if abi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(abi.PipeB200Error) as e:
        abi.Chain(2, [{"kind": "copy"}], buffer_frames=512, dtype=np.float64)
    assert e.value.code == abi.PB_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_invalid_descriptors_are_rejected_before_touching_the_device():
    for stages, code in (([{"kind": "resample", "up": 3, "down": 2, "taps": np.ones(6)}], abi.PB_ERR_UNSUPPORTED),
                         ([{"kind": "fir"}], abi.PB_ERR_INVALID),
                         ([{"kind": 9}], abi.PB_ERR_INVALID)):
        with pytest.raises(abi.PipeB200Error) as e:
            abi.Chain(2, stages, buffer_frames=16)
        assert e.value.code == code


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pipe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "_oracle" not in text and "pipe_oracle" not in text and "liboracle" not in text, f


def test_go_shim_uses_only_what_the_header_declares():
    # go/pipeb200.go cannot be compiled here (no Go toolchain): at least every C.<name> it refers to must exist in the header --
    # entry points, enum constants, struct types and the struct fields it fills.
    go = open(os.path.join(ROOT, "go", "pipeb200.go")).read()
    hdr = open(os.path.join(ROOT, "include", "pipe_b200.h")).read()
    used = set(re.findall(r"\bC\.((?:pb_|PB_)[A-Za-z0-9_]+)", go))
    assert {"pb_chain_create", "pb_chain_process", "pb_chain_reset", "pb_chain_sync", "pb_chain_destroy", "pb_last_error"} <= used
    for name in sorted(used):
        assert re.search(r"\b%s\b" % re.escape(name), hdr), f"go/pipeb200.go uses C.{name}, not in include/pipe_b200.h"
    body = re.search(r"typedef struct pb_chain_desc \{(.*?)\} pb_chain_desc;", hdr, flags=re.S)
    assert body, "pb_chain_desc not found in the header"
    fields = set(re.findall(r"\b([a-z_]+)\s*(?:\[[0-9]+\])?;", body.group(1)))
    lit = re.search(r"C\.pb_chain_desc\{(.*?)\n\t\t\}", go, flags=re.S)
    assert lit, "pb_chain_desc literal not found in the shim"
    for f in re.findall(r"\b([a-z_]+):", lit.group(1)):
        assert f in fields, f"go shim fills pb_chain_desc.{f}, which the header does not have"
