"""The C-ABI library loads without a GPU and exports every symbol include/trik_b200.h declares;
struct layouts agree between the C headers, the ctypes mirror and the oracle."""
import ctypes as C
import os
import re
import subprocess

import pytest

from trik_media_sensors_dsp_b200 import build, sensors, xdm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "trik_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    funcs = set(re.findall(r"\b(trikb200_\w+)\s*\(", text))
    data = set(re.findall(r"extern\s+\w+\s+(TRIKB200_\w+)\s*;", text))
    return funcs, data


def exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], check=True, capture_output=True, text=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_every_declared_symbol_is_exported():
    funcs, data = declared_symbols()
    assert len(funcs) >= 15 and len(data) == 10
    syms = exported(build.LIB)
    assert not (funcs - syms), "missing functions: %s" % sorted(funcs - syms)
    assert not (data - syms), "missing tables: %s" % sorted(data - syms)


def test_alias_libraries_export_the_reference_names():
    for kind in build.KINDS:
        syms = exported(build.alias_path(kind))
        assert {"TRIK_VIDTRANSCODE_CV_FXNS", "TRIK_VIDTRANSCODE_CV_IALG"} <= syms


def test_library_loads_and_sizes_match():
    lib = sensors.lib()
    for kind in range(5):
        assert lib.trikb200_sizeofInArgsAlg(kind) == C.sizeof(xdm.IN_ARGS_ALG[kind])
        assert lib.trikb200_sizeofOutArgsAlg(kind) == C.sizeof(xdm.OUT_ARGS_ALG[kind])
        assert lib.trikb200_sizeofInArgs(kind) == C.sizeof(xdm.in_args_type(kind))
        assert lib.trikb200_sizeofOutArgs(kind) == C.sizeof(xdm.out_args_type(kind))
    assert C.sizeof(xdm.TargetOutArgsAlg) == 16 and C.sizeof(xdm.ObjOutArgsAlg) == 36 and C.sizeof(xdm.MxnOutArgsAlg) == 400
    assert C.sizeof(xdm.RangeInArgsAlg) == 10 and C.sizeof(xdm.ObjInArgsAlg) == 12


def test_tables_are_wired():
    """alloc/free/init/process/control are set, the optional IALG slots are NULL
    (<sensor>/src/vidtranscode_cv_fxns.c:20-29)."""
    lib = sensors.lib()
    for kind in range(5):
        fx = lib.trikb200_fxns(kind).contents
        assert fx.ialg.algAlloc and fx.ialg.algFree and fx.ialg.algInit and fx.process and fx.control
        assert not fx.ialg.algActivate and not fx.ialg.algControl and not fx.ialg.algDeactivate
        assert not fx.ialg.algMoved and not fx.ialg.algNumAlloc
        tab = (xdm.IALG_MemRec * 4)()
        assert fx.ialg.algAlloc(None, None, tab) == 2           # two records (fxns.c:85-102)
        assert tab[0].space == xdm.IALG_EXTERNAL and tab[0].attrs == xdm.IALG_PERSIST
        assert tab[1].size == 0x1000 and tab[1].space == xdm.IALG_DARAM0 and tab[1].attrs == xdm.IALG_PERSIST
    assert not lib.trikb200_fxns(7)


def test_oracle_and_ctypes_layouts_agree():
    from oracle import ref
    for name, kind in xdm.KIND_OF.items():
        assert C.sizeof(ref.IN_ARGS[name]) == C.sizeof(xdm.IN_ARGS_ALG[kind])
        assert C.sizeof(ref.OUT_ARGS[name]) == C.sizeof(xdm.OUT_ARGS_ALG[kind])


def test_no_gpu_fails_loudly():
    """Without a CUDA device initObj must FAIL (no CPU fallback), with an error message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    c = sensors.Codec("wl")
    assert c.init_result == xdm.IALG_EFAIL
    assert sensors.last_error() != ""
    with pytest.raises(sensors.TrikB200Error):
        sensors.open_sensor("wl", 320, 240)
