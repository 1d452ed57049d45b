"""A plain C caller (tests/c/dropin_check.c) drives the host-built reference and this repository's alias
library through the SAME compiled code and the same struct definitions; results must be identical."""
import os
import subprocess

import pytest

from conftest import requires_ref
from oracle import ref as oracle
from trik_media_sensors_dsp_b200 import build

pytestmark = [pytest.mark.gpu, requires_ref]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("c") / "dropin_check")
    subprocess.run(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "c", "dropin_check.c"), "-ldl"], check=True)
    return exe


@pytest.mark.parametrize("kind", ["wl", "wo", "ol", "om", "oo"])
def test_c_caller_gets_identical_results(checker, kind):
    res = subprocess.run([checker, os.path.join(oracle.REF_DIR, "libtrikref_%s.so" % kind), build.alias_path(kind), kind],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "identical to the reference" in res.stdout
