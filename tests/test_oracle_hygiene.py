"""Oracle hygiene (layer 5): the C restatement built with
-fsanitize=address,undefined runs every sensor over noise / flat / structured frames at six sizes with
extreme arguments; any out-of-bounds access or undefined arithmetic aborts."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_is_clean_under_asan_ubsan():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "asan"], check=True)
    res = subprocess.run([os.path.join(ROOT, "oracle", "_asan_check")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "oracle hygiene ok" in res.stdout
