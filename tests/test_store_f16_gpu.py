"""VITCAP_STORE=fp16: the fast mode with EVERY 16-bit operand stored as an IEEE half (libvitcap_b200_f16.so, the same sources
built with -DVC_STORE_F16; DESIGN.md section 4a''). The storage type is process-wide, so the checks (tests/store_f16_check.py)
run in a child process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_half_storage_mode_against_its_oracle_and_the_fp32_reference():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VITCAP_STORE="fp16", PYTHONPATH=root)
    env.pop("VITCAP_LIB", None)
    r = subprocess.run([sys.executable, "-m", "tests.store_f16_check"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "STORE-F16-OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
