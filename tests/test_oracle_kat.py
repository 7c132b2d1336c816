"""The reference's known-answer tests, restated in oracle/kat_tests.cpp, must pass before the oracle is trusted."""
import subprocess

import oracle


def test_reference_kats_pin_the_oracle(oracle_lib):
    r = subprocess.run([oracle.KAT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "0 failures" in r.stdout
