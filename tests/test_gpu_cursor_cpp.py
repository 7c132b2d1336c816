"""C++ KeyCursor mirror (opengemini_b200/host) — built on CPU, parity-checked on the GPU by tests/cpp/cursor_test.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "cursor_test")


def _build():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "opengemini_b200", "libogpu.so")) or not os.path.exists(BIN):
        g.build()
    assert os.path.exists(BIN)


def test_cursor_host_lib_builds_and_fails_loudly_without_gpu():
    """No CPU fallback above the ABI either: without a device the C++ cursor test aborts at og_init (exit 2)."""
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    p = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert p.returncode == 2, p.stdout + p.stderr
    assert "og_init failed" in p.stdout


@pytest.mark.gpu
def test_cursor_cpp_parity_against_oracle():
    _build()
    p = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(p.stdout[-4000:])
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-2000:]
    assert " 0 failures" in p.stdout
