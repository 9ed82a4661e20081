"""The C++ host adapter (linevis_b200/host): LineRenderer / LineData / SettingsMap mirror of the reference interface."""
import os
import subprocess

import pytest

HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "linevis_b200", "host")


def _build():
    import linevis_b200.build as b
    b.build()
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)


def test_host_adapter_builds_and_fails_loudly_without_gpu():
    import torch
    _build()
    r = subprocess.run([os.path.join(HOST, "host_selftest")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_host_adapter_renders_on_gpu():
    _build()
    r = subprocess.run([os.path.join(HOST, "host_selftest")], capture_output=True, text=True)
    assert r.returncode == 0 and "host adapter OK (GPU)" in r.stdout, r.stdout + r.stderr
