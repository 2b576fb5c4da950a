"""Host-layer thread hygiene: tests/sanitize/sanitize.sh builds the C-ABI layer (handle.cu, stream_path.cu, bulk_path.cu), perseus_vrx.cpp,
copy_pool.cpp and
perseus_host.cpp with g++ against a CUDA stand-in (tests/sanitize/fake_cuda) under ThreadSanitizer and under
AddressSanitizer+UBSan and runs the multi-threaded stress driver tests/sanitize/host_stress.cpp under both."""
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(not shutil.which("/usr/bin/g++"), reason="system g++ not present")
def test_host_layer_is_clean_under_tsan_and_asan(tmp_path):
    probe = subprocess.run(["/usr/bin/g++", "-fsanitize=thread", "-x", "c++", "-", "-o", str(tmp_path / "p")], input="int main(){}",
                           text=True, capture_output=True)
    if probe.returncode != 0:
        pytest.skip("sanitizer runtimes not installed")
    r = subprocess.run([str(ROOT / "tests" / "sanitize" / "sanitize.sh"), str(tmp_path)], capture_output=True, text=True, timeout=600)
    logs = "".join((tmp_path / f"r2_sanitizer_host_{s}.txt").read_text() for s in ("tsan", "asan"))
    assert r.returncode == 0, logs[-4000:]
    assert logs.count("host_stress: all scenarios passed") == 4      # tsan, asan x membarrier, fence fallback
    assert "ThreadSanitizer" not in logs.replace("-fsanitize=thread", "") and "AddressSanitizer" not in logs and "runtime error" not in logs
