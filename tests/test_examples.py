"""examples/perseus_gpu_replay.c — the reference's perseustest flow (perseustest.c:93-409) on the virtual receiver and the
GPU callback, in plain C against include/perseus-gpu.h.  Built here with gcc exactly as a libperseus-sdr user would."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as O

LIBDIR = ROOT / "libperseus-sdr_b200" / "lib"


@pytest.fixture(scope="module")
def replay_bin(tmp_path_factory):
    out = tmp_path_factory.mktemp("bin") / "perseus_gpu_replay"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(ROOT / "include"),
                    str(ROOT / "examples" / "perseus_gpu_replay.c"), "-L", str(LIBDIR), "-lperseus_gpu", "-o", str(out)], check=True)
    return out


@pytest.fixture(scope="module")
def shard_bin(tmp_path_factory):
    out = tmp_path_factory.mktemp("bin") / "perseus_gpu_shard"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(ROOT / "include"),
                    str(ROOT / "examples" / "perseus_gpu_shard.c"), "-L", str(LIBDIR), "-lperseus_gpu", "-o", str(out)], check=True)
    return out


@pytest.fixture(scope="module")
def hostsink_bin(tmp_path_factory):
    out = tmp_path_factory.mktemp("bin") / "perseus_gpu_hostsink"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(ROOT / "include"),
                    str(ROOT / "examples" / "perseus_gpu_hostsink.c"), "-L", str(LIBDIR), "-lperseus_gpu", "-o", str(out)], check=True)
    return out


def run(binary, *args):
    env = dict(os.environ, LD_LIBRARY_PATH=f"{LIBDIR}:{os.environ.get('LD_LIBRARY_PATH', '')}")
    return subprocess.run([str(binary), *args], capture_output=True, text=True, env=env, timeout=120)


def test_example_builds_as_c99_and_fails_loudly_without_a_gpu(replay_bin, tmp_path):
    import torch
    r = run(replay_bin, "-h")
    assert r.returncode == 0 and "48000 95000 96000" in r.stderr
    if not torch.cuda.is_available():
        r = run(replay_bin, "-N", "4", "-o", str(tmp_path / "x"))
        assert r.returncode == 1 and "perseus_gpu_open" in r.stderr          # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("flag,mode", [((), O.MODE_I32), (("-p",), O.MODE_F32)])
def test_example_output_file_equals_reference_callbacks(replay_bin, tmp_path, coracle, flag, mode):
    out = tmp_path / "perseusdata"
    r = run(replay_bin, "-s", "96000", "-n", "6", "-b", "1024", "-N", "211", "-o", str(out), *flag)
    assert r.returncode == 0, r.stderr
    assert "Sample rate 96000 S/s, buffers of 6144 bytes" in r.stderr and "kSamples read: 216" in r.stderr   # 211*6144/6000
    wire = coracle.synth_random(211 * 6144)
    want = O.Ref().unpack(wire, mode, chunk=6144) if O.Ref.available() else coracle.unpack(wire, mode)
    assert out.read_bytes() == want.tobytes()


@pytest.mark.gpu
def test_example_streams_to_standard_output_like_perseustest_dash(replay_bin, coracle):
    """`perseustest -o -` leaves fout = stdout (perseustest.c:98,337) for a consumer on a pipe; so does the GPU path."""
    env = dict(os.environ, LD_LIBRARY_PATH=f"{LIBDIR}:{os.environ.get('LD_LIBRARY_PATH', '')}")
    r = subprocess.run([str(replay_bin), "-s", "2000000", "-n", "6", "-b", "1024", "-N", "77", "-o", "-", "-p"], capture_output=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == coracle.unpack(coracle.synth_random(77 * 6144), O.MODE_F32).tobytes()
    assert b"Bye" in r.stderr and not os.path.exists("-")


def test_hostsink_example_builds_as_c99_and_fails_loudly_without_a_gpu(hostsink_bin):
    import torch
    r = run(hostsink_bin, "-h")
    assert r.returncode == 0 and "perseus_gpu_hostsink" in r.stderr
    if not torch.cuda.is_available():
        r = run(hostsink_bin, "-N", "4")
        assert r.returncode == 1 and "perseus_gpu_open" in r.stderr           # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("args,n", [(("-N", "333"), 333), (("-s", "2000000", "-t", "1"), None)])
def test_hostsink_example_meters_exactly_what_the_reference_callbacks_would_hold(hostsink_bin, coracle, args, n):
    """examples/perseus_gpu_hostsink.c: a CPU consumer (level meter) fed by the host sink -- as fast as possible and paced like a
    2 MS/s receiver.  Its exact integer totals must equal those of the oracle's int32 unpack of the same wire stream."""
    r = run(hostsink_bin, *args)
    assert r.returncode == 0, r.stderr
    got = dict(zip(r.stdout.split()[0::2], map(int, r.stdout.split()[1::2])))
    if n is None:
        n = got["samples"] // 1024
        assert 1500 <= n <= 3000, got                      # 1 s (more on a busy box) at 1953 transfers/s
    want = coracle.unpack(coracle.synth_random(n * 6144), O.MODE_I32).view(np.int32).reshape(-1, 2).astype(np.int64)
    assert got["samples"] == n * 1024 and got["out_of_order"] == 0 and 1 <= got["blocks"] <= n
    assert (got["sum_i"], got["sum_q"], got["peak"]) == (int(want[:, 0].sum()), int(want[:, 1].sum()), int(np.abs(want).max()))


@pytest.mark.gpu
def test_example_rejects_illegal_buffer_sizes_like_the_reference(replay_bin, tmp_path):
    r = run(replay_bin, "-n", "5", "-b", "1024", "-N", "3", "-o", str(tmp_path / "x"))    # 5120 is not a multiple of 6144
    assert r.returncode == 1 and "integer multiple of 6144" in r.stderr                   # perseus-sdr.c:672
    r = run(replay_bin, "-n", "18", "-b", "1024", "-N", "3", "-o", str(tmp_path / "x"))   # 18432 > 16320
    assert r.returncode == 1 and "16320" in r.stderr                                       # perseus-sdr.c:663


@pytest.mark.gpu
def test_example_paced_run_streams_at_the_sample_rate(replay_bin, tmp_path):
    out = tmp_path / "paced"
    r = run(replay_bin, "-s", "2000000", "-t", "1", "-o", str(out), "-p")
    assert r.returncode == 0, r.stderr
    nsamples = out.stat().st_size // 8
    assert 1.5e6 < nsamples < 2.6e6 and out.stat().st_size % 8192 == 0
    wire = O.COracle().synth_random(6144 * 8)
    assert out.read_bytes()[: 8 * 8192] == O.COracle().unpack(wire, O.MODE_F32).tobytes()


def test_shard_example_builds_and_fails_loudly_without_a_gpu(shard_bin):
    import torch
    if not torch.cuda.is_available():
        r = run(shard_bin, "-n", "64")
        assert r.returncode == 1 and "no usable GPU" in r.stderr


@pytest.mark.gpu
def test_shard_example_checksum_equals_oracle_whole_recording(shard_bin, coracle):
    """One host thread, one handle per device, no collective: the per-shard checksums add up to the oracle's
    checksum of the whole recording, however many GPUs the box has."""
    import json
    import torch
    n = 4099
    wire = coracle.synth_random(n * 6144)
    for flag, mode in (((), O.MODE_I32), (("-p",), O.MODE_F32)):
        want = coracle.checksum32(coracle.unpack(wire, mode, nthreads=O.host_threads()))
        for g in sorted({1, torch.cuda.device_count()}):
            r = run(shard_bin, "-n", str(n), "-g", str(g), "-r", "3", *flag)
            assert r.returncode == 0, r.stderr
            d = json.loads(r.stdout)
            assert d["gpus"] == g and int(d["recording_checksum"], 16) == want
