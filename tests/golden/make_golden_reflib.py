#!/usr/bin/env python
"""Regenerates tests/golden/reflib.json by EXECUTING the reference's own library
(/root/reference/perseus-sdr.c, perseusfx2.c, perseus-in.c, perseuserr.c, compiled unmodified into
oracle/_ref/libperseus_sdr_ref.so over the fake libusb of oracle/fakeusb.c).

Run in the dev container, where /root/reference is mounted:
    python tests/golden/make_golden_reflib.py

What is frozen (each an output of reference code, not of a restatement):
  * rate_choice      perseus_set_sampling_rate(descr, x) -> which bitstream reached the FPGA: the result of the
                     reference's static getFpgaFile (perseus-sdr.c:776-811) for every table rate, every midpoint
                     between neighbours -1/0/+1, and the out-of-range ends
  * sampling_rates   perseus_get_sampling_rates (perseus-sdr.c:814-832), full and with short buffers
  * start_codes      perseus_start_async_input's return code and message (perseus-sdr.c:662-680) for a list of
                     buffer sizes, with 512-, 510- and 64-byte endpoints
  * bitstreams       name / rate / size / FNV-1a-64 of the ten *.rbs files (what the fake FPGA recognises)
  * firmware         record count, byte count and hash of the FX2 firmware the reference downloads
                     (perseus24v41_512.c via perseusfx2.c:164-202), parsed from the reference's table
"""
import ctypes as C
import json
import re
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

REFERENCE = Path("/root/reference")
RATES = [48000, 95000, 96000, 125000, 192000, 250000, 500000, 1000000, 1600000, 2000000]
SIZES = [6, 510, 1020, 1024, 3072, 6143, 6144, 6145, 6150, 12288, 15300, 16320, 16321, 18432, 20000]


def rate_probes():
    xs = set(RATES) | {-5, 0, 1, 47999, 48001, 2000001, 5_000_000}
    for a, b in zip(RATES, RATES[1:]):
        xs |= {(a + b) // 2 + d for d in (-1, 0, 1)} | {a + 1, b - 1}
    return sorted(xs)


def firmware_facts():
    text = (REFERENCE / "perseus24v41_512.c").read_text()
    recs = re.findall(r"\{\s*0x([0-9A-Fa-f]{2}),\s*0x([0-9A-Fa-f]{4}),\s*(\d+),\s*\{([^}]*)\}\s*\}", text)
    h, nbytes = 0xCBF29CE484222325, 0
    for _, addr, n, data in recs:
        addr, n = int(addr, 16), int(n)
        payload = bytes(int(x, 16) for x in data.split(",") if x.strip())
        assert len(payload) == n
        for b in bytes((addr & 0xFF, addr >> 8, n & 0xFF, n >> 8)) + payload:      # as oracle/fakeusb.c hashes each RAM write
            h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
        nbytes += n
    return {"source": "perseus24v41_512.c", "records": len(recs), "bytes": nbytes, "fnv1a64": f"{h:016x}"}


def main() -> None:
    rl = O.RefLib()
    noop = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)(lambda b, n, e: 0)
    cb = C.cast(noop, C.c_void_p)
    out = {"generated_by": "tests/golden/make_golden_reflib.py (executes oracle/_ref/libperseus_sdr_ref.so)"}

    with tempfile.TemporaryDirectory() as td:
        subprocess.run([sys.executable, str(ROOT / "oracle" / "gen_fpga_data.py"), str(REFERENCE), f"{td}/f.c", f"{td}/b.json"], check=True)
        out["bitstreams"] = json.loads(Path(f"{td}/b.json").read_text())["bitstreams"]
    out["firmware"] = firmware_facts()

    with rl.session(rate=0) as d:
        choice = {}
        for x in rate_probes():
            rc = rl.L.perseus_set_sampling_rate(d, x)
            assert rc == 0, (x, rl.errorstr())
            choice[str(x)] = rl.state()["fpga_rate"]
        out["rate_choice"] = choice
        buf = (C.c_int * 16)()
        rates = {}
        for size in (16, 10, 9, 4, 1, 0):
            rc = rl.L.perseus_get_sampling_rates(d, buf, size)
            rates[str(size)] = {"rc": rc, "values": list(buf)[:size], "message": rl.errorstr() if rc < 0 else ""}
        out["sampling_rates"] = rates

    codes = {}
    for ep in (512, 510, 64):
        with rl.session(rate=2_000_000, ep_max_packet=ep, limit=1) as d:
            for size in SIZES:
                rc = rl.L.perseus_start_async_input(d, size, cb, None)
                codes[f"{ep}:{size}"] = {"rc": rc, "message": rl.errorstr() if rc < 0 else ""}
                if rc == 0:
                    again = rl.L.perseus_start_async_input(d, size, cb, None)
                    codes[f"{ep}:{size}"]["second_start_rc"] = again
                    assert rl.L.perseus_stop_async_input(d) == 0
            rc = rl.L.perseus_stop_async_input(d)
            codes[f"{ep}:stop_when_not_started"] = {"rc": rc, "message": rl.errorstr()}
    out["start_codes"] = codes

    # the reference's own APPLICATION (examples/perseustest.c + fifo.c, unmodified: oracle/_ref/perseustest_ref) writing its
    # output file from the synthetic receiver: the `perseustest -o` format, end to end
    import os
    from oracle import oracle as Om
    co = Om.COracle()
    app = {"args": "-a -d 0 -s 250000 -n 6 -b 1024 -t 1 -o FILE [-p]", "limit": 200, "seed": 77}
    with tempfile.TemporaryDirectory() as td:
        for name, flag in (("int32", []), ("float", ["-p"])):
            out_file = f"{td}/{name}.bin"
            env = dict(os.environ, FAKEUSB_AUTOPLUG="1", FAKEUSB_LIMIT=str(app["limit"]), FAKEUSB_SEED=str(app["seed"]))
            subprocess.run([str(ROOT / "oracle" / "_ref" / "perseustest_ref"), "-a", "-d", "0", "-s", "250000", "-n", "6", "-b", "1024", "-t", "1",
                            "-o", out_file, *flag], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=60)
            data = Path(out_file).read_bytes()
            app[f"{name}_nbytes"] = len(data)
            app[f"{name}_fnv1a64"] = f"{co.fnv1a64(data):016x}"
    out["perseustest_app"] = app

    # examples/simple.c (the reference's `make check` program, its own copy of the int32 callback, simple.c:33-61): fixed 96 kS/s,
    # 6 x 1024-byte transfers, 10 s, writes ./perseusdata.bin
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, FAKEUSB_AUTOPLUG="1", FAKEUSB_LIMIT="200", FAKEUSB_SEED="77")
        subprocess.run([str(ROOT / "oracle" / "_ref" / "simple_ref")], cwd=td, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=60)
        data = Path(f"{td}/perseusdata.bin").read_bytes()
        out["simple_app"] = {"limit": 200, "seed": 77, "int32_nbytes": len(data), "int32_fnv1a64": f"{co.fnv1a64(data):016x}"}

    (HERE / "reflib.json").write_text(json.dumps(out, indent=1) + "\n")
    print(f"wrote {HERE / 'reflib.json'}: {len(out['rate_choice'])} rate probes, {len(codes)} start codes")


if __name__ == "__main__":
    main()
