#!/usr/bin/env python
"""Regenerates tests/golden/* from the reference's OWN callbacks (oracle/_ref).

Run in the dev container, where /root/reference is mounted:
    python tests/golden/make_golden.py
Everything written here is the byte-for-byte output of
/root/reference/examples/perseustest.c's user_data_callback_c_u / _c_f (compiled
verbatim by oracle/Makefile) on inputs defined in this script; the GPU box, which has
no /root/reference, checks the CUDA path and the restated oracle against these files.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from oracle import oracle as O  # noqa: E402


def le24(v: int) -> bytes:
    return bytes((v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF))


def main() -> None:
    ref, co = O.Ref(), O.COracle()
    meta = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref (reference callbacks, verbatim)",
            "build": ref.build_info()}

    # 1. known answers: every interesting 24-bit code, as I with Q = complement
    codes = [0x000000, 0x000001, 0x7FFFFF, 0x800000, 0xFFFFFF, 0x123456, 0x800001, 0x7FFFFE,
             0x0000FF, 0x00FF00, 0xFF0000, 0x00007F, 0x000080, 0x7F0000, 0x808080, 0x010203, 0xFEDCBA, 0x400000,
             0xC00000, 0x3FFFFF]
    wire = b"".join(le24(c) + le24(c ^ 0xFFFFFF) for c in codes)
    i32 = ref.unpack(np.frombuffer(wire, np.uint8), O.MODE_I32, chunk=len(wire))
    f32 = ref.unpack(np.frombuffer(wire, np.uint8), O.MODE_F32, chunk=len(wire))
    kat = []
    for k, c in enumerate(codes):
        kat.append({"code24": f"{c:06x}", "wire_hex": le24(c).hex(),
                    "int32_hex": f"{int(i32[k, 0]) & 0xFFFFFFFF:08x}",
                    "float_bits_hex": f"{int(f32[k, 0].view(np.uint32)):08x}",
                    "float_repr": repr(float(f32[k, 0])),
                    "q_code24": f"{c ^ 0xFFFFFF:06x}",
                    "q_int32_hex": f"{int(i32[k, 1]) & 0xFFFFFFFF:08x}",
                    "q_float_bits_hex": f"{int(f32[k, 1].view(np.uint32)):08x}"})
    (HERE / "kat.json").write_text(json.dumps({"meta": meta, "vectors": kat}, indent=1) + "\n")

    # 2. exhaustive ramp: all 2^24 I codes, fed in 6144-byte transfers; FNV-1a-64 of what was fwritten
    ramp = co.synth_ramp(1 << 24)
    hashes = {"pattern": "I=v, Q=(uint32(v*2654435761))>>8 for v in 0..2^24-1; 6144-byte transfers",
              "input_fnv1a64": f"{co.fnv1a64(ramp):016x}"}
    for name, mode in (("int32", O.MODE_I32), ("float", O.MODE_F32)):
        out = ref.unpack(ramp, mode, chunk=6144)
        hashes[f"{name}_fnv1a64"] = f"{co.fnv1a64(out):016x}"
        hashes[f"{name}_checksum32"] = f"{co.checksum32(out):016x}"
        hashes[f"{name}_nbytes"] = int(out.nbytes)
    (HERE / "exhaustive.json").write_text(json.dumps({"meta": meta, **hashes}, indent=1) + "\n")

    # 3. small binary fixtures: one default transfer (6144 B, perseustest.c:100-102), one legacy
    #    510-byte transfer (perseus-sdr.c:675), one ragged size (1000 B -> 166 samples, 4 bytes ignored)
    rnd = co.synth_random(6144 + 510 + 1000, seed=O.SYNTH_SEED)
    parts = {"xfer6144": rnd[:6144], "xfer510": rnd[6144:6654], "ragged1000": rnd[6654:]}
    index = {}
    for name, b in parts.items():
        (HERE / f"{name}.in.bin").write_bytes(b.tobytes())
        o_i = ref.unpack(b, O.MODE_I32, chunk=b.size)
        o_f = ref.unpack(b, O.MODE_F32, chunk=b.size)
        (HERE / f"{name}.i32.bin").write_bytes(o_i.tobytes())
        (HERE / f"{name}.f32.bin").write_bytes(o_f.tobytes())
        index[name] = {"in_bytes": int(b.size), "samples": int(o_i.shape[0])}
    # 4. hashes of a longer seeded stream (96 transfers ~ 1 s at 96 kS/s)
    big = co.synth_random(96 * 6144, seed=O.SYNTH_SEED + 1)
    index["stream96"] = {"seed": f"{O.SYNTH_SEED + 1:#x}", "in_bytes": int(big.size),
                         "input_fnv1a64": f"{co.fnv1a64(big):016x}",
                         "int32_fnv1a64": f"{co.fnv1a64(ref.unpack(big, O.MODE_I32)):016x}",
                         "float_fnv1a64": f"{co.fnv1a64(ref.unpack(big, O.MODE_F32)):016x}"}
    (HERE / "fixtures.json").write_text(json.dumps({"meta": meta, **index}, indent=1) + "\n")
    print("golden written:", sorted(p.name for p in HERE.iterdir()))


if __name__ == "__main__":
    main()
