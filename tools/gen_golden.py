#!/usr/bin/env python3
"""Regenerates tests/golden/*.json (dev-time; needs /root/reference for the fuzz inputs).

  fuzz_crashers.json  the 10 historical crasher inputs of the reference's fuzzing_test.go:22-107, as hex
                      (Go string literals decoded: \\xNN = one byte, \\uNNNN = UTF-8 of the code point)
  oracle_pcm.json     sha256 / length / sample rate of the ORACLE's PCM for the two audio fixtures and a fixed set of
                      synthetic streams (tools/synth), so that the exact-build GPU tests also compare against
                      committed values and the oracle itself is pinned against accidental edits
"""
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def go_string_bytes(lit: str) -> bytes:
    out = bytearray()
    i = 0
    while i < len(lit):
        c = lit[i]
        if c == "\\":
            n = lit[i + 1]
            if n == "x":
                out.append(int(lit[i + 2:i + 4], 16)); i += 4
            elif n == "u":
                out += chr(int(lit[i + 2:i + 6], 16)).encode("utf-8"); i += 6
            elif n == "t":
                out.append(9); i += 2
            elif n == "n":
                out.append(10); i += 2
            elif n == "\\":
                out.append(92); i += 2
            elif n == '"':
                out.append(34); i += 2
            else:
                raise ValueError(f"escape \\{n}")
        else:
            out += c.encode("utf-8"); i += 1
    return bytes(out)


def fuzz_inputs():
    src = open(os.path.join(REF, "fuzzing_test.go"), encoding="utf-8").read()
    body = src[src.index("inputs := []string{"):src.index("for _, input := range inputs")]
    inputs, cur = [], None
    for line in body.splitlines():
        line = line.strip()
        if line.startswith("//") or not line.startswith('"'):
            continue
        m = re.match(r'"((?:[^"\\]|\\.)*)"\s*(\+|,)', line)
        piece = go_string_bytes(m.group(1))
        cur = piece if cur is None else cur + piece
        if m.group(2) == ",":
            inputs.append(cur); cur = None
    return inputs


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    if os.path.isdir(REF):
        ins = fuzz_inputs()
        assert len(ins) == 10, len(ins)
        with open(os.path.join(gold, "fuzz_crashers.json"), "w") as f:
            json.dump({"source": "fuzzing_test.go:22-107", "inputs_hex": [b.hex() for b in ins]}, f, indent=1)
    import oracle
    from tools.synth import synth
    cases = {}
    for name in ("classic_lame", "mpeg2"):
        with open(os.path.join(gold, "fixtures", name + ".mp3"), "rb") as f:
            cases[name] = f.read()
    for name, cfg in golden_synth_cases(synth):
        cases[name] = synth.stream(cfg)
    out = {}
    for name, data in cases.items():
        d = oracle.OracleDecoder(data)
        pcm, err = d.read_all() if d.ok() else (b"", d.open_err)
        out[name] = {"input_sha256": hashlib.sha256(data).hexdigest(), "pcm_sha256": hashlib.sha256(pcm).hexdigest(),
                     "pcm_bytes": len(pcm), "err": err, "sample_rate": d.sample_rate() if d.ok() else 0}
    with open(os.path.join(gold, "oracle_pcm.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", len(out), "golden PCM digests")


def golden_synth_cases(synth):
    cs = [("cfg3_0", synth.cfg3(0, 80)), ("cfg4_3", synth.cfg4(3, 80)), ("cfg4_19_lsf", synth.cfg4(19, 80)),
          ("cfg4_39_lsf_mono", synth.cfg4(39, 80)), ("cfg5", synth.cfg5(60))]
    cs += [(f"wild{i}", synth.wild(i)) for i in range(6)]
    cs += [(f"fuzz{i}", synth.fuzz(i)) for i in range(6)]
    return cs


if __name__ == "__main__":
    main()
