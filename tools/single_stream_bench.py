"""BASELINE.json configs[0] / configs[1]: one stream through the drop-in Decoder (NewDecoder + io.ReadAll, bench_test.go:40-55)
on one GPU, next to the oracle on one host core.  Prints one JSON line per fixture."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402
import oracle  # noqa: E402  (checker / CPU baseline only)

pkg = load_package()
for name in ("classic_lame", "mpeg2"):
    data = open(os.path.join(ROOT, "tests", "golden", "fixtures", name + ".mp3"), "rb").read()
    res = {}
    for chunk in (64, 256, 4096):
        eng = pkg.Engine(0, chunk_frames=chunk)
        eng.new_decoder(data).read_all()  # warm-up (workspace, tables)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            pcm, err = eng.new_decoder(data).read_all()
            best = min(best, time.perf_counter() - t0)
        res[f"gpu_decoder_chunk{chunk}_ms"] = best * 1e3
        eng.close()
    t0 = time.perf_counter()
    ref, _ = oracle.OracleDecoder(data).read_all()
    cpu = time.perf_counter() - t0
    samples = len(ref) // 4
    diff = np.abs(np.frombuffer(pcm, np.int16).astype(np.int32) - np.frombuffer(ref, np.int16).astype(np.int32))
    res.update({"stream": name, "stereo_samples": samples, "cpu_oracle_1core_ms": cpu * 1e3,
                "gpu_msamples_per_s_chunk4096": samples / res["gpu_decoder_chunk4096_ms"] / 1e3,
                "cpu_msamples_per_s": samples / cpu / 1e6, "max_abs_diff_lsb": int(diff.max())})
    print(json.dumps(res))
