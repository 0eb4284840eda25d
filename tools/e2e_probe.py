"""End-to-end probe: mp3_decode_batch (raw .mp3 bytes -> PCM in pinned host memory) on N synthetic streams, a few calls,
printing the per-call wall time, the D2H rate and the plain-copy ceiling.  Used to tune the host pipeline (MP3HOST_CHUNKS)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from tools.synth import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=1149)
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--calls", type=int, default=6)
    ap.add_argument("--devices", type=int, default=1)
    args = ap.parse_args()
    pkg = load_package()
    cores = os.cpu_count() or 1
    gen = {"cfg3": synth.cfg3, "cfg4": synth.cfg4}[args.workload]
    buf, offs, lens = synth.batch([gen(i, args.frames) for i in range(args.streams)], cores)
    sb = pkg.StreamBuffer(buf, offs * args.devices, lens * args.devices)
    eng = pkg.Engine(devices=list(range(args.devices)), host_threads=cores)
    ceil = eng.measure_d2h_ceiling(2 << 30, 3)
    eng.decode_batch(sb)
    eng.decode_batch(sb)
    best = None
    for _ in range(args.calls):
        res, pcm, tm = eng.decode_batch(sb)
        nbytes = sum(r["pcm_bytes"] for r in res)
        row = {"total_ms": tm["total_s"] * 1e3, "parse_ms": tm["parse_s"] * 1e3, "gather_ms": tm["gather_s"] * 1e3,
               "device_ms": tm["device_s"] * 1e3, "d2h_gbs": nbytes / tm["total_s"] / 1e9, "gsamples_s": nbytes / 4 / tm["total_s"] / 1e9}
        if best is None or row["total_ms"] < best["total_ms"]:
            best = row
    best["ceiling_gbs"] = ceil["aggregate_gbs"]
    best["frac_of_ceiling"] = best["d2h_gbs"] / ceil["aggregate_gbs"]
    best["chunks_env"] = os.environ.get("MP3HOST_CHUNKS", "")
    print(json.dumps(best))
    eng.close()


if __name__ == "__main__":
    main()
