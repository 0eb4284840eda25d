"""cfg5 probe: mp3_decode_stream_split of one long stream on a fresh engine and on an engine that has served batches,
printing the engine's own timings per call."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402
from tools.synth import synth  # noqa: E402


def main():
    pkg = load_package()
    cores = os.cpu_count() or 1
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 413438
    data = bench.make_long_stream(synth, frames, cores)
    ix = pkg.StreamIndex(data)
    print("cores", cores, "frames", ix.frames(), "bytes", data.size)
    eng = pkg.Engine(devices=[0], host_threads=cores)
    for k in range(4):
        t = time.time()
        pcm, rc, tm = eng.decode_stream_split(ix)
        print("fresh", k, rc, round((time.time() - t) * 1e3, 1), json.dumps({a: round(b * 1e3, 1) if isinstance(b, float) else b for a, b in tm.items()}))
    buf, offs, lens = synth.batch([synth.cfg3(i, 1149) for i in range(1024)], cores)
    sb = pkg.StreamBuffer(buf, offs, lens)
    for k in range(2):
        eng.decode_batch(sb)
    for k in range(3):
        t = time.time()
        pcm, rc, tm = eng.decode_stream_split(ix)
        print("after batches", k, rc, round((time.time() - t) * 1e3, 1), json.dumps({a: round(b * 1e3, 1) if isinstance(b, float) else b for a, b in tm.items()}))
    eng.close()


if __name__ == "__main__":
    main()
