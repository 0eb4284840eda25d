"""Write N synthetic streams of a workload preset as .mp3 files (input for go-mp3_b200/go/bench_batch_test.go)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tools.synth import synth  # noqa: E402

if __name__ == "__main__":
    out, n = sys.argv[1], int(sys.argv[2])
    kind = sys.argv[3] if len(sys.argv) > 3 else "cfg3"
    os.makedirs(out, exist_ok=True)
    gen = {"cfg3": synth.cfg3, "cfg4": synth.cfg4}[kind]
    for i in range(n):
        with open(os.path.join(out, f"{kind}_{i:05d}.mp3"), "wb") as f:
            f.write(synth.stream(gen(i)))
