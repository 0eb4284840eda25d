"""Small decode workload for compute-sanitizer: fixtures prefix + wild + fuzz streams through the host and device APIs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from tools.synth import synth  # noqa: E402

pkg = load_package()
streams = [open(os.path.join(ROOT, "tests/golden/fixtures/classic_lame.mp3"), "rb").read()[:60000],
           open(os.path.join(ROOT, "tests/golden/fixtures/mpeg2.mp3"), "rb").read()[:40000]]
streams += [synth.stream(synth.cfg4(i, 40)) for i in (3, 19, 39)]
streams += [synth.stream(synth.wild(i)) for i in range(8)] + [synth.stream(synth.fuzz(i)) for i in range(12)]
pb = pkg.parse_streams(streams)
for wave in (0, 7):
    g = pkg.GpuEngine(0, wave_granules=wave, keep_intermediates=True)
    pcm = g.decode(pb.main_data, pb.main_data_len, pb.units)
    print("wave", wave, "granules", pb.n_granules, "checksum", int(pcm.astype(np.int64).sum()))
    g.close()
eng = pkg.Engine(0, chunk_frames=5)
d = eng.new_decoder(streams[0])
d.read_all()
d.seek_to_time(10**9)
d.read(5000)
print("decoder ok")
