"""Frame-range decode (BASELINE.json configs[4]): a stream cut at frame boundaries — one range per GPU, or a random-access
read — gives the same PCM as the linear decode, with the reservoir resolved on the host and a two-granule halo rebuilding
the overlap / V-history seams.  Also checks the seam state against the oracle's linear decode."""
import numpy as np
import pytest

import oracle
from tools.synth import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,cfg", [("cfg5", synth.cfg5(240)), ("cfg4_lsf", synth.cfg4(19, 300)), ("cfg4_mixed", synth.cfg4(3, 240))])
@pytest.mark.parametrize("exact", [True, False])
def test_ranges_concatenate_to_linear_decode(pkg, name, cfg, exact):
    data = synth.stream(cfg)
    pb = pkg.parse_streams([data])
    g = pkg.GpuEngine(0, exact=exact)
    full = g.decode(pb.main_data, pb.main_data_len, pb.units)
    frames = pb.streams[0]["frames"]
    cuts = [0] + sorted(np.random.default_rng(5).choice(np.arange(1, frames), 7, replace=False).tolist()) + [frames]
    parts = [g.decode_frames(pb, 0, a, b) for a, b in zip(cuts[:-1], cuts[1:])]   # 8 ranges, as on 8 GPUs
    assert np.array_equal(np.concatenate(parts), full)
    if exact:
        ref, err = oracle.OracleDecoder(data).read_all()
        assert np.array_equal(full.reshape(-1), np.frombuffer(ref, np.int16))
    g.close()


def test_random_access_reads(pkg):
    """1,000 random (frame, length) reads of one stream equal the slices of its linear decode."""
    data = synth.stream(synth.cfg5(400))
    pb = pkg.parse_streams([data])
    g = pkg.GpuEngine(0)
    full = g.decode(pb.main_data, pb.main_data_len, pb.units)
    rng = np.random.default_rng(7)
    for _ in range(200):
        f0 = int(rng.integers(0, 399))
        f1 = min(400, f0 + int(rng.integers(1, 6)))
        assert np.array_equal(g.decode_frames(pb, 0, f0, f1), full[f0 * 1152:f1 * 1152])
    g.close()


def test_range_in_batch_of_streams(pkg, classic_lame):
    streams = [synth.stream(synth.cfg3(0, 50)), classic_lame, synth.stream(synth.cfg4(39, 90))]
    pb = pkg.parse_streams(streams)
    g = pkg.GpuEngine(0, exact=True)
    full = g.decode(pb.main_data, pb.main_data_len, pb.units)
    for s, (a, b) in ((0, (10, 50)), (1, (100, 385)), (1, (0, 1)), (2, (37, 90)), (2, (1, 2))):
        st = pb.streams[s]
        spf = st["pcm_bytes"] // 4 // st["frames"]
        o = st["pcm_offset"] // 4
        assert np.array_equal(g.decode_frames(pb, s, a, b), full[o + a * spf:o + b * spf]), (s, a, b)
    g.close()
