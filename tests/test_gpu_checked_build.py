"""Memory safety without compute-sanitizer (closed on the GPU pool): libmp3gpu_checked.so is the product kernels with every
input-dependent global load / store guarded (MP3_CHECK, go-mp3_b200/csrc/unit_logic.h); a guard that fails is reported as
an error of the call.  The whole range of inputs the parity tests use goes through it here — fixtures, CBR / VBR batches
cut into odd waves, quirk and fuzz streams, garbage descriptors, frame ranges, the streaming Decoder — with no guard
failing and the PCM identical to the product build's."""
import numpy as np
import pytest

from tools.synth import synth

pytestmark = pytest.mark.gpu


def both(pkg, **kw):
    return pkg.GpuEngine(0, checked=True, **kw), pkg.GpuEngine(0, **kw)


def test_checked_build_fixtures_batches_and_quirk_streams(pkg, classic_lame, mpeg2):
    streams = [classic_lame, mpeg2]
    streams += [synth.stream(synth.cfg3(i, 60)) for i in range(6)] + [synth.stream(synth.cfg4(i, 80)) for i in range(3, 43)]
    streams += [synth.stream(synth.wild(i, 48)) for i in range(40)] + [synth.stream(synth.fuzz(i, 24)) for i in range(60)]
    streams += [synth.stream(synth.cfg5(150))]
    pb = pkg.parse_streams(streams)
    for wave in (0, 997, 64):
        c, p = both(pkg, wave_granules=wave)
        a = c.decode(pb.main_data, pb.main_data_len, pb.units)   # raises Mp3Error if a guard failed
        b = p.decode(pb.main_data, pb.main_data_len, pb.units)
        assert np.array_equal(a, b), wave
        c.close(); p.close()


def test_checked_build_garbage_descriptors(pkg, classic_lame):
    rng = np.random.default_rng(99)
    pb = pkg.parse_streams([classic_lame])
    c = pkg.GpuEngine(0, checked=True)
    n = 4096
    bad = np.zeros(n * 2, dtype=pkg.UNIT_DTYPE)
    bad["bit_start"] = rng.integers(0, 16 * pb.main_data_len, n * 2, dtype=np.uint64)
    bad["bit_start"][::7] = np.uint64(2**63)
    bad["buf_end_rel"] = rng.integers(-5000, 2**31 - 1, n * 2, dtype=np.int64).astype(np.int32)
    for w in ("w0", "w1", "w2"):
        bad[w] = rng.integers(0, 2**32, n * 2, dtype=np.uint64).astype(np.uint32)
    bad["w2"] |= np.uint32(pkg.W2_VALID)
    c.decode(pb.main_data, pb.main_data_len, bad)   # no guard fails: garbage positions are clipped, garbage fields stay in range
    good = c.decode(pb.main_data, pb.main_data_len, pb.units)
    p = pkg.GpuEngine(0)
    assert np.array_equal(good, p.decode(pb.main_data, pb.main_data_len, pb.units))
    c.close(); p.close()


def test_checked_build_host_api_paths(pkg, classic_lame):
    """DecodeBatch (chunked, two device engines per GPU), frame ranges, split decode and the streaming Decoder on the checked build."""
    cfgs = [synth.cfg4(i, 40 + (i % 5)) for i in range(300)]
    buf, offs, lens = synth.batch(cfgs, 8)
    sb = pkg.StreamBuffer(buf, offs, lens)
    ce = pkg.Engine(devices=[0, 0], checked=True, chunk_frames=19, host_threads=8)
    pe = pkg.Engine(device=0, host_threads=8)
    ra, pa, _ = ce.decode_batch(sb)
    rb, pb_, _ = pe.decode_batch(sb)
    for a, b in zip(ra, rb):
        assert (a["pcm_bytes"], a["status"]) == (b["pcm_bytes"], b["status"])
        assert np.array_equal(pa[a["pcm_offset"]:a["pcm_offset"] + a["pcm_bytes"]], pb_[b["pcm_offset"]:b["pcm_offset"] + b["pcm_bytes"]])
    data = synth.stream(synth.cfg5(500))
    ix = pkg.StreamIndex(data)
    lin, rc, _ = pe.decode_stream_split(ix)
    lin = lin.copy()
    out, rc2, _ = ce.decode_stream_split(ix)
    assert rc == rc2 == 0 and np.array_equal(lin, out)
    got, rc3 = ce.decode_frames(ix, 123, 321, slot=1)
    assert rc3 == 0 and np.array_equal(got, lin[123 * 4608:321 * 4608])
    d = ce.new_decoder(classic_lame, slot=1)
    d.seek_to_time(3_000_000_000)
    a, _ = d.read_all()
    d2 = pe.new_decoder(classic_lame)
    d2.seek_to_time(3_000_000_000)
    b, _ = d2.read_all()
    assert a == b and len(a) > 500_000
    d.close(); d2.close(); ce.close(); pe.close()
