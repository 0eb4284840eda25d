"""The host mirror of *mp3.Decoder and DecodeBatch (include/mp3host.h) against the oracle's Decoder (decode.go).

Ports the behaviours of time_seek_test.go (on classic_lame.mp3 — classic.mp3 is not shipped — and mpeg2.mp3) and of
trailing_tags_test.go to the GPU-backed Decoder: Length/Duration/Position/Remaining/Progress/SamplePosition/SampleCount,
SeekToTime/SeekToSample/Skip clamping and alignment, seek determinism, non-seekable sources, error kinds.
The exact build is used so PCM is compared bit for bit; chunked decode-ahead must be invisible to the caller.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from test_oracle_reference_vectors import ape_header, id3v2, minimal_frame
from tools.synth import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def eng(pkg):
    e = pkg.Engine(0, chunk_frames=37, exact=True)
    yield e
    e.close()


def both(eng, data, seekable=True):
    return eng.new_decoder(data, seekable), oracle.OracleDecoder(data, seekable)


def same_state(d, o):
    assert d.length() == o.length() and d.sample_rate() == o.sample_rate()
    assert d.bytes_per_frame() == o.bytes_per_frame()
    assert d.duration_ns() == o.duration_ns() and d.position_ns() == o.position_ns()
    assert d.remaining_ns() == o.remaining_ns() and d.progress() == o.progress()
    assert d.sample_position() == o.sample_position() and d.sample_count() == o.sample_count()


def test_open_state_and_readall(eng, classic_lame, mpeg2):
    for data in (classic_lame, mpeg2):
        d, o = both(eng, data)
        same_state(d, o)           # TestDuration_Seekable / _MPEG2, TestPosition_Initial, TestRemaining_Initial, ...
        a, ea = d.read_all()
        b, eb = o.read_all()
        assert ea == eb == 0 and a == b
        same_state(d, o)           # TestProgress_End
        assert d.read(16) == (b"", 1)  # io.EOF again


def test_small_reads_track_position(eng, classic_lame):  # TestPosition_AfterRead, TestSamplePosition_AfterRead
    d, o = both(eng, classic_lame)
    for n in (1, 3, 4096, 4608, 5000, 17):
        a, ea = d.read(n)
        b, eb = o.read(n)
        assert (a, ea) == (b, eb)
        same_state(d, o)


def test_non_seekable(eng, classic_lame):  # TestDuration/Position/SeekToTime/Skip/Remaining/Progress/SampleCount _NonSeekable
    d, o = both(eng, classic_lame, seekable=False)
    assert d.length() == -1 and d.duration_ns() == -1 and d.remaining_ns() == -1 and d.progress() == -1 and d.sample_count() == -1
    same_state(d, o)
    for fn, arg in (("seek_to_time", 10**9), ("skip", 10**9), ("seek_to_sample", 1000)):
        with pytest.raises(Exception) as ei:
            getattr(d, fn)(arg)
        assert ei.value.code == -11  # "mp3: seek not supported on non-seekable source"
        assert getattr(o, fn)(arg) == -11
    a, _ = d.read(10000)
    b, _ = o.read(10000)
    assert a == b


@pytest.mark.parametrize("t_ms", [0, 1, 26, 500, 3000, 5000, -5, 10**7])
def test_seek_to_time_matches_oracle(eng, classic_lame, t_ms):  # TestSeekToTime_Start/Middle/End/Negative/BeyondEnd/Alignment
    d, o = both(eng, classic_lame)
    d.seek_to_time(t_ms * 10**6)
    assert o.seek_to_time(t_ms * 10**6) == 0
    same_state(d, o)
    assert d.sample_position() * 4 % 4 == 0
    a, ea = d.read(20000)
    b, eb = o.read(20000)
    assert (a, ea) == (b, eb)      # includes the reference's post-seek quirk (Q9): frame f-1 decoded from zero state


def test_skip_and_seek_to_sample(eng, mpeg2):  # TestSkip_*, TestSeekToSample_Valid/_Clamping
    d, o = both(eng, mpeg2)
    for fn, arg in (("skip", 2 * 10**9), ("skip", -10**9), ("skip", -10**12), ("skip", 10**13), ("seek_to_sample", 44100),
                    ("seek_to_sample", -7), ("seek_to_sample", 10**12), ("seek_to_sample", 123457)):
        getattr(d, fn)(arg)
        assert getattr(o, fn)(arg) == 0
        same_state(d, o)
        a, ea = d.read(9000)
        b, eb = o.read(9000)
        assert (a, ea) == (b, eb), (fn, arg)
        same_state(d, o)


def test_audio_integrity_across_seeks(eng, classic_lame):  # TestIntegration_AudioIntegrity, _SeekToStart (:1010-1144)
    d = eng.new_decoder(classic_lame)
    first, _ = d.read(50000)
    d.seek_to_time(2 * 10**9)
    a, _ = d.read(30000)
    d.seek_to_time(4 * 10**9)
    d.read(1000)
    d.seek_to_time(2 * 10**9)
    b, _ = d.read(30000)
    assert a == b                  # seek -> read -> seek away -> seek back gives byte-identical PCM
    d.seek(0, 0)
    again, _ = d.read(50000)
    assert again == first          # seek to 0 reproduces the first bytes


@pytest.mark.parametrize("off,whence", [(-100, 0), (-10**9, 1), (-10**12, 2), (0, 2), (7, 0), (4608 * 3 + 5, 0), (0, 1), (5, 9)])
def test_raw_seek_matches_oracle(eng, classic_lame, off, whence):  # TestSeek_Negative*ShouldNotPanic, whence errors
    d, o = both(eng, classic_lame)
    d.read(10000); o.read(10000)
    ro, eo = o.seek(off, whence)
    try:
        rd, ed = d.seek(off, whence), 0
    except Exception as ex:
        rd, ed = 0, ex.code
    assert (rd, ed) == (ro, eo)
    a, ea = d.read(12000)
    b, eb = o.read(12000)
    assert (a, ea) == (b, eb)
    same_state(d, o)


def test_trailing_and_leading_tags(eng):  # trailing_tags_test.go
    f = minimal_frame()
    for data in (f * 10 + ape_header(18) + b"ARTIST\x00Test Artist", id3v2(100) + f * 15 + bytes(70000), id3v2(50) + id3v2(80) + f * 8):
        d, o = both(eng, data)
        same_state(d, o)
        mid = (d.length() // 2) & ~3
        assert d.seek(mid, 0) == mid and o.seek(mid, 0) == (mid, 0)
        a, ea = d.read_all()
        b, eb = o.read_all()
        assert (a, ea) == (b, eb) and len(a) == d.length() - mid
        assert d.seek(0, 2) == d.length()


def test_new_decoder_errors_match(eng):
    with open(os.path.join(GOLD, "fuzz_crashers.json")) as fh:
        cases = [bytes.fromhex(h) for h in json.load(fh)["inputs_hex"]]
    cases += [b"", b"\xff", b"ID3", b"TAG" + bytes(10), b"\xff\xe3\x90\x44" + bytes(500), b"\xff\xfb\x00\x44" + bytes(500), bytes(70000)]
    for data in cases:
        o = oracle.OracleDecoder(data)
        try:
            d = eng.new_decoder(data)
            assert o.ok()
            a, ea = d.read_all()
            b, eb = o.read_all()
            assert (a, ea) == (b, eb)
        except Exception as ex:
            assert not o.ok()
            assert ex.code == (1 if o.open_err in (1, -1, -2) else o.open_err)


def test_error_mid_stream_then_continue(eng):
    """A fatal frame (big_values > 288) ends io.ReadAll with the reference's error after the PCM decoded so far; the
    next Read resumes behind it with zero state (d.frame = nil, decode.go:47)."""
    for i in range(24):
        data = synth.stream(synth.fuzz(i))
        d, o = both(eng, data) if oracle.OracleDecoder(data).ok() else (None, None)
        if d is None:
            continue
        for _ in range(4):
            a, ea = d.read_all()
            b, eb = o.read_all()
            assert (a, ea) == (b, eb), i


def test_decode_batch_matches_oracle(pkg, classic_lame, mpeg2):
    with open(os.path.join(GOLD, "oracle_pcm.json")) as fh:
        gold = json.load(fh)
    from tools import gen_golden
    cases = {"classic_lame": classic_lame, "mpeg2": mpeg2}
    cases.update({n: synth.stream(c) for n, c in gen_golden.golden_synth_cases(synth)})
    names = list(cases)
    e = pkg.Engine(0, host_threads=4, wave_granules=1000, exact=True)
    res, pcm, tm = e.decode_batch([cases[n] for n in names] + [b"", b"junk" * 100])
    assert len(res) == len(names) + 2
    for n, r in zip(names, res):
        g = gold[n]
        seg = pcm[r["pcm_offset"]:r["pcm_offset"] + r["pcm_bytes"]].tobytes()
        assert r["pcm_bytes"] == g["pcm_bytes"] and r["status"] == g["err"], n
        assert hashlib.sha256(seg).hexdigest() == g["pcm_sha256"], n   # committed golden digest of the oracle's PCM
    assert res[-2]["status"] == 1 and res[-1]["status"] == 1 and res[-1]["pcm_bytes"] == 0  # io.EOF from NewDecoder
    assert tm["n_granules"] * 2304 == tm["pcm_bytes"]
    e.close()


def test_decode_batch_chunked_pipeline_equals_single_call(pkg, classic_lame):
    """Batches of >= 256 streams are cut into chunks whose parse/gather overlaps the previous chunk's device call
    (mp3host.cc).  The result must be byte-identical to decoding the same streams in small (single-chunk) batches,
    with every stream's PCM laid out back to back; empty / junk / truncated streams ride along."""
    streams = []
    for i in range(300):
        c = synth.cfg4(1000 + i, 6 + i % 5) if i % 3 else synth.cfg3(2000 + i, 5 + i % 7)
        streams.append(synth.stream(c))
    streams[17] = b""
    streams[130] = b"junk" * 64
    streams[131] = classic_lame[:5000]          # cut inside a frame
    streams[299] = classic_lame[:20000]
    e = pkg.Engine(0, host_threads=4, exact=True)
    res, pcm, tm = e.decode_batch(streams)
    got = [bytes(pcm[r["pcm_offset"]:r["pcm_offset"] + r["pcm_bytes"]]) for r in res]
    stat = [(r["status"], r["frames"], r["sample_rate"], r["pcm_bytes"]) for r in res]
    # back to back, in order
    off = 0
    for r in res:
        assert r["pcm_offset"] == off
        off += r["pcm_bytes"]
    assert off == tm["pcm_bytes"] == tm["n_granules"] * 2304
    for lo in range(0, 300, 100):                # 100 streams per call: the single-chunk path
        res1, pcm1, _ = e.decode_batch(streams[lo:lo + 100])
        for k, r in enumerate(res1):
            assert (r["status"], r["frames"], r["sample_rate"], r["pcm_bytes"]) == stat[lo + k], lo + k
            assert bytes(pcm1[r["pcm_offset"]:r["pcm_offset"] + r["pcm_bytes"]]) == got[lo + k], lo + k
    assert res[17]["status"] == 1 and res[130]["status"] == 1 and res[131]["pcm_bytes"] > 0
    e.close()
