"""Pins the oracle (oracle/mp3_oracle.c) against every vector the reference's own tests hold for this path.

  internal/bits/bits_test.go:23-112               bit reader values + out-of-bounds behaviour
  internal/frameheader/frameheader_test.go:39-286  frame geometry, sync limit, resync, Layer 1/2 rejection
  internal/maindata/huffman_test.go:14-46          region-count overflow clamps instead of failing
  trailing_tags_test.go:101-550                    Length()/ReadAll lengths with trailing/leading tags, Seek
  fuzzing_test.go:22-107                           historical crasher inputs
  fixture invariants         frame counts, PCM lengths, sample rates
"""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import oracle

L = oracle.lib()
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits_of(data: bytes):
    b = oracle.OrcBits()
    buf = C.create_string_buffer(data, len(data))
    L.orc_bits_init(C.byref(b), buf, len(data))
    return b, buf


# ---- bits_test.go ----------------------------------------------------------------------------------
def test_bit_out_of_bounds_reports_error():  # bits_test.go:23-39
    b, _k = bits_of(b"\xff\xff")
    for _ in range(16):
        L.orc_bits_bit(C.byref(b))
        assert b.err == 0
    assert L.orc_bits_bit(C.byref(b)) == 0 and b.err != 0


def test_bits_out_of_bounds_reports_error():  # bits_test.go:41-69
    b, _k = bits_of(b"\xab\xcd")
    assert L.orc_bits_bits(C.byref(b), 8) == 0xAB and b.err == 0
    assert L.orc_bits_bits(C.byref(b), 8) == 0xCD and b.err == 0
    assert L.orc_bits_bits(C.byref(b), 8) == 0 and b.err != 0


def test_bits_partial_out_of_bounds():  # bits_test.go:71-86
    b, _k = bits_of(b"\xff")
    L.orc_bits_bits(C.byref(b), 4)
    assert b.err == 0
    pos = L.orc_bits_pos(C.byref(b))
    assert L.orc_bits_bits(C.byref(b), 8) == 0 and b.err != 0
    assert L.orc_bits_pos(C.byref(b)) == pos  # a read that would cross the end does not advance (bits.go:65-68)


def test_bits_values():  # bits_test.go:88-112
    b, _k = bits_of(bytes([85, 170, 204, 51]))
    assert [L.orc_bits_bits(C.byref(b), 1) for _ in range(4)] == [0, 1, 0, 1]
    assert L.orc_bits_bits(C.byref(b), 8) == 90
    assert L.orc_bits_bits(C.byref(b), 12) == 2764


# ---- frameheader_test.go ----------------------------------------------------------------------------
def mpeg1(sf):
    return 0xFFFB9000 | ((sf & 3) << 10)


def mpeg2(sf):
    return 0xFFF39000 | ((sf & 3) << 10)


def test_samples_per_frame():  # :39-55
    assert L.orc_header_samples_per_frame(mpeg1(0)) == 1152
    assert L.orc_header_samples_per_frame(mpeg2(0)) == 576


def test_frame_duration():  # :57-105
    assert abs(L.orc_header_frame_duration_ns(mpeg1(0)) - 10**9 * 1152 // 44100) <= 1000
    assert abs(L.orc_header_frame_duration_ns(mpeg1(1)) - 24_000_000) <= 1000
    assert abs(L.orc_header_frame_duration_ns(mpeg2(0)) - 10**9 * 576 // 22050) <= 1000


def test_bytes_per_second():  # :107-138
    assert L.orc_header_bytes_per_second(mpeg1(0)) == 44100 * 4
    assert L.orc_header_bytes_per_second(mpeg1(1)) == 48000 * 4
    assert L.orc_header_bytes_per_second(mpeg1(2)) == 32000 * 4
    assert L.orc_header_bytes_per_second(mpeg2(0)) == 22050 * 4


def fh_read(data: bytes, pos=0):
    h, start, newpos, searched = C.c_uint32(), C.c_int64(), C.c_size_t(), C.c_int64()
    rc = L.orc_frameheader_read_mem(data, len(data), pos, C.byref(h), C.byref(start), C.byref(newpos), C.byref(searched))
    return rc, h.value, start.value


def test_read_sync_search_limit():  # :158-177
    rc, _, _ = fh_read(bytes(70000))
    assert rc == -2  # *SyncSearchLimitError


def test_read_valid_header_within_limit():  # :179-204
    data = bytearray(1004)
    data[1000:1004] = (0xFFFB9044).to_bytes(4, "big")
    rc, h, pos = fh_read(bytes(data))
    assert rc == 0 and pos == 1000 and L.orc_header_is_valid(h)


@pytest.mark.parametrize("h,want", [(0xFFFB9044, 1), (0xFFFF9044, 0), (0xFFFD9044, 0), (0xFFFFC420, 0)])
def test_is_valid_rejects_non_layer3(h, want):  # :206-247
    assert L.orc_header_is_valid(h) == want


def test_read_skips_non_layer3_headers():  # :249-286
    data = bytearray(204)
    data[100:104] = (0xFFFFC420).to_bytes(4, "big")
    data[200:204] = (0xFFFBB200).to_bytes(4, "big")
    rc, h, pos = fh_read(bytes(data))
    assert rc == 0 and pos == 200 and h == 0xFFFBB200


# ---- helpers of trailing_tags_test.go:15-98 ---------------------------------------------------------
def minimal_frame() -> bytes:
    f = bytearray(417)
    f[0:4] = b"\xff\xfb\x90\x44"
    return bytes(f)


def ape_header(tag_size: int) -> bytes:
    h = bytearray(32)
    h[0:8] = b"APETAGEX"
    h[8:12] = (2000).to_bytes(4, "little")
    h[12:16] = tag_size.to_bytes(4, "little")
    h[16:20] = (1).to_bytes(4, "little")
    h[20:24] = bytes([0xA0, 0, 0, 0x80])
    return bytes(h)


def id3v1() -> bytes:
    t = bytearray(128)
    t[0:3] = b"TAG"
    t[3:13] = b"Test Title"
    return bytes(t)


def id3v2(payload: int) -> bytes:
    return b"ID3\x04\x00\x00" + bytes([(payload >> 21) & 0x7F, (payload >> 14) & 0x7F, (payload >> 7) & 0x7F, payload & 0x7F]) + bytes(payload)


def decode_all(data: bytes):
    d = oracle.OracleDecoder(data)
    assert d.ok(), d.open_err
    length = d.length()
    pcm, err = d.read_all()
    return d, length, pcm, err


TRAILERS = {
    "ape": lambda: ape_header(18) + b"ARTIST\x00Test Artist",     # :101-153
    "id3v1": id3v1,                                                # :155-198
    "garbage100k": lambda: bytes([0x00, 0x01, 0x02, 0x03] * 25600),  # :264-307 (no sync words)
    "fake_syncs_70k": lambda: bytes(70000),                        # :409-479
}


@pytest.mark.parametrize("kind", list(TRAILERS))
def test_trailing_tags_lengths(kind):
    n = 10
    data = minimal_frame() * n + TRAILERS[kind]()
    d, length, pcm, err = decode_all(data)
    assert length == n * 1152 * 4
    assert err == 0 and len(pcm) == n * 1152 * 4


def test_id3v2_and_trailing_ape():  # :200-262
    data = id3v2(100) + minimal_frame() * 15 + ape_header(18) + b"ARTIST\x00Test Artist"
    d, length, pcm, err = decode_all(data)
    assert length == 15 * 4608 and len(pcm) == 15 * 4608 and err == 0


def test_multiple_consecutive_id3v2_tags():  # :481-530
    data = id3v2(50) + id3v2(80) + id3v2(10) + minimal_frame() * 8
    d, length, pcm, err = decode_all(data)
    assert length == 8 * 4608 and len(pcm) == 8 * 4608 and err == 0


def test_seek_with_trailing_tags():  # :309-372
    data = minimal_frame() * 20 + ape_header(18) + b"ARTIST\x00Test Artist"
    d = oracle.OracleDecoder(data)
    total = d.length()
    assert total > 0
    mid = (total // 2) & ~3
    pos, err = d.seek(mid, 0)
    assert err == 0 and pos == mid
    rest, err = d.read_all()
    assert err == 0 and len(rest) == total - mid
    pos, err = d.seek(0, 2)
    assert err == 0 and pos == total


def test_sync_limit_error_kind():  # :532-550
    assert b"no valid frame header found within" in L.orc_error_string(-2)


# ---- huffman_test.go:14-46 --------------------------------------------------------------------------
def pack_bits(fields):
    v, n = 0, 0
    for val, width in fields:
        v = (v << width) | (val & ((1 << width) - 1))
        n += width
    return v.to_bytes((n + 7) // 8, "big") if n % 8 == 0 else (v << (8 - n % 8)).to_bytes((n + 7) // 8, "big")


def test_region_count_overflow_clamps():
    """Region0Count=15, Region1Count=7 gives j = 24 > 22: clamp to 576, no error (maindata/huffman.go:57-63)."""
    fields = [(0, 9), (0, 3), (0, 8)]  # main_data_begin, private, scfsi
    for _ in range(4):  # gr x ch
        fields += [(100, 12), (10, 9), (200, 8), (0, 4), (0, 1), (0, 5), (0, 5), (0, 5), (15, 4), (7, 3), (0, 1), (0, 1), (0, 1)]
    side = pack_bits(fields)
    assert len(side) == 32
    frame = b"\xff\xfb\x90\x00" + side + bytes(417 - 36)
    d, length, pcm, err = decode_all(frame * 3)
    assert err == 0 and len(pcm) == 3 * 4608


# ---- fuzzing_test.go:22-107 -------------------------------------------------------------------------
def test_fuzz_crashers_do_not_crash():
    with open(os.path.join(GOLD, "fuzz_crashers.json")) as f:
        ins = [bytes.fromhex(h) for h in json.load(f)["inputs_hex"]]
    assert len(ins) == 10
    for data in ins:
        d = oracle.OracleDecoder(data)
        if d.ok():
            d.read_all()


# ---- fixtures + committed golden digests -------------------------------------------------------------
def test_fixture_invariants(classic_lame, mpeg2):
    for data, frames, nbytes, rate in ((classic_lame, 385, 1774080, 44100), (mpeg2, 2872, 6617088, 22050)):
        d = oracle.OracleDecoder(data)
        assert d.ok() and d.sample_rate() == rate and d.length() == nbytes and d.num_frame_starts() == frames
        pcm, err = d.read_all()
        assert err == 0 and len(pcm) == nbytes
        a = np.frombuffer(pcm, np.int16).astype(np.float64)
        # decoded audio, not noise: little clipping and most energy at low frequencies (first difference small)
        assert (np.abs(a) >= 32767).mean() < 1e-3
        assert np.abs(np.diff(a[::2])).mean() < 0.35 * np.abs(a).mean()


def test_oracle_matches_committed_golden_pcm(fixtures_dir):
    from tools import gen_golden
    from tools.synth import synth
    with open(os.path.join(GOLD, "oracle_pcm.json")) as f:
        gold = json.load(f)
    cases = {n: open(os.path.join(fixtures_dir, n + ".mp3"), "rb").read() for n in ("classic_lame", "mpeg2")}
    cases.update({n: synth.stream(c) for n, c in gen_golden.golden_synth_cases(synth)})
    assert set(cases) == set(gold)
    for name, data in cases.items():
        g = gold[name]
        assert hashlib.sha256(data).hexdigest() == g["input_sha256"], f"synthesiser output changed for {name}"
        d = oracle.OracleDecoder(data)
        pcm, err = d.read_all() if d.ok() else (b"", d.open_err)
        assert (len(pcm), err) == (g["pcm_bytes"], g["err"]), name
        assert hashlib.sha256(pcm).hexdigest() == g["pcm_sha256"], name
