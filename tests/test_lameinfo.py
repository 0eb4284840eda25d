"""The reference's lameinfo tests (lameinfo/lameinfo_test.go:12-588), test for test, against the C++ mirror behind
include/mp3host.h (mp3_lameinfo_*).  CPU only: lameinfo is metadata, nothing here touches the device."""
import os
import struct

import pytest

from conftest import ROOT, load_package

pkg = load_package()
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")
FLAG_FC, FLAG_BC, FLAG_TOC, FLAG_VS = pkg.LAME_FLAG_FRAME_COUNT, pkg.LAME_FLAG_BYTE_COUNT, pkg.LAME_FLAG_TOC, pkg.LAME_FLAG_VBR_SCALE


def build_test_frame(is_xing=False, flags=0, frame_count=0, byte_count=0, vbr_scale=0, lame_version=b"", encoder_delay=0,
                     encoder_padding=0):
    """lameinfo_test.go:12-115 buildTestFrame: MPEG-1 Layer III stereo 128 kbps header, 32 zero bytes of side info, tag."""
    frame = bytearray(b"\xff\xfb\x90\x00") + bytes(32)
    frame += b"Xing" if is_xing else b"Info"
    frame += bytes([0, 0, 0, flags & 0xFF])
    if flags & FLAG_FC:
        frame += struct.pack(">I", frame_count)
    if flags & FLAG_BC:
        frame += struct.pack(">I", byte_count)
    if flags & FLAG_TOC:
        frame += bytes(range(100))
    if flags & FLAG_VS:
        frame += struct.pack(">I", vbr_scale)
    if lame_version:
        frame += lame_version[:9].ljust(9, b"\x00")
        frame += bytes(12)
        frame += bytes([(encoder_delay >> 4) & 0xFF, ((encoder_delay << 4) & 0xF0) | ((encoder_padding >> 8) & 0x0F), encoder_padding & 0xFF])
        frame += bytes(12)
    if len(frame) < 417:
        frame += bytes(417 - len(frame))
    return bytes(frame)


def test_parse_xing_header():  # :129
    info = pkg.lameinfo_parse(build_test_frame(is_xing=True, flags=FLAG_FC | FLAG_BC, frame_count=1000, byte_count=500000))
    assert info.is_xing
    assert info.has_frame_count() and info.frame_count == 1000
    assert info.has_byte_count() and info.byte_count == 500000
    assert not info.has_toc() and not info.has_vbr_scale() and not info.has_lame_info()


def test_parse_info_header():  # :168
    info = pkg.lameinfo_parse(build_test_frame(is_xing=False, flags=FLAG_FC, frame_count=2000))
    assert not info.is_xing and info.frame_count == 2000


def test_parse_all_flags():  # :188
    info = pkg.lameinfo_parse(build_test_frame(is_xing=True, flags=FLAG_FC | FLAG_BC | FLAG_TOC | FLAG_VS, frame_count=5000,
                                               byte_count=2500000, vbr_scale=75))
    assert info.has_frame_count() and info.frame_count == 5000
    assert info.has_byte_count() and info.byte_count == 2500000
    assert info.has_toc() and info.toc == bytes(range(100))
    assert info.has_vbr_scale() and info.vbr_scale == 75


def test_parse_lame_info():  # :223
    info = pkg.lameinfo_parse(build_test_frame(is_xing=True, flags=FLAG_FC, frame_count=3000, lame_version=b"LAME3.100",
                                               encoder_delay=576, encoder_padding=1848))
    assert info.has_lame_info()
    assert info.lame_version == b"LAME3.100" and info.encoder_delay == 576 and info.encoder_padding == 1848


def test_total_delay():  # :252
    assert pkg.LameInfo().total_delay() == pkg.LAME_DECODER_DELAY == 529
    assert pkg.LameInfo(lame_version=b"LAME3.100", encoder_delay=576).total_delay() == 576 + 529


def test_total_padding():  # :270
    assert pkg.LameInfo().total_padding() == 0
    assert pkg.LameInfo(lame_version=b"LAME3.100", encoder_padding=1848).total_padding() == 1848 - 529
    assert pkg.LameInfo(lame_version=b"LAME3.100", encoder_padding=100).total_padding() == 0  # less than the decoder delay


def test_parse_no_xing_header():  # :297
    frame = b"\xff\xfb\x90\x00" + bytes(32) + b"XXXX" + bytes(400)
    with pytest.raises(pkg.NoXingHeader):
        pkg.lameinfo_parse(frame)


def test_parse_too_short():  # :314
    with pytest.raises(pkg.NoXingHeader):
        pkg.lameinfo_parse(b"\xff\xfb")


def test_parse_invalid_sync():  # :322
    with pytest.raises(pkg.NoXingHeader):
        pkg.lameinfo_parse(bytes(100))


def test_parse_from_reader():  # :332
    frame = build_test_frame(is_xing=True, flags=FLAG_FC | FLAG_BC, frame_count=1234, byte_count=567890, lame_version=b"LAME3.99",
                             encoder_delay=576, encoder_padding=1152)
    info = pkg.lameinfo_parse_from_reader(frame)
    assert info.frame_count == 1234 and info.byte_count == 567890
    assert info.lame_version == b"LAME3.99\x00"  # the field is 9 bytes; the reference keeps the padding NUL
    assert info.encoder_delay == 576 and info.encoder_padding == 1152


def test_parse_from_reader_read_errors():
    """io.ReadFull's errors pass through ParseFromReader (lameinfo.go:291-293,323-325)."""
    frame = build_test_frame(is_xing=True, flags=FLAG_FC, frame_count=1)
    for data, code in ((b"", pkg.MP3_EOF), (frame[:2], pkg.MP3_ERR_UNEXPECTED_EOF), (frame[:4], pkg.MP3_EOF),
                       (frame[:100], pkg.MP3_ERR_UNEXPECTED_EOF)):
        with pytest.raises(pkg.Mp3Error) as ex:
            pkg.lameinfo_parse_from_reader(data)
        assert ex.value.code == code, (len(data), ex.value.code)
    # header fields ParseFromReader refuses before sizing the frame (:306-308): reserved version, reserved layer, free
    # format, bitrate index 15, reserved sampling rate
    for hdr in (b"\xff\xeb\x90\x00", b"\xff\xf9\x90\x00", b"\xff\xfb\x00\x00", b"\xff\xfb\xf0\x00", b"\xff\xfb\x9c\x00"):
        with pytest.raises(pkg.NoXingHeader):
            pkg.lameinfo_parse_from_reader(hdr + bytes(500))


def test_parse_mpeg2_mono():  # :366
    frame = b"\xff\xf3\x50\xc0" + bytes(9) + b"Info" + bytes([0, 0, 0, FLAG_FC]) + bytes([0, 0, 0x03, 0xE8]) + bytes(200)
    assert pkg.lameinfo_parse(frame).frame_count == 1000


@pytest.mark.parametrize("version,want", [(b"LAME3.100", True), (b"LAME3.99", True), (b"L3.99abc", True), (b"Gogo3dex", True),
                                          (b"GOGO    ", True), (b"XXXXXXXX", False), (b"LAM", False), (b"", False)])
def test_is_lame_version(version, want):  # :404
    assert pkg.is_lame_version(version) == want


@pytest.mark.parametrize("delay,padding", [(0, 0), (576, 1848), (576, 0), (0, 1152), (4095, 4095), (1, 1), (256, 512), (2048, 2048)])
def test_encoder_delay_padding_bit_packing(delay, padding):  # :428
    info = pkg.lameinfo_parse(build_test_frame(is_xing=True, flags=0, lame_version=b"LAME3.100", encoder_delay=delay,
                                               encoder_padding=padding))
    assert info.encoder_delay == delay and info.encoder_padding == padding


def test_parse_real_lame_file():  # :471 (example/classic_lame.mp3, encoded with lame -V2)
    with open(f"{FIX}/classic_lame.mp3", "rb") as f:
        data = f.read()
    info = pkg.lameinfo_parse_from_reader(data)
    assert info.is_xing
    assert info.has_frame_count() and info.has_byte_count() and info.has_toc() and info.has_vbr_scale()
    assert 300 <= info.frame_count <= 500
    assert len(data) // 2 <= info.byte_count <= len(data)
    assert info.has_lame_info() and info.lame_version == b"LAME3.100"
    assert info.encoder_delay == 576
    assert 0 < info.encoder_padding <= 2000
    assert info.total_delay() == info.encoder_delay + 529
    assert info.vbr_scale <= 100
    # what this file says exactly (pinned here; the reference's test logs them)
    assert (info.frame_count, info.byte_count, info.vbr_scale, info.encoder_padding, info.total_padding()) == (384, 228657, 80, 792, 263)
    assert info.toc[0] == 0 and all(a <= b for a, b in zip(info.toc, info.toc[1:]))


def test_parse_real_mpeg2_file():  # :575 (no LAME header); :560 needs example/classic.mp3, which the reference tree lacks
    with open(f"{FIX}/mpeg2.mp3", "rb") as f:
        data = f.read()
    with pytest.raises(pkg.NoXingHeader):
        pkg.lameinfo_parse_from_reader(data)


def test_toc_offset_interpolates_like_the_xing_spec():
    """mp3_lameinfo_toc_offset (not in the reference, which never uses the TOC it parses)."""
    toc = bytes(min(255, int(256 * (k / 100.0) ** 1.5)) for k in range(100))  # a VBR-like, monotonic table
    info = pkg.LameInfo(flags=pkg.LAME_FLAG_TOC | pkg.LAME_FLAG_BYTE_COUNT, byte_count=1_000_000)
    for k in range(100):
        info._raw.toc[k] = toc[k]
    assert info.toc_offset(0.0) == 0
    assert info.toc_offset(1.0) == 1_000_000                       # entry "100" is 256 by definition
    assert info.toc_offset(0.50) == int(toc[50] / 256 * 1_000_000)
    mid = info.toc_offset(0.505)                                      # half-way between entries 50 and 51
    assert mid == int((toc[50] + (toc[51] - toc[50]) * 0.5) / 256 * 1_000_000)
    assert info.toc_offset(-3.0) == 0 and info.toc_offset(7.0) == 1_000_000 and info.toc_offset(float("nan")) == 0
    offs = [info.toc_offset(k / 1000.0) for k in range(1001)]
    assert offs == sorted(offs)
    # no byte count in the tag: the caller's stream size is used; no TOC: -1
    info2 = pkg.LameInfo(flags=pkg.LAME_FLAG_TOC)
    for k in range(100):
        info2._raw.toc[k] = toc[k]
    assert info2.toc_offset(1.0, stream_bytes=5000) == 5000
    assert pkg.LameInfo(flags=pkg.LAME_FLAG_BYTE_COUNT, byte_count=10).toc_offset(0.5) == -1
    # the real LAME file's TOC leads to frame boundaries' neighbourhood: offsets are monotonic and within the file
    with open(f"{FIX}/classic_lame.mp3", "rb") as f:
        data = f.read()
    real = pkg.lameinfo_parse_from_reader(data)
    assert real.has_toc()
    o = [real.toc_offset(k / 20.0, len(data)) for k in range(21)]
    assert o == sorted(o) and o[0] == 0 and 0 < o[10] < o[20] <= len(data)
