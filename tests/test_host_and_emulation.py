"""CPU tests of the product's host stage and of the K1/K2 unit logic (compiled for the host by tests/hostemu).

* mp3_parse_streams (tags, header sync, side info, reservoir -> bit-slices) against the oracle's Decoder on the
  fixtures, synthetic streams, fuzzed streams, tag/garbage cases and the reference's crasher inputs: same frame
  counts, PCM lengths, sample rates and terminal statuses.
* K1 (scalefactors + Huffman) and K2 (requantise/reorder/stereo/alias) logic bit-exact against the oracle's taps.
* Huffman LUT == reference tree walk, exhaustively per code word.
* Device tables == oracle tables bitwise, and the symmetries k_hybrid/k_synth rely on.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import common
import hostemu_lib
import oracle
from test_oracle_reference_vectors import ape_header, id3v1, id3v2, minimal_frame
from tools.synth import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def oracle_summary(data: bytes):
    d = oracle.OracleDecoder(data)
    if not d.ok():
        return {"opened": False, "err": d.open_err}
    pcm, err = d.read_all()
    return {"opened": True, "err": err, "pcm_bytes": len(pcm), "rate": d.sample_rate()}


def check_host_vs_oracle(pkg, data: bytes):
    res = pkg.parse_streams([data]).streams[0]
    o = oracle_summary(data)
    if not o["opened"]:
        assert res["frames"] == 0 and res["pcm_bytes"] == 0
        # NewDecoder's error: EOF family -> io.EOF (1), else the reference's error code
        assert res["status"] == (1 if o["err"] in (1, -1, -2) else o["err"])
        return
    assert res["pcm_bytes"] == o["pcm_bytes"] and res["sample_rate"] == o["rate"]
    assert res["status"] == o["err"]


def stream_cases():
    cs = [(f"cfg3_{i}", synth.cfg3(i, 40)) for i in range(2)]
    cs += [(f"cfg4_{i}", synth.cfg4(i, 60)) for i in (0, 3, 10, 19, 39, 59)]
    cs += [("cfg5", synth.cfg5(40))]
    cs += [(f"wild{i}", synth.wild(i)) for i in range(16)]
    cs += [(f"fuzz{i}", synth.fuzz(i)) for i in range(24)]
    return cs



def check_staged_cursor(pb, is16, meta, sf):
    """k_huffman's staged cursor (tiles of units, the stretch they read staged big-endian; positions outside the stretch
    read from main_data) decodes exactly what the register-window cursor does — one piece per unit as k_huffman stages
    them, whole stretches of many units, staging areas far too small (most reads take the fallback), odd tile sizes."""
    for tile, cap16 in ((1, 1 << 20), (1, 6), (256, 1 << 20), (64, 3), (100, 40), (256, 0)):  # tile 1 = k_huffman's per-unit pieces
        a, b, c = hostemu_lib.huffman_staged(pb.main_data, pb.units, tile, cap16)
        assert np.array_equal(a, is16) and np.array_equal(b, meta) and np.array_equal(c, sf), (tile, cap16)

@pytest.mark.parametrize("name,cfg", stream_cases(), ids=[n for n, _ in stream_cases()])
def test_host_stage_and_unit_logic_vs_oracle(pkg, name, cfg):
    data = synth.stream(cfg)
    check_host_vs_oracle(pkg, data)
    pb = pkg.parse_streams([data])
    dec, pcm, err, taps = oracle.decode_with_taps(data, cfg.n_frames + 2, stages=True)
    if dec is None or pb.n_granules == 0:
        return
    o = common.oracle_units_view(taps, taps.n_frames)
    assert len(pb.units) == len(o["live"])
    is16, meta, sf = hostemu_lib.huffman(pb.main_data, pb.units)
    assert np.array_equal(is16, o["is_"])                      # bit-exact Huffman integers
    assert np.array_equal(meta & 0x3FF, o["count1"])
    assert np.array_equal(sf[:, :22], o["scalefac_l"]) and np.array_equal(sf[:, 22:61], o["scalefac_s"])
    check_staged_cursor(pb, is16, meta, sf)
    xr = hostemu_lib.requant(pb.units, is16, meta, sf).reshape(-1, 576)
    assert np.array_equal(xr.view(np.uint32), o["xr_alias"].view(np.uint32))  # bit-exact spectrum after K2


def test_fixtures_unit_logic(pkg, classic_lame, mpeg2):
    for data, nf in ((classic_lame, 385), (mpeg2, 2872)):
        check_host_vs_oracle(pkg, data)
        pb = pkg.parse_streams([data])
        _, _, _, taps = oracle.decode_with_taps(data, nf + 2, stages=True)
        o = common.oracle_units_view(taps, taps.n_frames)
        is16, meta, sf = hostemu_lib.huffman(pb.main_data, pb.units)
        assert np.array_equal(is16, o["is_"]) and np.array_equal(meta & 0x3FF, o["count1"])
        check_staged_cursor(pb, is16, meta, sf)
        xr = hostemu_lib.requant(pb.units, is16, meta, sf).reshape(-1, 576)
        assert np.array_equal(xr.view(np.uint32), o["xr_alias"].view(np.uint32))


def edge_streams():
    f = minimal_frame()
    with open(os.path.join(GOLD, "fuzz_crashers.json")) as fh:
        crashers = [bytes.fromhex(h) for h in json.load(fh)["inputs_hex"]]
    cs = {
        "empty": b"", "one_byte": b"\xff", "three_bytes": b"\xff\xfb\x90", "header_only": b"\xff\xfb\x90\x44",
        "truncated_side_info": f[:20], "truncated_main_data": f[:200], "one_frame": f, "frame_and_half": f + f[:208],
        "id3v1_only": id3v1(), "id3v2_only": id3v2(64), "id3v2_truncated": id3v2(64)[:40], "tag_then_frames": id3v1() + f * 3,
        "ape_trailer": f * 5 + ape_header(18) + b"ARTIST\x00Test Artist", "garbage_70k": f * 2 + bytes(70000),
        "garbage_then_frames": bytes(range(1, 200)) + f * 4, "free_format": b"\xff\xfb\x00\x44" + bytes(413) + f,
        "mpeg25": b"\xff\xe3\x90\x44" + bytes(413) + f * 2, "layer2_only": b"\xff\xfd\x90\x44" * 100,
        "crc_frame": b"\xff\xfa\x90\x44" + bytes(413) + f,
        "huge_frame_320k_32k": b"\xff\xfb\xe8\x44" + bytes(1436) + f,
    }
    cs.update({f"crasher{i}": c for i, c in enumerate(crashers)})
    return cs


@pytest.mark.parametrize("name", list(edge_streams()))
def test_host_stage_edge_streams(pkg, name):
    check_host_vs_oracle(pkg, edge_streams()[name])


def test_batch_layout(pkg, classic_lame):
    """Streams are laid back to back: bit positions, PCM offsets and per-stream results of a batch equal the singles."""
    streams = [classic_lame[:30000], synth.stream(synth.cfg4(3, 20)), b"", synth.stream(synth.cfg4(39, 30)), b"\xff\xfb"]
    pb = pkg.parse_streams(streams, host_threads=3)
    off_gr = 0
    for i, s in enumerate(streams):
        one = pkg.parse_streams([s])
        r = pb.streams[i]
        assert (r["pcm_bytes"], r["frames"], r["status"], r["sample_rate"]) == tuple(
            one.streams[0][k] for k in ("pcm_bytes", "frames", "status", "sample_rate"))
        assert r["pcm_offset"] == off_gr * 2304
        n = one.n_granules
        a, b = pb.units[2 * off_gr:2 * (off_gr + n)], one.units
        assert np.array_equal(a["w0"], b["w0"]) and np.array_equal(a["w2"], b["w2"])
        if n:
            base = int(a["bit_start"][0]) - int(b["bit_start"][0])
            assert base % 32 == 0  # every stream's main data starts on a 4-byte boundary
            assert np.array_equal(a["bit_start"] - np.uint64(base), b["bit_start"])
        off_gr += n
    assert off_gr == pb.n_granules


def test_decode_batch_arena_bound_covers_every_stream(pkg, classic_lame, mpeg2):
    """DecodeBatch sizes its pinned arenas from a header-only frame walk before parsing (mp3host.cc): the bound must
    never be below what the parser produces — on well-formed streams it is exact, on truncated, fuzzed and malformed
    ones it may only be larger."""
    cases = dict(edge_streams())
    cases.update({"classic_lame": classic_lame, "mpeg2": mpeg2, "cut": classic_lame[:12345], "empty": b""})
    for i in range(40):
        cases[f"fuzz{i}"] = synth.stream(synth.fuzz(i))
        cases[f"wild{i}"] = synth.stream(synth.wild(i, 12))
    for i in range(8):
        cases[f"cfg4_{i}"] = synth.stream(synth.cfg4(i, 25))
    exact = 0
    for name, data in cases.items():
        ub = pkg.unit_slots_upper_bound(data)
        mb = pkg.main_bytes_upper_bound(data)
        pb = pkg.parse_streams([data])
        n = pb.n_granules * 2
        assert ub >= n, (name, ub, n)
        assert mb >= pb.main_data_len - 3, (name, mb, pb.main_data_len)  # main_data_len is padded to 4 bytes
        exact += ub == n and mb <= pb.main_data_len <= mb + 3
    assert exact >= len(cases) // 2


# ---- Huffman LUT == reference tree walk ----------------------------------------------------------------
def test_huffman_lut_equals_tree_walk_exhaustive():
    L, E = oracle.lib(), hostemu_lib.lib()
    rng = np.random.default_rng(7)
    for table in range(34):
        n, lin = C.c_int(), C.c_int()
        L.orc_huffman_table_info(table, C.byref(n), C.byref(lin))
        codes = []
        for i in range(n.value):
            x, y, hl, hc = C.c_int(), C.c_int(), C.c_int(), C.c_uint32()
            L.orc_huffman_table_code(table, i, C.byref(x), C.byref(y), C.byref(hl), C.byref(hc))
            codes.append((hl.value, hc.value))
        if not codes:
            codes = [(0, 0)]
        for hl, hc in codes:
            for _ in range(3):  # random continuations: linbits, signs, following data
                tail = int(rng.integers(0, 1 << 40))
                word = ((hc << 40) | tail) << (64 - 40 - hl) if hl else tail << 24
                data = int(word).to_bytes(8, "big")
                b = oracle.OrcBits()
                buf = C.create_string_buffer(data, 8)
                L.orc_bits_init(C.byref(b), buf, 8)
                ref = (C.c_int * 4)()
                assert L.orc_huffman_decode(C.byref(b), table, C.byref(ref)) == 0
                got = (C.c_int * 4)()
                used = E.emu_huff_one(table, data, 8, C.byref(got))
                assert list(got) == list(ref), (table, hl, hc)
                assert used == L.orc_bits_pos(C.byref(b))


def test_huffman_truncated_buffers_match():
    """Q3: reads at/after the buffer end return 0 and do not advance — LUT path == bit-serial path on short buffers."""
    L, E = oracle.lib(), hostemu_lib.lib()
    rng = np.random.default_rng(11)
    for _ in range(4000):
        table = int(rng.integers(0, 34))
        nbytes = int(rng.integers(1, 4))
        data = bytes(rng.integers(0, 256, nbytes, dtype=np.uint8))
        b = oracle.OrcBits()
        buf = C.create_string_buffer(data, nbytes)
        L.orc_bits_init(C.byref(b), buf, nbytes)
        ref = (C.c_int * 4)()
        L.orc_huffman_decode(C.byref(b), table, C.byref(ref))
        got = (C.c_int * 4)()
        used = E.emu_huff_one(table, data, nbytes, C.byref(got))
        assert list(got) == list(ref) and used == L.orc_bits_pos(C.byref(b)), (table, data.hex())


# ---- tables ------------------------------------------------------------------------------------------
def test_device_tables_equal_oracle_tables_bitwise():
    L = oracle.lib()
    pairs = [(0, L.orc_table_cos_n36, 18 * 36), (1, L.orc_table_cos_n12, 72), (2, L.orc_table_imdct_win, 144),
             (3, L.orc_table_synth_nwin, 2048), (4, L.orc_table_synth_dtbl, 512)]
    for which, fn, n in pairs:
        ref = np.ctypeslib.as_array(fn(), shape=(n,))
        got = hostemu_lib.table(which)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), which
    ref = np.ctypeslib.as_array(L.orc_table_powtab34(), shape=(8207,))
    assert np.array_equal(hostemu_lib.powtab34().view(np.uint64), ref.view(np.uint64))


def test_table_symmetries_used_by_the_kernels():
    c36 = hostemu_lib.table(0).reshape(18, 36)
    c12 = hostemu_lib.table(1).reshape(6, 12)
    N = hostemu_lib.table(3).reshape(64, 32)
    for p in range(9):
        assert np.array_equal(c36[:, 17 - p], -c36[:, p]) and np.array_equal(c36[:, 35 - p], c36[:, 18 + p])
    for p in range(3):
        assert np.array_equal(c12[:, 5 - p], -c12[:, p]) and np.array_equal(c12[:, 11 - p], c12[:, 6 + p])
    for i in range(16):
        assert np.array_equal(N[32 - i], -N[i])
    for k in range(1, 16):
        assert np.array_equal(N[48 + k], N[48 - k])
    assert np.all(N[48] == -1.0) and np.all(np.abs(N[16]) < 1e-13) and np.any(N[16] != 0)


def test_unit_logic_sweep_vs_oracle(pkg):
    """A wider net than the parametrised cases above: 540 more synthetic / wild / fuzzed streams (64 k units) through
    the host stage and the K1/K2 unit logic, every Huffman integer, count1, scalefactor and requantised/stereo/alias
    spectrum value bit-identical to the oracle's taps."""
    cases = [synth.wild(i, 40) for i in range(16, 216)] + [synth.fuzz(i, 30) for i in range(24, 224)]
    cases += [synth.cfg4(i, 50) for i in range(60, 160)] + [synth.cfg3(i, 30) for i in range(2, 42)]
    units = 0
    for cfg in cases:
        data = synth.stream(cfg)
        pb = pkg.parse_streams([data])
        dec, pcm, err, taps = oracle.decode_with_taps(data, cfg.n_frames + 2, stages=True)
        if dec is None or pb.n_granules == 0:
            continue
        o = common.oracle_units_view(taps, taps.n_frames)
        assert len(pb.units) == len(o["live"])
        is16, meta, sf = hostemu_lib.huffman(pb.main_data, pb.units)
        assert np.array_equal(is16, o["is_"]) and np.array_equal(meta & 0x3FF, o["count1"])
        assert np.array_equal(sf[:, :22], o["scalefac_l"]) and np.array_equal(sf[:, 22:61], o["scalefac_s"])
        a, b, c = hostemu_lib.huffman_staged(pb.main_data, pb.units, 1, 1 << 20)
        assert np.array_equal(a, is16) and np.array_equal(b, meta) and np.array_equal(c, sf)
        a, b, c = hostemu_lib.huffman_staged(pb.main_data, pb.units, 96, 24)
        assert np.array_equal(a, is16) and np.array_equal(b, meta) and np.array_equal(c, sf)
        xr = hostemu_lib.requant(pb.units, is16, meta, sf).reshape(-1, 576)
        same = (xr.view(np.uint32) == o["xr_alias"].view(np.uint32)) | (np.isnan(xr) & np.isnan(o["xr_alias"]))
        assert same.all()
        units += len(pb.units)
    assert units > 60000
