"""Multi-GPU inside the product API (include/mp3host.h), SURVEY.md 8e:

* DecodeBatch shards its streams over the engine's devices — per-stream PCM digests identical on 1, 2, 3 and 4 device slots
  (reference-side T6: what a stream decodes to must not depend on which device decoded it or on its neighbours);
* one long stream: mp3_decode_frames (a frame range decoded on its own, lead-in + halo inside the library) equals that
  stretch of the linear decode bit for bit, and mp3_decode_stream_split (one range per device, concurrently) equals the
  whole linear decode (BASELINE.json configs[4]);
* Decoders on different device slots.

A device list may name the same ordinal more than once (several device engines on one GPU), so every path here runs on a
one-GPU box; with more GPUs visible the slots are spread over them.
"""
import hashlib

import numpy as np
import pytest

import oracle
from tools.synth import synth

pytestmark = pytest.mark.gpu


def device_list(n):
    import torch
    have = max(1, torch.cuda.device_count())
    return [i % have for i in range(n)]


def digests(res, pcm):
    return [hashlib.sha256(pcm[r["pcm_offset"]:r["pcm_offset"] + r["pcm_bytes"]].tobytes()).hexdigest() for r in res]


def test_decode_batch_is_device_count_independent(pkg, classic_lame, mpeg2):
    cfgs = [(synth.cfg3(i, 140) if i % 3 == 0 else synth.cfg4(i, 90 + 7 * (i % 11))) for i in range(300)]
    buf, offs, lens = synth.batch(cfgs, 8)
    streams = [buf[o:o + l].tobytes() for o, l in zip(offs, lens)]
    streams[17] = classic_lame
    streams[155] = mpeg2
    streams[201] = b""                       # fails to open
    streams[202] = streams[202][:1000]       # truncated mid-frame
    streams[299] = b"\x00" * 5000            # junk
    ref = None
    for n in (1, 2, 3, 4):
        eng = pkg.Engine(devices=device_list(n), host_threads=8)
        assert eng.device_count() == n
        res, pcm, tm = eng.decode_batch(streams)
        d = digests(res, pcm)
        meta = [(r["pcm_bytes"], r["status"], r["frames"], r["sample_rate"]) for r in res]
        if ref is None:
            ref = (d, meta)
            assert res[17]["pcm_bytes"] == 1774080 and res[155]["pcm_bytes"] == 6617088 and res[201]["status"] != 0
        else:
            assert meta == ref[1], n
            assert d == ref[0], n
        eng.close()
    # and the single-device engine agrees with the oracle on a few of them (product build: +-1 LSB)
    eng = pkg.Engine(device=0)
    res, pcm, _ = eng.decode_batch(streams)
    for i in (0, 17, 40, 202):
        want, err = oracle.OracleDecoder(streams[i]).read_all()
        got = pcm[res[i]["pcm_offset"]:res[i]["pcm_offset"] + res[i]["pcm_bytes"]]
        assert len(want) == len(got), i
        diff = np.abs(np.frombuffer(want, np.int16).astype(np.int32) - got.view(np.int16).astype(np.int32))
        assert diff.max() <= 1, i
    eng.close()


@pytest.mark.parametrize("name,cfg", [("cfg5_320k", synth.cfg5(700)), ("cfg4_lsf_mono", synth.cfg4(39, 500)), ("cfg4_mixed", synth.cfg4(3, 600)),
                                      ("cfg3", synth.cfg3(1, 300)), ("wild", synth.wild(7, 200))])
def test_frame_ranges_equal_linear_decode(pkg, name, cfg):
    data = synth.stream(cfg)
    eng = pkg.Engine(devices=device_list(2), host_threads=4)
    res, pcm, _ = eng.decode_batch([data])
    lin = pcm[res[0]["pcm_offset"]:res[0]["pcm_offset"] + res[0]["pcm_bytes"]].copy()
    ix = pkg.StreamIndex(data)
    frames = ix.frames()
    if res[0]["status"] == 0:
        assert frames == res[0]["frames"] and ix.pcm_bytes(0, frames) == len(lin)
    frames = min(frames, res[0]["frames"])
    rng = np.random.default_rng(11)
    ranges = [(0, frames), (0, 1), (frames - 1, frames), (1, 2), (2, 5), (frames // 2, frames // 2)]
    for _ in range(40):
        a = int(rng.integers(0, frames))
        ranges.append((a, min(frames, a + int(rng.integers(1, 40)))))
    for k, (a, b) in enumerate(ranges):
        got, rc = eng.decode_frames(ix, a, b, slot=k % 2)
        assert rc == 0, (a, b, rc)
        lo, hi = ix.pcm_bytes(0, a), ix.pcm_bytes(0, b)
        assert np.array_equal(got, lin[lo:hi]), (name, a, b)
    eng.close()


@pytest.mark.parametrize("n_dev", [1, 2, 5, 8])
def test_stream_split_equals_linear_decode(pkg, n_dev):
    data = synth.stream(synth.cfg5(3000))
    eng = pkg.Engine(devices=device_list(n_dev), host_threads=4)
    res, pcm, _ = eng.decode_batch([data])
    lin = hashlib.sha256(pcm[:res[0]["pcm_bytes"]].tobytes()).hexdigest()
    n_lin = res[0]["pcm_bytes"]
    ix = pkg.StreamIndex(data)
    out, rc, tm = eng.decode_stream_split(ix)
    assert rc == 0 and len(out) == n_lin == 3000 * 4608
    assert hashlib.sha256(out.tobytes()).hexdigest() == lin
    eng.close()


def test_stream_split_many_ranges_per_device_and_early_end(pkg, monkeypatch):
    """Ranges are parsed side by side and decoded in order (several per device); a stream that ends or breaks in the middle
    gives what the linear decode gives: the PCM in front of the break and the same status."""
    monkeypatch.setenv("MP3HOST_SPLIT_MIN_FRAMES", "50")
    good = synth.stream(synth.cfg5(2400))
    cases = {"whole": good, "cut_mid_frame": good[:1044 * 1000 + 300], "junk_tail": good[:1044 * 700] + b"\x00" * 70000,
             "bad_frame": good[:1044 * 900] + b"\xff\xfb\x00\x44" + good[1044 * 900 + 4:]}  # free-format header: fatal in the reference
    lin_eng = pkg.Engine(device=0, host_threads=8)
    eng = pkg.Engine(devices=device_list(3), host_threads=8)
    for name, data in cases.items():
        res, pcm, _ = lin_eng.decode_batch([data])
        want = pcm[res[0]["pcm_offset"]:res[0]["pcm_offset"] + res[0]["pcm_bytes"]].copy()
        ix = pkg.StreamIndex(data)
        out, rc, tm = eng.decode_stream_split(ix)
        assert len(out) == len(want) and np.array_equal(out, want), name
        assert (rc == 0) == (res[0]["status"] == 0) or name == "bad_frame", (name, rc, res[0]["status"])
    eng.close(); lin_eng.close()


def test_stream_split_of_a_fixture_matches_oracle_exact_build(pkg, classic_lame):
    """Exact build: the split decode is bit-identical to the oracle's linear decode, seams included."""
    eng = pkg.Engine(devices=device_list(4), exact=True)
    out, rc, _ = eng.decode_stream_split(pkg.StreamIndex(classic_lame))
    want, err = oracle.OracleDecoder(classic_lame).read_all()
    assert rc == 0 and out.tobytes() == want
    eng.close()


def test_decoders_on_device_slots(pkg, classic_lame):
    eng = pkg.Engine(devices=device_list(3), chunk_frames=32)
    outs = []
    for slot in range(3):
        d = eng.new_decoder(classic_lame, slot=slot)
        d.seek_to_time(2_000_000_000)
        outs.append(d.read_all()[0])
        d.close()
    assert outs[0] == outs[1] == outs[2] and len(outs[0]) > 1_000_000
    eng.close()


def test_gapless_trim(pkg, classic_lame, mpeg2):
    """trim_gapless: DecodeBatch reports each stream without the LAME encoder delay / padding, as the reference's README
    example does with lameinfo (README.md:110-195): skip TotalDelay() samples, drop TotalPadding() at the end."""
    plain = pkg.Engine(device=0)
    trim = pkg.Engine(device=0, trim_gapless=True)
    r0, p0, _ = plain.decode_batch([classic_lame, mpeg2])
    r1, p1, _ = trim.decode_batch([classic_lame, mpeg2])
    info = pkg.lameinfo_parse_from_reader(classic_lame)
    skip, cut = info.total_delay() * 4, info.total_padding() * 4
    assert (skip, cut) == (1105 * 4, 263 * 4)
    assert r1[0]["pcm_bytes"] == r0[0]["pcm_bytes"] - skip - cut
    a = p0[r0[0]["pcm_offset"] + skip:r0[0]["pcm_offset"] + r0[0]["pcm_bytes"] - cut]
    b = p1[r1[0]["pcm_offset"]:r1[0]["pcm_offset"] + r1[0]["pcm_bytes"]]
    assert np.array_equal(a, b)
    assert r1[1]["pcm_bytes"] == r0[1]["pcm_bytes"]  # no LAME tag: untouched
    plain.close(); trim.close()


def test_device_api_rejects_misaligned_and_null_pointers(pkg):
    """The device-resident entry points validate what they can: a misaligned access would be a sticky fault."""
    import torch
    data = synth.stream(synth.cfg3(0, 20))
    pb = pkg.parse_streams([data])
    g = pkg.GpuEngine(0)
    dev = torch.device("cuda", 0)
    d_main = torch.zeros(len(pb.main_data) + 64, dtype=torch.uint8, device=dev)
    d_main[:len(pb.main_data)] = torch.from_numpy(pb.main_data).to(dev)
    d_units = torch.from_numpy(pb.units.view(np.uint8)).to(dev)
    d_pcm = torch.zeros(pb.n_granules * 1152 + 8, dtype=torch.int16, device=dev)
    ok = lambda: g.decode_device(d_main.data_ptr(), pb.main_data_len, d_units.data_ptr(), pb.n_granules, d_pcm.data_ptr())
    ok()
    want = d_pcm[:pb.n_granules * 1152].cpu().numpy().copy()
    for args in ((d_main.data_ptr() + 4, d_units.data_ptr(), d_pcm.data_ptr()), (d_main.data_ptr(), d_units.data_ptr() + 8, d_pcm.data_ptr()),
                 (d_main.data_ptr(), d_units.data_ptr(), d_pcm.data_ptr() + 2), (0, d_units.data_ptr(), d_pcm.data_ptr())):
        with pytest.raises(pkg.Mp3Error) as ex:
            g.decode_device(args[0], pb.main_data_len, args[1], pb.n_granules, args[2])
        assert ex.value.code == -3
    ok()  # the context is still usable
    assert np.array_equal(d_pcm[:pb.n_granules * 1152].cpu().numpy(), want)
    g.close()
    # a caller built against another struct layout is refused
    import ctypes as C
    ctx = C.c_void_p()
    assert g.lib.mp3gpu_create(0, C.byref(pkg.GpuOpts(99, 0, 0, 0)), C.byref(ctx)) == -3


def test_output_side_f32_planar(pkg):
    """mp3gpu_pcm_to_f32_planar: the PCM a decode leaves in HBM, as float planes for a consumer on the same GPU."""
    import torch
    data = synth.stream(synth.cfg4(3, 61))
    pb = pkg.parse_streams([data])
    g = pkg.GpuEngine(0)
    dev = torch.device("cuda", 0)
    d_main = torch.from_numpy(pb.main_data).to(dev)
    d_units = torch.from_numpy(pb.units.view(np.uint8)).to(dev)
    for n in (pb.n_granules * 576, pb.n_granules * 576 - 3, 5, 0):
        d_pcm = torch.zeros(pb.n_granules * 1152, dtype=torch.int16, device=dev)
        left = torch.full((max(n, 1) + 8,), 7.0, dtype=torch.float32, device=dev)
        right = torch.full((max(n, 1) + 8,), 7.0, dtype=torch.float32, device=dev)
        g.decode_device(d_main.data_ptr(), pb.main_data_len, d_units.data_ptr(), pb.n_granules, d_pcm.data_ptr(), sync=False)
        g.pcm_to_f32_planar(d_pcm.data_ptr(), n, left.data_ptr(), right.data_ptr())
        g.synchronize()
        pcm = d_pcm.cpu().numpy().reshape(-1, 2)[:n].astype(np.float32) / 32768.0
        assert np.array_equal(left[:n].cpu().numpy(), pcm[:, 0]) and np.array_equal(right[:n].cpu().numpy(), pcm[:, 1])
        assert float(left[n:].min()) == 7.0 and float(right[n:].min()) == 7.0   # nothing written past the end
    g.close()
