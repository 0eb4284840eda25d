"""Independent sanity of the oracle against the Layer III signal model (not against the Go code): a stream whose
granules carry ONE spectral line k must decode to a narrow-band signal centred at (k + 1/2) * fs / 1152, its level must
follow global_gain as 2^(gain/4), and a line in an odd subband must not be mirrored (frequency inversion, subband order,
IMDCT/synthesis phase conventions).  The reference ships no golden PCM ("parity unpinned"); this pins the
oracle's pipeline to the standard's mathematics, so a mis-read of the reference that still "sounds plausible" is caught.
"""
import numpy as np
import pytest

import oracle
from test_oracle_reference_vectors import pack_bits

FS = 44100


def tone_stream(line: int, gain: int, frames: int = 40) -> bytes:
    """MPEG-1 44.1 kHz mono, 128 kbps, long blocks, table 1 (codes: (0,0)=1, (0,1)=001, (1,0)=01, (1,1)=000)."""
    npairs = line // 2 + 1
    bits = [(1, 1)] * (npairs - 1) + ([(1, 2)] if line % 2 == 0 else [(1, 3)]) + [(0, 1)]  # last pair (1,0)/(0,1), sign +
    p23 = sum(w for _, w in bits)
    gr = [(p23, 12), (npairs, 9), (gain, 8), (0, 4), (0, 1), (1, 5), (1, 5), (1, 5), (15, 4), (7, 3), (0, 1), (0, 1), (0, 1)]
    side = pack_bits([(0, 9), (0, 5), (0, 4)] + gr + gr)
    assert len(side) == 17
    payload = pack_bits(bits + bits)
    frame = b"\xff\xfb\x90\xc4" + side + payload
    frame += bytes(417 - len(frame))
    return frame * frames


def spectrum_peak(pcm: np.ndarray):
    x = pcm[4608:].astype(np.float64)  # skip the start-up transient
    x = x * np.hanning(x.size)
    mag = np.abs(np.fft.rfft(x))
    k = int(np.argmax(mag))
    return k * FS / x.size, mag


@pytest.mark.parametrize("line", [3, 10, 17, 18, 25, 40, 100, 201, 302, 450])
def test_single_line_decodes_to_its_frequency(line):
    d = oracle.OracleDecoder(tone_stream(line, 190))
    assert d.ok()
    pcm, err = d.read_all()
    assert err == 0
    left = np.frombuffer(pcm, np.int16)[::2]
    f_peak, mag = spectrum_peak(left)
    f_expect = (line + 0.5) * FS / 1152
    assert abs(f_peak - f_expect) < FS / 1152, (line, f_peak, f_expect)   # within one MDCT bin
    # energy is concentrated around the line: > 90 % within +-2 bins
    freqs = np.arange(mag.size) * FS / (2 * (mag.size - 1))
    near = np.abs(freqs - f_expect) < 2.5 * FS / 1152
    assert (mag[near] ** 2).sum() > 0.9 * (mag ** 2).sum()


def test_level_follows_global_gain():
    rms = []
    for gain in (170, 178, 186):
        pcm, _ = oracle.OracleDecoder(tone_stream(40, gain)).read_all()
        a = np.frombuffer(pcm, np.int16)[::2][4608:].astype(np.float64)
        rms.append(np.sqrt((a ** 2).mean()))
    assert rms[1] / rms[0] == pytest.approx(4.0, rel=0.02) and rms[2] / rms[1] == pytest.approx(4.0, rel=0.02)  # 2^(8/4)
    # absolute level: one line of magnitude 2^((gain-210)/4) gives a sinusoid of that order (x 32767), not x10 off
    expect = 2.0 ** ((170 - 210) / 4) * 32767
    assert 0.2 * expect < rms[0] < 3.0 * expect
