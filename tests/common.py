"""Shared helpers for the parity tests: map the oracle's per-frame taps onto the engine's per-granule unit slots."""
import numpy as np


def oracle_units_view(taps, n_frames):
    """Flatten oracle taps [frame][gr][ch] into the engine's unit-slot order: 2 slots per granule, LSF frames
    contribute one granule.  Returns dict of arrays indexed by unit slot, plus `live` (slot holds a channel)."""
    hdr = taps.header[:n_frames]
    lsf = ((hdr >> 19) & 3) != 3
    mono = ((hdr >> 6) & 3) == 3
    sel = []  # (frame, gr, ch, live)
    for f in range(n_frames):
        ngr = 1 if lsf[f] else 2
        for gr in range(ngr):
            sel.append((f, gr, 0, True))
            sel.append((f, gr, 1, not mono[f]))
    idx = np.array([(f, g, c) for f, g, c, _ in sel])
    live = np.array([l for *_, l in sel])
    out = {"live": live, "frame": idx[:, 0]}
    for name in ("is_", "count1", "scalefac_l", "scalefac_s", "part2_start", "xr_requant", "xr_reorder", "xr_stereo",
                 "xr_alias", "hybrid"):
        a = getattr(taps, name, None)
        if a is None:
            continue
        v = a[idx[:, 0], idx[:, 1], idx[:, 2]].copy()
        v[~live] = 0
        out[name] = v
    return out


def pcm_stats(a: np.ndarray, b: np.ndarray):
    """max |diff| and exact-match fraction between two int16 PCM arrays."""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return int(d.max()) if d.size else 0, float((d == 0).mean()) if d.size else 1.0
