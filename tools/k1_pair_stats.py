"""Code-length statistics of the big_values pairs of a synthetic workload (host emulation; no GPU): how many LUT lookups a
decoder would need that takes two pairs at once whenever both fit into b index bits."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package  # noqa: E402
from tools.synth import synth  # noqa: E402
import hostemu_lib  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    pkg = load_package()
    gen = {"cfg3": synth.cfg3, "cfg4": synth.cfg4}[wl]
    buf, offs, lens = synth.batch([gen(i, 200) for i in range(n)], 4)
    pb = pkg.parse_streams(pkg.StreamBuffer(buf, offs, lens))
    main, units = pb.main_data, pb.units
    import ctypes as C
    L = hostemu_lib.lib()
    hist = np.zeros(64, np.int64)
    nl = np.zeros(17, np.int64)
    npairs = C.c_longlong(0)
    main = np.ascontiguousarray(main)
    units = np.ascontiguousarray(units)
    L.emu_pair_stats(C.c_void_p(main.ctypes.data), C.c_ulonglong((len(main) - 64) * 8), C.c_void_p(units.ctypes.data), C.c_longlong(len(units)),
                     C.c_void_p(hist.ctypes.data), C.c_void_p(nl.ctypes.data), C.byref(npairs))
    tot = npairs.value
    print(wl, "pairs", tot, "mean bits/pair", float((hist * np.arange(64)).sum()) / max(tot, 1))
    print("length histogram (share):", {i: round(float(h) / tot, 3) for i, h in enumerate(hist) if h})
    for b in (6, 7, 8, 9, 10, 12, 16):
        print(f"dual lookup with {b}-bit index: {nl[b] / tot:.3f} lookups per pair")


if __name__ == "__main__":
    main()
