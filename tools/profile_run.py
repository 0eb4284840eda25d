"""Small fixed workload for ncu captures: a few device-resident passes over N synthetic cfg3/cfg4 streams."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from tools.synth import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=512)
    ap.add_argument("--frames", type=int, default=1149)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--wave", type=int, default=0, help="granules per kernel wave (0 = engine default)")
    args = ap.parse_args()
    import torch
    pkg = load_package()
    gen = {"cfg3": synth.cfg3, "cfg4": synth.cfg4}[args.workload]
    buf, offs, lens = synth.batch([gen(i, args.frames) for i in range(args.streams)], os.cpu_count() or 1)
    pb = pkg.parse_streams(pkg.StreamBuffer(buf, offs, lens), os.cpu_count() or 1)
    dev = torch.device("cuda", 0)
    d_main = torch.from_numpy(pb.main_data).to(dev)
    d_units = torch.from_numpy(pb.units.view(np.uint8)).to(dev)
    d_pcm = torch.empty(pb.n_granules * 1152, dtype=torch.int16, device=dev)
    eng = pkg.GpuEngine(0, wave_granules=args.wave)
    for _ in range(args.passes):
        eng.decode_device(d_main.data_ptr(), pb.main_data_len, d_units.data_ptr(), pb.n_granules, d_pcm.data_ptr())
        print(eng.timings())


if __name__ == "__main__":
    main()
