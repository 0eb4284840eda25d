#!/usr/bin/env python
"""bench.py — decoded stereo PCM Msamples/s of the B200 MP3 Layer III decode path.

    python bench.py --gpus N --steps K --warmup W            (N = 1; for N > 1 launch under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path (k_huffman: scalefactors + Huffman; k_hybrid: requantise/stereo/alias + IMDCT;
k_synth: polyphase synthesis + int16 store) over one batch of synthetic streams.  Workload (BASELINE.json configs[2]):
4,096 synthetic 30 s MPEG-1 Layer III 128 kbps CBR stereo streams, long blocks only, per GPU (weak scaling:
every rank decodes its own 4,096 streams; streams are independent, so there is no data-path collective).

`value`  : whole-job stereo samples/s with main data + unit descriptors already resident in HBM and PCM left in
           HBM; timed with CUDA events on the engine's compute stream around K back-to-back passes, max over ranks.
`e2e`    : the same metric through the C-ABI call with HOST buffers (mp3gpu_decode, include/mp3gpu.h): pinned host
           main data + descriptors -> device, PCM -> pinned host, all inside the timed region.
`roofline`: the dominant kernel against its bound (FP32 FMA issue or HBM), see DESIGN.md for the per-unit figures.
`cpu_baseline`: the oracle (C restatement of go-mp3; no Go toolchain in the image) on all host cores, bounded sample.
Inputs (2 GB main data + 0.6 GB descriptors) and outputs (21.7 GB PCM) per pass are far larger than the 126 MB L2,
so no explicit L2 flush is needed between timed iterations.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "decoded_stereo_pcm_msamples_per_s"
UNIT = "Msamples/s"

# Algorithmic work per granule-channel (SURVEY.md 8d / DESIGN.md): bytes each kernel must move, and flops.
# bits = main-data bytes per unit, measured from the batch.
# FLOPS_REFERENCE = the reference's direct forms (long blocks): k_hybrid = K2 (requantise 576 + MS 1,152 + alias 744)
#   + K3 (IMDCT 41,472 + window 1,152 + overlap 576); k_synth = K4 (matrixing 73,728 + window 9,216 + sums 9,216).
# FLOPS = what the shipped (fused-multiply-add) build executes: K3 is an 18-point DCT-IV through two 9-point DCT-IIIs
#   (258 flops per subband incl. window and overlap-add), K4's matrixing a 32-point Lee DCT (80 mul + 209 add per
#   slot) followed by the 512-tap window (1,024 flops) and the scale (2).  fp32 fractions use FLOPS; the roofline
#   picks, per kernel, whichever of HBM and FP32 it sits closer to.
FLOPS_REFERENCE = {"k1_huffman": 0.0, "k_hybrid": 2472.0 + 43200.0, "k_synth": 92160.0}
FLOPS = {"k1_huffman": 0.0, "k_hybrid": 2472.0 + 32 * 258.0, "k_synth": 18 * (289.0 + 1024.0 + 2.0)}


def kernel_bytes(main_bytes_per_unit):
    return {
        "k1_huffman": main_bytes_per_unit + 32 + 1152 + 4 + 32,   # bits + descriptor -> int16 lines + count1 + scalefactors
        "k_hybrid": 32 + 1152 + 4 + 32 + 2304,                   # K1 output + descriptor -> f32 subband samples
        "k_synth": 2304 + 1152,                                   # subband samples -> int16 PCM of this channel
    }


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id):
        self.gpu_id = gpu_id
        self.rows = []
        self.stamps = []
        self.stop_flag = False
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_id), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())
            self.stamps.append(time.time())

    def samples_since(self, t0):
        return sum(1 for t in self.stamps if t >= t0)

    def keep_since(self, t0):
        """Drop the samples taken before t0 (nvidia-smi is started early because it needs ~1 s to print its first row)."""
        keep = [i for i, t in enumerate(self.stamps) if t >= t0]
        self.rows = [self.rows[i] for i in keep]
        self.stamps = [self.stamps[i] for i in keep]

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(synth, first_stream, n_streams, n_frames, threads, kind):
    gen = {"cfg3": synth.cfg3, "cfg4": synth.cfg4}[kind]
    cfgs = [gen(first_stream + i, n_frames) for i in range(n_streams)]
    return synth.batch(cfgs, threads)


def run_reference(args, rank, world):
    """CPU arm: the oracle (port of the reference's algorithm; Go is not installed) on all host cores."""
    if rank != 0:
        return
    import oracle
    from tools.synth import synth
    cores = os.cpu_count() or 1
    per_step = max(cores * 4, 16)
    buf, offs, lens = make_workload(synth, 0, per_step, args.frames, cores, args.workload)
    streams = [buf[o:o + l].tobytes() for o, l in zip(offs, lens)]
    for _ in range(min(args.warmup, 1)):
        oracle.decode_streams_mt(streams[:cores], cores)
    t_tot, samples = 0.0, 0
    for _ in range(args.steps):
        secs, pcm_bytes, _ = oracle.decode_streams_mt(streams, cores)
        t_tot += secs
        samples += sum(pcm_bytes) // 4
    val = samples / t_tot / 1e6
    sample = f"{per_step} of the {args.streams} streams of the workload per step ({args.frames} frames each), {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_tot / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "streams_per_gpu": args.streams, "frames_per_stream": args.frames,
                       "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(args):
    if args.workload == "cfg3":
        return (f"{args.streams} synthetic 30 s MPEG-1 L3 128 kbps CBR stereo streams per GPU, long blocks only "
                f"(BASELINE.json configs[2]; {args.frames} frames each)")
    return (f"{args.streams} synthetic VBR streams per GPU with long/short/mixed blocks, MS+intensity stereo, deep "
            f"reservoir, 5% LSF (BASELINE.json configs[3]; {args.frames} frames each)")


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    from tools.synth import synth
    pkg = load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cores = os.cpu_count() or 1
    threads = max(1, cores // world)

    # ---- workload: synthesise, host-parse (tags/headers/side info/reservoir), stage in HBM -------------------------
    t0 = time.time()
    buf, offs, lens = make_workload(synth, rank * args.streams, args.streams, args.frames, threads, args.workload)
    sb = pkg.StreamBuffer(buf, offs, lens)
    t1 = time.time()
    pb = pkg.parse_streams(sb, threads)
    t2 = time.time()
    assert all(s["status"] == 0 and s["frames"] == args.frames for s in pb.streams), "synthetic stream failed to parse"
    n_gr = pb.n_granules
    n_units_valid = int(((pb.units["w2"] >> 25) & 1).sum())
    samples_per_step = n_gr * 576  # stereo samples
    d_main = torch.from_numpy(pb.main_data).to(dev)
    d_units = torch.from_numpy(pb.units.view(np.uint8)).to(dev)
    d_pcm = torch.empty(n_gr * 1152, dtype=torch.int16, device=dev)
    torch.cuda.synchronize()
    eng = pkg.GpuEngine(local_rank, wave_granules=args.wave)
    info = eng.device_info()

    def one_pass(sync):
        eng.decode_device(d_main.data_ptr(), pb.main_data_len, d_units.data_ptr(), n_gr, d_pcm.data_ptr(), sync=sync)

    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    sampler = ClockSampler(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
    sampler.start()  # started here: nvidia-smi takes about a second to deliver its first row; rows before the timed region are dropped
    for _ in range(max(args.warmup, 3)):
        one_pass(True)
    fp32_peak = eng.fp32_peak_tflops()

    # ---- timed region: K passes, CUDA events on the compute stream ------------------------------------------------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eng.synchronize()
    t_timed0 = time.time()
    ksum = {"k1_huffman": 0.0, "k_hybrid": 0.0, "k_synth": 0.0}
    launches = 0
    eng.event_record(0)
    for _ in range(args.steps):
        one_pass(False)
    eng.event_record(1)
    total_ms = eng.event_elapsed_ms(0, 1)
    eng.synchronize()
    torch.cuda.synchronize()
    # per-kernel event times: K more passes, read back after each (kept outside the event pair above so that the
    # read-back synchronisation never sits inside the headline number)
    for _ in range(args.steps):
        one_pass(True)
        t = eng.timings()
        for k in ksum:
            ksum[k] += t[k + "_ms"]
        launches = t["launches"]
    # the clock samples must come from the loaded GPU: keep decoding (outside the event pair) until nvidia-smi has
    # delivered a few rows since the timed region began
    while sampler.proc is not None and sampler.samples_since(t_timed0) < 3 and time.time() - t_timed0 < 4.0:
        one_pass(True)
        eng.synchronize()
    sampler.keep_since(t_timed0)
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = samples_per_step * world / (ms_per_step * 1e-3) / 1e6

    # ---- parity spot check against the oracle on the first stream of this rank (checker only) ---------------------
    parity = None
    if rank == 0:
        import oracle
        worst, fracs, checked = 0, [], []
        for i in sorted({0, args.streams // 2, args.streams - 1}):
            st = pb.streams[i]
            ref, err = oracle.OracleDecoder(sb.stream(i)).read_all()
            ref = np.frombuffer(ref, dtype=np.int16)
            o = st["pcm_offset"] // 2
            got = d_pcm[o:o + ref.size].cpu().numpy()
            diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
            worst = max(worst, int(diff.max()))
            fracs.append(float((diff == 0).mean()))
            checked.append(i)
        ref0, _ = oracle.OracleDecoder(sb.stream(0)).read_all()
        ref = np.frombuffer(ref0, dtype=np.int16)
        parity = {"streams_checked": checked, "max_abs_diff_lsb": worst, "exact_fraction": min(fracs),
                  "tolerance_lsb": 1}

    # ---- e2e: host buffers through the C ABI (pinned; H2D + kernels + D2H inside the timed region) ----------------
    e2e = None
    # Pinned host memory is bounded to ~48 GB over all ranks of the node: at N = 1 and 2 every rank runs its whole batch
    # through the host API; beyond that a stream-ordered prefix of the batch (per-GPU e2e is PCIe-bound and does not depend
    # on the batch size).  The prefix is a whole number of streams, so it is a self-contained submission.
    e2e_streams = args.streams
    per_stream_pcm = (n_gr * 2304) // max(args.streams, 1)
    budget = int(48e9) // world
    if per_stream_pcm * e2e_streams > budget:
        e2e_streams = max(1, budget // max(per_stream_pcm, 1))
    e2e_gr = sum(s["pcm_bytes"] for s in pb.streams[:e2e_streams]) // 2304
    full_gr, full_main_len = n_gr, pb.main_data_len
    if e2e_streams < args.streams:
        u = pb.units[: e2e_gr * 2]
        valid = (u["w2"] >> 25) & 1 == 1
        main_end = int(((u["bit_start"][valid].astype(np.int64) + u["buf_end_rel"][valid]).max() + 7) // 8)
        n_gr, main_len_e2e = e2e_gr, min(main_end, pb.main_data_len)
    else:
        main_len_e2e = pb.main_data_len
    h2d = main_len_e2e + n_gr * 2 * 32
    d2h = n_gr * 2304
    del d_pcm
    torch.cuda.empty_cache()
    numa_node = eng.bind_host_to_gpu_numa_node() if world > 1 else None  # keep each rank's pinned buffers next to its GPU
    try:
        p_main = eng.host_alloc(main_len_e2e + 64)
        p_units = eng.host_alloc(n_gr * 2 * 32)
        p_pcm = eng.host_alloc(d2h)
        C.memset(p_main, 0, main_len_e2e + 64)
        C.memmove(p_main, pb.main_data.ctypes.data, main_len_e2e)
        C.memmove(p_units, pb.units.ctypes.data, n_gr * 2 * 32)
        e2e_steps = max(1, min(args.steps, 3))
        eng.decode_host(p_main, main_len_e2e, p_units, n_gr, p_pcm)  # warm-up (allocates the staging ring)
        if world > 1:
            dist.barrier()
        tw = time.perf_counter()
        for _ in range(e2e_steps):
            eng.decode_host(p_main, main_len_e2e, p_units, n_gr, p_pcm)
        e2e_s = (time.perf_counter() - tw) / e2e_steps
        e2e_t = eng.timings()
        if world > 1:
            tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt.item())
        if rank == 0 and parity is not None:
            got = np.ctypeslib.as_array(C.cast(p_pcm, C.POINTER(C.c_int16)), shape=(ref.size,))
            parity["e2e_max_abs_diff_lsb"] = int(np.abs(got.astype(np.int32) - ref.astype(np.int32)).max())
        e2e = {"value": n_gr * 576 * world / e2e_s / 1e6, "unit": UNIT, "streams_per_gpu": int(e2e_streams), "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "api": "mp3gpu_decode (include/mp3gpu.h), pinned host buffers", "numa_node": numa_node,
               "h2d_ms": e2e_t["h2d_ms"], "d2h_ms": e2e_t["d2h_ms"]}
        for p in (p_main, p_units, p_pcm):
            eng.host_free(p)
    except MemoryError as ex:
        e2e = {"value": None, "unit": UNIT, "error": str(ex)}
    n_gr = full_gr

    # ---- informational: DecodeBatch from raw .mp3 bytes (host parse + gather + device), rank 0 at N = 1 ------------
    decode_batch = None
    if rank == 0 and world == 1 and not args.no_decode_batch:
        try:
            heng = pkg.Engine(local_rank, host_threads=cores)
            n_b = min(args.streams, 1024)  # bounded: the first 1,024 streams
            sub = pkg.StreamBuffer(buf, offs[:n_b], lens[:n_b])
            heng.decode_batch(sub)            # first call allocates the pinned arenas
            res_b, pcm_b, tm = heng.decode_batch(sub)
            decode_batch = {"streams": n_b, "value": tm["pcm_bytes"] / 4 / tm["total_s"] / 1e6, "unit": UNIT,
                            "parse_s": tm["parse_s"], "gather_s": tm["gather_s"], "device_s": tm["device_s"],
                            "host_threads": cores, "api": "mp3_decode_batch (include/mp3host.h): raw .mp3 bytes in, PCM out"}
            heng.close()
        except Exception as ex:  # informational only
            decode_batch = {"error": str(ex)}

    # ---- CPU baseline (rank 0, N = 1): oracle on all host cores over a bounded sample -----------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        n_s = min(args.streams, max(cores * 4, 16))
        streams = [sb.stream(i) for i in range(n_s)]
        secs, pcm_bytes, _ = oracle.decode_streams_mt(streams, cores)
        cpu_baseline = {"value": sum(pcm_bytes) / 4 / secs / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {n_s} of the {args.streams} streams ({args.frames} frames each), one thread per "
                                  f"stream on {cores} threads, {secs:.1f} s wall"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline ------------------------------------------------------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    kb = kernel_bytes(pb.main_data_len / max(n_units_valid, 1))
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except OSError:
        pass
    kernels = {}
    step_kernel_ms = sum(ksum.values()) / args.steps
    for k, ms_sum in ksum.items():
        ms = ms_sum / args.steps                     # all waves of one pass
        per_launch_ms = ms / max(launches // 3, 1)   # one launch = one wave of this kernel
        gbs = kb[k] * n_units_valid / (ms * 1e-3) / 1e9
        tfs = FLOPS[k] * n_units_valid / (ms * 1e-3) / 1e12
        kernels[k] = {"ms_per_step": ms, "ms_per_launch": per_launch_ms, "share": ms / step_kernel_ms,
                      "alg_bytes_per_unit": kb[k], "alg_flops_per_unit": FLOPS[k],
                      "reference_direct_form_flops_per_unit": FLOPS_REFERENCE[k], "hbm_gbs": gbs,
                      "hbm_frac": gbs / hbm_peak, "fp32_tflops": tfs, "fp32_frac": tfs / fp32_peak}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    kd = kernels[dom]
    units_per_launch = n_units_valid / max(launches // 3, 1)
    if kd["fp32_frac"] >= kd["hbm_frac"]:
        roofline = {"kernel": dom, "bound": "fp32", "achieved": kd["fp32_tflops"], "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": kd["fp32_frac"],
                    "peak_source": "FFMA micro-benchmark run in this process (mp3gpu_measure_fp32_peak); "
                                   "MEASURED_PEAKS.json has no non-tensor fp32 entry; tensor cores are not used "
                                   "(fp32 parity rules out tf32)"}
    else:
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kd["hbm_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kd["hbm_frac"], "peak_source": hbm_src}
    roofline["units_per_launch"] = units_per_launch
    roofline["traffic"] = traffic.get(dom)
    roofline["also"] = {"hbm_frac": kd["hbm_frac"], "fp32_frac": kd["fp32_frac"], "hbm_peak_gbs": hbm_peak,
                        "fp32_peak_tflops": fp32_peak}
    # whole-pipeline view: algorithmic bytes of the fused minimum (bits + descriptor in, PCM out) and all flops
    pipe_bytes = (pb.main_data_len + n_gr * 2 * 32 + n_gr * 2304)
    pipe_flops = sum(FLOPS.values()) * n_units_valid
    pipeline = {"hbm_gbs_fused_minimum": pipe_bytes / (ms_per_step * 1e-3) / 1e9,
                "fp32_tflops": pipe_flops / (ms_per_step * 1e-3) / 1e12,
                "fp32_frac": pipe_flops / (ms_per_step * 1e-3) / 1e12 / fp32_peak}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "streams_per_gpu": args.streams, "frames_per_stream": args.frames,
                       "granules_per_gpu": int(n_gr), "granule_channels_per_gpu": n_units_valid,
                       "main_data_bytes_per_gpu": int(pb.main_data_len), "pcm_bytes_per_gpu": int(n_gr * 2304),
                       "l2": "inputs (main data + descriptors) and outputs per pass are >> 126 MB L2; no explicit flush",
                       "parallelism": f"streams sharded over {world} GPU(s), no collective", "wave_granules": args.wave or 2097152,
                       "device": info, "setup_s": {"synthesise": t1 - t0, "host_parse": t2 - t1}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches * args.steps), "roofline": roofline,
            "kernels": kernels, "pipeline": pipeline, "cpu_baseline": cpu_baseline, "parity": parity,
            "decode_batch_from_mp3_bytes": decode_batch}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg4"])
    ap.add_argument("--streams", type=int, default=4096, help="streams per GPU")
    ap.add_argument("--frames", type=int, default=1149, help="frames per stream (1149 = 30 s at 44.1 kHz)")
    ap.add_argument("--wave", type=int, default=0, help="granules per kernel wave (0 = engine default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode-batch", action="store_true")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
