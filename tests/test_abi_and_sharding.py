"""C-ABI surface (no compute without a GPU) and the multi-GPU sharding logic (gloo, world_size 2, on CPU)."""
import ctypes as C
import os
import re
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header: str):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mp3(?:gpu)?_[a-z0-9_]+)\s*\(", src)))


@pytest.mark.parametrize("header,lib", [("mp3gpu.h", "libmp3gpu.so"), ("mp3gpu.h", "libmp3gpu_exact.so"), ("mp3host.h", "libmp3host.so")])
def test_library_exports_every_declared_symbol(header, lib):
    names = declared_functions(header)
    assert len(names) >= 15
    so = C.CDLL(os.path.join(ROOT, "go-mp3_b200", lib))
    missing = [n for n in names if not hasattr(so, n)]
    assert not missing, missing


def test_unit_descriptor_is_32_bytes(pkg):
    assert C.sizeof(pkg.Unit) == 32 and pkg.UNIT_DTYPE.itemsize == 32


def test_no_cpu_fallback_without_device(pkg):
    """Without a CUDA device the engine refuses to start: the product has no CPU decode path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.Mp3Error):
        pkg.GpuEngine(0)
    with pytest.raises(pkg.Mp3Error):
        pkg.Engine(0)


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use oracle/."""
    pk = os.path.join(ROOT, "go-mp3_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".cuh", ".h", ".inc")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "mp3_oracle" not in txt and "orc_" not in txt, os.path.join(dirpath, f)
    for lib in ("libmp3gpu.so", "libmp3host.so"):
        out = subprocess.run(["ldd", os.path.join(pk, lib)], capture_output=True, text=True).stdout
        assert "oracle" not in out


def test_shard_streams_partitions_exactly(pkg):
    for n in (0, 1, 7, 8, 4096, 65537):
        for world in (1, 2, 3, 8):
            parts = [pkg.shard_streams(n, world, r) for r in range(world)]
            assert sum(b - a for a, b in parts) == n
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_sharded_host_stage_gloo(tmp_path):
    """world_size-2 gloo run of the N>1 path's host logic: each rank synthesises + parses its own shard of the stream
    set (no data-path collective), and the gathered per-stream results equal a single-rank run."""
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        from __graft_entry__ import load_package
        from tools.synth import synth
        pkg = load_package()
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        N = 11
        a, b = pkg.shard_streams(N, world, rank)
        streams = [synth.stream(synth.cfg4(i, 12)) for i in range(a, b)]
        pb = pkg.parse_streams(streams, host_threads=2)
        mine = [(a + i, s["frames"], s["pcm_bytes"], s["status"]) for i, s in enumerate(pb.streams)]
        out = [None] * world
        dist.all_gather_object(out, mine)
        t = torch.tensor([float(pb.n_granules)])
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if rank == 0:
            allr = sorted(x for part in out for x in part)
            whole = pkg.parse_streams([synth.stream(synth.cfg4(i, 12)) for i in range(N)])
            ref = [(i, s["frames"], s["pcm_bytes"], s["status"]) for i, s in enumerate(whole.streams)]
            assert allr == ref, (allr, ref)
            assert int(t.item()) == whole.n_granules
            print("SHARD_OK")
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "SHARD_OK" in r.stdout
