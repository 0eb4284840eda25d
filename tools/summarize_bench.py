"""Prints the scaling table of DESIGN.md (e) from profiles/r02_bench_n{1,2,4,8}.json."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
print("| GPUs | cfg 3 device-resident | cfg 3 public API (D2H GB/s; ceiling; fraction) | cfg 4 device-resident | cfg 4 public API | cfg 5 split decode |")
print("|---|---|---|---|---|---|")
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")
    if not os.path.exists(p):
        continue
    d = json.load(open(p))
    e, c4, c5 = d["e2e"], d.get("cfg4") or {}, d.get("cfg5") or {}
    e4 = c4.get("e2e") or {}
    print(f"| {n} | {d['value'] / 1e3:.1f} | {e['value'] / 1e3:.1f} ({e['d2h_gbs']:.0f}; {e['copy_ceiling']['aggregate_gbs']:.0f}; {e['frac_of_copy_ceiling']:.2f}) | "
          f"{c4.get('value', 0) / 1e3:.1f} | {e4.get('value', 0) / 1e3:.1f} ({e4.get('frac_of_copy_ceiling', 0):.2f}) | "
          f"{c5.get('ms_per_decode', 0):.1f} ms ({c5.get('value', 0) / 1e3:.1f} G); one device {c5.get('linear_one_device_ms', 0):.0f} ms |")
