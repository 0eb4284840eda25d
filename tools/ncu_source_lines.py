"""Aggregate an ncu source page (ncu -i X.ncu-rep --page source --csv --print-source cuda,sass) by source line:
share of stall samples, share of issue slots, lanes per instruction.  Usage: ncu_source_lines.py report.ncu-rep [kernel-regex] [top]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if kern:
        cmd += ["--kernel-name", "regex:" + kern]
    txt = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file, line, c = None, None, {}
    agg = collections.OrderedDict()
    tot = [0, 0, 0]
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            c = {n: i for i, n in enumerate(r)}
            continue
        if r[0] == "Function Name":
            continue
        if r[0] != "":
            line = (cur_file, int(r[0]), r[1].strip()[:110])
            continue
        if r[2] in ("...", "-"):
            continue
        try:
            ie = int(r[c["Instructions Executed"]]); te = int(r[c["Thread Instructions Executed"]]); sm = int(r[c["# Samples"]])
        except (ValueError, KeyError, IndexError):
            continue
        a = agg.setdefault(line, [0, 0, 0, 0])
        a[0] += ie; a[1] += te; a[2] += 1; a[3] += sm
        tot[0] += ie; tot[1] += te; tot[2] += sm
    print(f"warp-instr {tot[0]}  thread-instr {tot[1]}  lanes/instr {tot[1] / max(tot[0], 1):.2f}  samples {tot[2]}")
    for (f, l, s), (ie, te, n, sm) in sorted(agg.items(), key=lambda kv: -kv[1][3])[:top]:
        print(f"samples {sm / max(tot[2], 1) * 100:5.1f}%  issue {ie / max(tot[0], 1) * 100:5.1f}%  lanes {te / max(ie, 1):5.1f}  sass {n:3d}  {f}:{l}  {s}")


if __name__ == "__main__":
    main()
