"""Per-kernel SASS evidence for profiles/: code size, instruction-class histogram and the mnemonics that matter on sm_100a.

    python tools/sass_histogram.py go-mp3_b200/libmp3gpu.so profiles/r02_sass_histogram.json

Reads `cuobjdump -sass` (works without a GPU)."""
import collections
import json
import re
import subprocess
import sys

CLASSES = [
    ("fp32_fma", r"^(FFMA|FFMA2)"), ("fp32_other", r"^(FADD|FMUL|FMNMX|FSEL|FSET|FSETP|FCHK|MUFU|F2I|I2F|F2F|FADD2|FMUL2)"),
    ("int_alu", r"^(IADD|IADD3|IMAD|LEA|LOP|LOP3|SHF|SHL|SHR|PRMT|SGXT|IABS|IMNMX|VIADD|VIMNMX|VIADDMNMX|POPC|FLO|BREV|BMSK|ISETP|SEL|PLOP3|R2P|P2R|MOV|CS2R|S2R|S2UR|UMOV|UIADD3|ULOP3|ULEA|USHF|UIMAD|R2UR|LDC|LDCU|UISETP|USEL|UPLOP3|REDUX)"),
    ("async_copy_LDGSTS", r"^(LDGSTS|LDGDEPBAR|DEPBAR)"), ("tma_UBLKCP_UTMA", r"^(UBLKCP|UTMALDG|UTMASTG|UTMAPF|SYNCS)"),
    ("shared_ld_st", r"^(LDS|STS|LDSM|ATOMS)"), ("global_ld", r"^(LDG|LD\.)"), ("global_st", r"^(STG|ST\.|RED|ATOMG|ATOM)"),
    ("tensor_UTC_HMMA", r"^(UTC|HMMA|IMMA|DMMA|LDTM|STTM|HGMMA)"), ("shuffle_vote", r"^(SHFL|VOTE|MATCH|VOTEU)"),
    ("branch_sync", r"^(BRA|BRX|JMP|EXIT|RET|CALL|BSSY|BSYNC|BAR|WARPSYNC|NANOSLEEP|YIELD|BPT|ERRBAR|MEMBAR|BMOV|BREAK)"),
]


def main():
    so, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?PT?\d*\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    report = {}
    for k, c in kernels.items():
        total = sum(c.values())
        cls = collections.OrderedDict((n, 0) for n, _ in CLASSES)
        other = 0
        for mn, cnt in c.items():
            for n, rx in CLASSES:
                if re.match(rx, mn):
                    cls[n] += cnt
                    break
            else:
                other += cnt
        cls["other"] = other
        report[k] = {"instructions": total, "code_bytes": total * 16, "classes": cls,
                     "top_mnemonics": dict(collections.Counter({m.split(".")[0]: 0 for m in c}) | collections.Counter()),}
        top = collections.Counter()
        for mn, cnt in c.items():
            top[mn.split(".")[0]] += cnt
        report[k]["top_mnemonics"] = dict(top.most_common(14))
    json.dump({"library": so, "note": "static SASS of sm_100a (cuobjdump -sass); 16 bytes per instruction; L1.5 instruction cache 32 KB",
               "kernels": report}, open(out, "w"), indent=1)
    for k, r in report.items():
        print(f"{k}: {r['instructions']} instr = {r['code_bytes'] / 1024:.1f} KB; FFMA {r['classes']['fp32_fma']}, LDGSTS {r['classes']['async_copy_LDGSTS']}, "
              f"TMA {r['classes']['tma_UBLKCP_UTMA']}, tensor {r['classes']['tensor_UTC_HMMA']}")


if __name__ == "__main__":
    main()
