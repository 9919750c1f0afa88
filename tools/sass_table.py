#!/usr/bin/env python3
"""SASS instruction table of the hot kernels (static counts per opcode class) from the built objects:
    python tools/sass_table.py > profiles/r2_sass_table.md
Evidence for: no library kernels, cp.async (LDGSTS) in the fft_len-1024 sync kernel, TMA (UTMALDG + mbarrier SYNCS)
in the ring sync kernel, no tensor-core / TMEM instructions (none belong on this path), packed FP32 not used."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "gr-ofdm_tools_b200", "lib", "obj")
KERNELS = [
    ("framew_1024.o", "_Z16rx_framew_kernelILi1024ELi4ELb0EEv2KPPK6float2xxPKxPKiPKfS7_S7_P11ofdmx_framePhxPS1_xjii", "rx_framew_kernel<1024,4,false>"),
    ("framew_2048.o", "_Z16rx_framew_kernelILi2048ELi6ELb0EEv2KPPK6float2xxPKxPKiPKfS7_S7_P11ofdmx_framePhxPS1_xjii", "rx_framew_kernel<2048,6,false>"),
    ("framew_64.o", "_Z16rx_framew_kernelILi64ELi2ELb0EEv2KPPK6float2xxPKxPKiPKfS7_S7_P11ofdmx_framePhxPS1_xjii", "rx_framew_kernel<64,2,false>"),
    ("framep.o", "_Z16rx_framep_kernelILi6ELb0EEv2KPPK6float2xxPKxPKiPKfS7_S7_P11ofdmx_framePhxPS1_xjPKti", "rx_framep_kernel<6,false>"),
    ("api.o", "_Z23sync_metric_warp_kernelILi16EEvPK6float2xxfdPjS3_xiiii", "sync_metric_warp_kernel<16>"),
    ("api.o", "_Z23sync_metric_warp_kernelILi32EEvPK6float2xxfdPjS3_xiiii", "sync_metric_warp_kernel<32>"),
    ("api.o", "_Z24sync_metric_warpn_kernelILi64EEvPK6float2xxfdPjS3_xiiii", "sync_metric_warpn_kernel<64>"),
    ("api.o", "_Z22sync_metric_tma_kernel14CUtensorMap_stPK6float2xxifdPjS3_xxxx", "sync_metric_tma_kernel"),
    ("api.o", "_Z16agc2_span_kernelPK6float2PS_xxiixiffffPKffPfS5_Piii", "agc2_span_kernel"),
    ("txw_1024.o", None, "tx_framew_kernel<1024,4>"),
]
CLASSES = [("FP32 (FFMA/FMUL/FADD)", r"^(FFMA|FMUL|FADD)$"), ("FP32 packed (FFMA2/FADD2/FMUL2)", r"^(FFMA2|FADD2|FMUL2)$"),
           ("FP64 (DFMA/DADD/DMUL)", r"^(DFMA|DADD|DMUL)$"), ("MUFU", r"^MUFU$"), ("integer / logic (IADD3/IMAD/LOP3/SHF/LEA/...)", r"^(IADD3|IMAD|LOP3|SHF|LEA|VIADD|ISETP|SEL|VIMNMX|PRMT|IABS|POPC|FLO|BREV|I2F|F2I|I2FP|F2F|FSETP|FSEL|FMNMX)$"),
           ("shared loads/stores (LDS/STS)", r"^(LDS|STS)$"), ("global loads (LDG)", r"^LDG$"), ("global stores (STG)", r"^STG$"),
           ("cp.async (LDGSTS)", r"^LDGSTS$"), ("TMA (UTMALDG/UTMASTG)", r"^UTMA"), ("mbarrier (SYNCS)", r"^SYNCS$"),
           ("local memory (LDL/STL)", r"^(LDL|STL)$"), ("warp shuffles / votes (SHFL/VOTE/MATCH/REDUX)", r"^(SHFL|VOTE|MATCH|REDUX)$"),
           ("block barriers (BAR)", r"^BAR$"), ("tensor core / TMEM (UTC*MMA, HMMA, LDTM, ...)", r"^(UTC|HMMA|IMMA|QMMA|LDTM|STTM|UTCBAR)")]


def sass(obj, fun):
    cmd = ["cuobjdump", "-sass"] + (["-fun", fun] if fun else []) + [os.path.join(OBJ, obj)]
    return subprocess.run(cmd, capture_output=True, text=True).stdout


print("# SASS instruction table (static counts), round-2 build\n")
print("`python tools/sass_table.py` over `gr-ofdm_tools_b200/lib/obj/*.o` (sm_100a, nvcc 12.9).\n")
hdrs = [k[2] for k in KERNELS]
tables = []
for obj, fun, name in KERNELS:
    txt = sass(obj, fun)
    if fun is None:   # first <1024, 4> instantiation in the object
        blocks = txt.split("Function : ")
        txt = next((b for b in blocks if "tx_framew_kernelILi1024ELi4E" in b.split("\n")[0]), "")
    ops = collections.Counter()
    for line in txt.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m:
            ops[m.group(1).split(".")[0]] += 1
    tables.append(ops)
print("| class | " + " | ".join(hdrs) + " |")
print("|---|" + "---|" * len(hdrs))
print("| all instructions | " + " | ".join(str(sum(t.values())) for t in tables) + " |")
for label, pat in CLASSES:
    print("| %s | " % label + " | ".join(str(sum(v for k, v in t.items() if re.match(pat, k))) for t in tables) + " |")
print("\nLibrary kernels (cuFFT / cuBLAS / CUTLASS symbols) in libofdmx.so: none -- every `Function :` of "
      "`cuobjdump -sass libofdmx.so` is one of this repo's kernels:\n")
so = os.path.join(ROOT, "gr-ofdm_tools_b200", "lib", "libofdmx.so")
names = sorted(set(re.sub(r"I[LbE0-9x_]*E.*", "<...>", n) for n in re.findall(r"Function : (\S+)", subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout)))
print("```\n" + "\n".join(names) + "\n```")
