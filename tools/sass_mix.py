"""Static instruction mix of the hot kernels (cuobjdump -sass on the built library): how many SASS instructions of each class
a kernel's code holds — tensor-pipe issue (UTC*MMA), TMEM traffic (LDTM / STTM), bulk copies (UBLKCP), barriers (SYNCS, BAR),
shared-memory loads / stores, conversions (F2F / F2FP: the fp32 -> bf16 hi / lo split of every epilogue), FP32 math.
STATIC counts (loop bodies count once): they show what the CUDA-core side of a chain kernel is made of, not its run time.
Runs anywhere: `python tools/sass_mix.py > profiles/sass_r2_mix_hot_kernels.txt`."""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nsdp_b200 import _lib  # noqa: E402

HOT = ("vattn_bwd_oh_kernel", "vattn_fwd_oh_kernel", "dw_tc_kernel", "resnet_tail_bwd_tc_kernel", "resnet_tail_tc_kernel",
       "vattn_bwd_tc_kernel", "vattn_fwd_tc_kernel", "fused_mlp_tc_kernel", "fused_mlp_bwd_tc_kernel", "knn_scan_kernel",
       "fps_kernel")
CLASSES = [
    ("tensor issue", r"^UTC\w*MMA"), ("tensor commit", r"^UTCBAR"), ("TMEM ld", r"^LDTM"), ("TMEM st", r"^STTM"),
    ("bulk copy", r"^UBLKCP|^UTMA"), ("mbarrier", r"^SYNCS"), ("CTA barrier", r"^BAR|^UCGABAR"),
    ("smem ld", r"^LDS"), ("smem st", r"^STS"), ("global ld", r"^LDG"), ("global st", r"^STG"), ("atomics", r"^RED[^U]|^REDG|^ATOM"),
    ("convert", r"^F2F|^F2FP|^I2F|^F2I|^PRMT"), ("fp32 math", r"^FFMA|^FMUL|^FADD|^FMNMX|^FSEL|^FSETP|^MUFU"),
    ("int / addr / pred", r"^IMAD|^IADD|^VIADD|^LEA|^LOP|^PLOP|^SHF|^ISETP|^UISETP|^MOV|^UMOV|^UIADD|^ULEA|^ULOP|^USHF|^UIMAD|^S2R"
                          r"|^S2UR|^R2UR|^SEL|^USEL|^CS2R|^LDC|^LDCU|^ULDC|^IMNMX|^VIMNMX|^P2R|^R2P|^IABS|^FLO|^POPC|^BREV|^UFLO|^UPOPC"),
    ("fence / nop", r"^NOP|^FENCE|^MEMBAR|^YIELD|^DEPBAR|^ERRBAR|^CCTL|^BPT|^NANOSLEEP|^ENDCOLLECTIVE"),
    ("shuffle / vote", r"^SHFL|^VOTE|^REDUX|^ELECT|^MATCH"), ("branch", r"^BRA|^BSSY|^BSYNC|^EXIT|^CALL|^RET|^WARPSYNC"),
]
txt = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
print(f"# {os.path.basename(_lib.LIB_PATH)}: static SASS instruction mix of the hot kernels (sm_100a)")
print("# " + "kernel".ljust(78) + " total " + " ".join(c[0].replace(' ', '_') for c in CLASSES) + " other")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    if not any(h in name for h in HOT):
        continue
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem).replace("void ", "").replace("nsdp::", "")
    ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", f)
    cnt = collections.Counter()
    for op in ops:
        for cname, pat in CLASSES:
            if re.match(pat, op):
                cnt[cname] += 1
                break
        else:
            cnt["other"] += 1
    total = sum(cnt.values())
    print(dem[:78].ljust(80) + f"{total:6d} " + " ".join(f"{cnt[c[0]]:5d}" for c in CLASSES) + f" {cnt['other']:5d}")
