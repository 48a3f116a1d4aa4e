#!/usr/bin/env python
"""SASS evidence for the hot kernels of the built objects (runs here, no GPU): per kernel the counts of the mnemonics
that prove the tcgen05 / TMEM / bulk-copy path, and a short excerpt around the first tensor-core instruction chain.

    python tools/sass_excerpt.py > profiles/r02_sass_excerpts.txt
"""
import collections
import os
import re
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "attentive_dfprior_b200", "csrc")
KERNELS = [("dfp_dec_fwd_tc.o", "k_decoder_fwd_tcILi2ELi0"), ("dfp_dec_fwd_tc.o", "k_decoder_fwd_tc_allILi0"),
           ("dfp_dec_dh_tc.o", "k_decoder_dh_tcILi2ELi0"), ("dfp_dec_dw_tc.o", "k_decoder_dw_tcILi2"),
           ("dfp_attn_tc.o", "k_attn_fwd_tc"), ("dfp_attn_tc.o", "k_attn_dh_tc"), ("dfp_attn_tc.o", "k_attn_dw_tc")]
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "REDG", "RED.", "MUFU.SIN", "MUFU.COS", "ELECT", "R2UR", "BAR.SYNC", "BAR.ARV",
        "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "F2FP", "UTMALDG", "FENCE.VIEW.ASYNC"]


def main():
    for obj, pat in KERNELS:
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
        funcs = re.split(r"\n\s*Function : ", out)
        for f in funcs[1:]:
            name = f.split("\n", 1)[0].strip()
            if pat not in name:
                continue
            lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,5}\*/", l)]
            ins = [re.sub(r"/\*.*?\*/", "", l).strip().rstrip(";").strip() for l in lines]
            ins = [i for i in ins if i]
            cnt = collections.Counter()
            for i in ins:
                for k in KEYS:
                    if k in i:
                        cnt[k] += 1
            print("=" * 110)
            print(name, "--", len(ins), "SASS instructions")
            print("  " + "  ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k]))
            first = next((n for n, i in enumerate(ins) if "UTCHMMA" in i), None)
            if first is not None:
                print("  excerpt around the first tensor-core chain:")
                for i in ins[max(0, first - 6):first + 14]:
                    print("      " + i)
            break


if __name__ == "__main__":
    main()
