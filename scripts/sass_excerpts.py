"""Static SASS evidence of the shipped library: python scripts/sass_excerpts.py > profiles/r02b_sass_excerpts.txt"""
import re, subprocess, collections, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "sngnn_b200", "lib", "libsng.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = [("simknn_stage1_kernelILi4ELb0", "K1 main pass, small K (split mode)"), ("simknn_stage1_kernelILi1ELb0", "K1, large K / streamed query block"),
        ("simknn_stage1_kernelILi4ELb1", "K1 seed pass"), ("gemm_nt_f16_kernel", "toolbox N x N producer"),
        ("edge_fwd_staged_kernelILi8ELb0ELb0ELb0", "K2, C = 32"), ("edge_fwd_staged_kernelILi8ELb1ELb0ELb0", "K2 + K4 fused"),
        ("edge_fwd_narrow_kernelILb1ELb0", "narrow rows, fused forward"), ("edge_bwd_target_staged_kernelILi8", "K2b pass T"),
        ("edge_bwd_source_staged_kernelILi8ELb1ELb0", "K2b pass S, fused"), ("edge_bwd_source_narrow_kernelILb1", "narrow rows, pass S"),
        ("lin_bwd_partial_kernelILi1ELb1", "lin backward, flat staging"), ("lin_norm_kernelILi8", "lin + bias + 1/norm")]
ops = ["UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "SYNCS", "LDGSTS", "CREDUX", "LDS", "STS", "SHFL", "FFMA", "FMNMX3", "MATCH", "REDUX"]
funcs = re.split(r"\n\s*Function : ", out)
print("SASS evidence (cuobjdump -sass sngnn_b200/lib/libsng.so, sm_100a), final state of round 2.  Counts are static instruction counts per kernel.")
print("tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2), TMA -> UTMALDG, tcgen05.commit -> UTCBAR, tcgen05.ld -> LDTM, cp.async -> LDGSTS, warp max -> CREDUX.\n")
for key, label in want:
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        if key in name:
            cnt = collections.Counter()
            first = {}
            for line in f.splitlines():
                m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Za-z0-9_.]+)?\s", line)
                if m and m.group(1) in ops:
                    cnt[m.group(1)] += 1
                    first.setdefault(m.group(1), re.sub(r"\s+/\*[0-9a-fx]+\*/\s*$", "", line.split("*/", 1)[1]).strip())
            print(f"== {label}\n   mangled: {name}\n   " + ", ".join(f"{k}: {cnt[k]}" for k in ops if cnt[k]))
            for k in ("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "LDGSTS", "CREDUX"):
                if k in first:
                    print("      " + first[k])
            print()
            break
