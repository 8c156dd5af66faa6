#!/usr/bin/env python
"""Write the SASS of the hot kernels of libeigb200.so to profiles/sass/<kernel>.sass (cuobjdump -sass, encodings stripped) plus a mnemonic summary that
shows which Blackwell paths each kernel uses (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, SYNCS = mbarrier).
Usage: python tools/dump_sass.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "task-level-insights-from-eigenvalues-across-sequence-models_b200", "libeigb200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
# (file stem, regex on the demangled kernel name, pick)
WANT = [("k1_row8_kernel", r"k1_row8_kernel<"), ("k1_partials_kernel", r"k1_partials_kernel<8>"), ("k1_gate_kernel", r"k1_gate_kernel"),
        ("ssd_scan_v3_kernel_c2", r"ssd_scan_v3_kernel<16, 64, 8, true, true>"), ("ssd_tc_kernel", r"ssd_tc_kernel"),
        ("diag_scan_kernel", r"diag_scan_kernel"), ("eigvals_kernel", r"eigvals_kernel"), ("dplr_abar_kernel", r"dplr_abar"),
        ("gemm_tc_ts_kernel_inproj_f16", r"gemm_tc_ts_kernel<0, false, 4, true>"),
        ("gemm_tc_ts_kernel_glu_tf32", r"gemm_tc_ts_kernel<2, false, 4, false>"),
        ("gemm_out_glu_kernel", r"gemm_out_glu_kernel"), ("mamba_front_kernel", r"mamba_front_kernel<true>"), ("gemm_tc_stream_kernel", r"gemm_tc_stream_kernel<0, false"),
        ("gemm_tc_stream_kernel_glu_f16_rawa", r"gemm_tc_stream_kernel<2, true, true>"), ("linattn_mma_kernel", r"linattn_mma_kernel<false>"),
        ("linattn_mma_kernel_conv", r"linattn_mma_kernel<true>"), ("diag_scan_kernel_pipelined", r"diag_scan_kernel<1, 0, 32, true>"),
        ("embedding128_kernel", r"embedding128_kernel"),
        ("embedding_kernel", r"embedding_kernel<8>"), ("count_moments_kernel", r"count_moments_kernel"), ("ratio_hist_kernel", r"ratio_hist_kernel"),
        ("linattn_forward_col_kernel", r"linattn_forward_col_kernel<64>"), ("softmax_nu_kernel", r"softmax_nu_kernel")]
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "ELECT", "LDGSTS", "HMMA", "FFMA", "DFMA", "MUFU", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "SHFL"]


def main():
    os.makedirs(OUT, exist_ok=True)
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    dem = subprocess.run(["c++filt"], input=txt, capture_output=True, text=True).stdout
    blocks = re.split(r"(?m)^\s*Function : ", dem)[1:]
    summary = []
    for stem, rx in WANT:
        hit = [b for b in blocks if re.search(rx, b.split("\n", 1)[0])]
        if not hit:
            summary.append("%-34s NOT FOUND (%s)" % (stem, rx)); continue
        b = hit[0]
        name, body = b.split("\n", 1)
        lines = []
        for ln in body.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", ln)
            if m:
                lines.append("/*%s*/ %s" % (m.group(1), m.group(2).rstrip(" ;") + " ;"))
            elif re.match(r"\s+\.L_x_\d+:", ln):
                lines.append(ln.strip())
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", l.split("*/ ", 1)[1]).split(".")[0].split(" ")[0] for l in lines if l.startswith("/*"))
        with open(os.path.join(OUT, stem + ".sass"), "w") as f:
            f.write("// %s\n// cuobjdump -sass libeigb200.so (sm_100a), instruction encodings stripped; %d instructions\n" % (name.strip(), sum(ops.values())))
            f.write("\n".join(lines) + "\n")
        keys = " ".join("%s=%d" % (k, sum(v for o, v in ops.items() if o.startswith(k))) for k in KEY if any(o.startswith(k) for o in ops))
        summary.append("%-34s %5d instr  %s" % (stem, sum(ops.values()), keys))
    with open(os.path.join(OUT, "SUMMARY.txt"), "w") as f:
        f.write("SASS of the hot kernels (tools/dump_sass.py).  Mnemonics: UTCHMMA = tcgen05.mma (kind::f16 / tf32), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,\n"
                "UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA), SYNCS = mbarrier, LDGSTS = cp.async.\n\n" + "\n".join(summary) + "\n")
    print("\n".join(summary))


if __name__ == "__main__":
    main()
