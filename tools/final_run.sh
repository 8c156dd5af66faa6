#!/bin/bash
# End-of-session evidence run on one B200 (under gpurun): full GPU test suite, the bench line, launch list under ncu, per-kernel micro-benchmarks,
# one bench line per BASELINE config with its CPU baseline.  usage: tools/final_run.sh <tag>   -> gpurun_out/<tag>_*
tag=${1:-final}
o=gpurun_out
timeout 300 python __graft_entry__.py --smoke > $o/${tag}_smoke.log 2>&1; tail -2 $o/${tag}_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $o/${tag}_tests.log 2>&1; tail -3 $o/${tag}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; tail -c 400 $o/${tag}_bench.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-other-configs --no-cpu-baseline --no-gpu-baseline --no-weak-curve > $o/${tag}_ncu1.log 2>&1
timeout 400 python tools/kbench.py front_c2 tail_c2 emb_c2 k1_c2 ssd_c2 diag_c3 k3_c4 lin_c5_glu_f16 lin_c5_in_f16 lin_c5_qkv_f16 lin_c5_out_f16 lin_c3_out ssd_c5 linattn_c5 linattn_c5_col linattn_c5_conv conv_c5 \
  --iters 20 --json $o/${tag}_kbench.json > $o/${tag}_kbench.log 2>&1; tail -3 $o/${tag}_kbench.log
for c in c1 c3-lru c3-s5 c4 c5-mamba c5-normattn; do
  timeout 400 python bench.py --config $c --steps 5 --warmup 3 > $o/${tag}_cfg_$c.json 2> $o/${tag}_cfg_$c.err; tail -c 200 $o/${tag}_cfg_$c.json; echo
done
