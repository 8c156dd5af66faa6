cd $GRAFT_REPO_ROOT
for lib in ab/lib_base.so ""; do
  if [ -n "$lib" ]; then export EIGB200_LIB=$PWD/$lib; else unset EIGB200_LIB; fi
  echo "== ${lib:-new}"
  timeout 200 python tools/kbench.py lin_in_ln_tc3 lin_out_tc3 lin_glu_tc3 --iters 30 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('%-20s %.3f ms  %.0f GB/s' % (d['case'], d['ms_median'], d['GBps']))
"
done
unset EIGB200_LIB
timeout 300 python -m pytest tests/test_blocks_gpu.py -x -q -k "linear or gemm or tc or glu" 2>&1 | tail -2
timeout 200 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
