cd $GRAFT_REPO_ROOT
for lib in "" ab/lib_abl_epi.so ab/lib_abl_conv.so ab/lib_abl_both.so; do
  echo "=== LIB ${lib:-default}"
  if [ -n "$lib" ]; then export EIGB200_LIB=$PWD/$lib; fi
  timeout 200 python tools/kbench.py lin_in_tc3 lin_in_tc1 lin_out_tc3 lin_out_tc1 lin_glu_tc3 lin_glu_tc1 lin_out_none_tc3 --iters 20 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('%-20s %.3f ms  %.0f GB/s' % (d['case'], d['ms_median'], d['GBps']))
"
done
