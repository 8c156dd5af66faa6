cd $GRAFT_REPO_ROOT
EIGB200_SSD_FORM=mma timeout 300 python -m pytest tests/test_scans_gpu.py -q -k "ssd" 2>&1 | tail -15
for f in scan mma; do
  EIGB200_SSD_FORM=$f timeout 120 python tools/kbench.py ssd_c2 --iters 20 2>&1 | tail -1
done
