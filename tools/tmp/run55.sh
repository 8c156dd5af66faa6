cd $GRAFT_REPO_ROOT
EIGB200_SSD_FORM=mma timeout 120 python tools/kbench.py ssd_c2 --iters 20 2>&1 | tail -1 | cut -c1-120
EIGB200_SSD_PB=64 EIGB200_SSD_FORM=mma timeout 120 python tools/kbench.py ssd_c2 --iters 20 2>&1 | tail -1 | cut -c1-120
