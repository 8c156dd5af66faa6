cd $GRAFT_REPO_ROOT
EIGB200_SSD_FORM=mma timeout 300 ncu --set full --import-source on --clock-control none -k regex:ssd_chunk_mma -s 3 -c 1 -o gpurun_out/ssd_mma python tools/kbench.py ssd_c2 --iters 3 > gpurun_out/ncu_ssd_mma.log 2>&1
tail -3 gpurun_out/ncu_ssd_mma.log
