cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_ts -s 14 -c 3 -o gpurun_out/gemm3 python bench.py --steps 1 --warmup 1 --no-graph > gpurun_out/ncu_gemm3.log 2>&1
tail -2 gpurun_out/ncu_gemm3.log
