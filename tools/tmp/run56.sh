cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for d in 0 2; do
  EIGB200_GEMM_DIRECT_STORE=$d timeout 200 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('DIRECT$d', d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
done
for d in 0 2; do
  EIGB200_GEMM_DIRECT_STORE=$d timeout 200 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('DIRECT$d', d['value'], d['ms_per_step'])"
done
