python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['kernels']: print(k['kernel'], round(k['avg_ms'],3), k['launches_per_step'], round(k['share'],3), round(k['algo_GBps']))"
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
