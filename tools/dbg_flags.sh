for a in "" "--cin 64 --cout 64 --L 76800" "--cin 128 --cout 128 --L 19200" "--fir"; do echo "== $a"; W2S_DEBUG_FLAGS=64 python tools/profile_conv.py $a --iters 3 2>&1 | grep -E "blocked|CTA entry" | cut -c1-230; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
