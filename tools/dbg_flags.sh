W2S_DEBUG_FLAGS=64 python tools/profile_conv.py --cin 64 --cout 64 --L 76800 --iters 3 2>&1 | grep -E "blocked|CTA entry" | cut -c1-230
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
