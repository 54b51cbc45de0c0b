python tools/time_wide.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_forward_gpu.py -q -m gpu -s -p no:cacheprovider -k "eog" 2>&1 | grep -E "max-abs|passed|failed" | cut -c1-150
