#!/bin/bash
# fixed cost of a stream-kernel launch: CTA-0 milestones and CTA exit spread on small layers (debug build)
for a in "--cin 64 --cout 64 --L 9600" "--cin 64 --cout 64 --L 38400" "--cin 128 --cout 128 --L 9600" "--cin 16 --cout 16 --stride 2 --L 153600" "--cin 32 --cout 32 --L 38400"; do
  echo "== profile_conv $a"
  W2S_LIB_VARIANT=dbg W2S_DEBUG_FLAGS=64 python tools/profile_conv.py $a --iters 5 2>&1 | cut -c1-400
done
W2S_LIB_VARIANT=ring timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -x -q -k eog -s 2>&1 | grep -E "EOG wide|passed|failed"
