#!/bin/bash
# Per-role wait accounting of the streaming conv kernel family (W2S_DEBUG_FLAGS=64, see conv_stream.cuh):
# which stage each kernel is waiting for, and the spread of CTA finish times.
for a in "--fir" "" "--stride 2" "--cin 32 --cout 32 --L 307200" "--cin 64 --cout 64 --L 76800" "--cin 128 --cout 128 --L 19200"; do
  echo "== profile_conv $a"
  W2S_DEBUG_FLAGS=64 python tools/profile_conv.py $a --iters 3 2>&1 | grep -E "blocked|CTA entry" | cut -c1-230
done
