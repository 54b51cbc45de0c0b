#!/bin/bash
for r in 1 2; do
  echo "default (proportional):"; python tools/time_masked_forward.py 2>&1 | tail -1
  echo "equal split:";            W2S_LIB_VARIANT=equal python tools/time_masked_forward.py 2>&1 | tail -1
  echo "no pairs:";               W2S_ENC_PAIRS=0 python tools/time_masked_forward.py 2>&1 | tail -1
done
