#!/bin/bash
# Build an A/B variant of the library next to the default one: tools/ab_build.sh <name> "<nvcc flags>"
# then run anything with W2S_LIB_VARIANT=<name>.
W2S_LIB_VARIANT=$1 W2S_NVCC_FLAGS="$2" python -c "
import sys; sys.path.insert(0, '.')
from wav2sleep_b200 import _lib
print(_lib.build(force=True))"
