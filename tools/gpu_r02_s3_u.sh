#!/bin/bash
# Round 2, session 3: GPU suite on the final source, then compute-sanitizer memcheck over the kernels this session
# changed (direct-from-global transform, paired launches, warp-parallel live list): encoder conv kernel tests incl. masks
# and partial tiles, the paired-vs-single forward test, the smoke forward.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/u_tests.log 2>&1
echo "tests rc=$?"; tail -n 3 gpurun_out/u_tests.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider \
   -k "encoder_conv_layer or many_tiles_with_mask or row_mask_skips" > gpurun_out/u_memcheck_kernels.log 2>&1
echo "memcheck kernels rc=$?"; tail -n 3 gpurun_out/u_memcheck_kernels.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_forward_gpu.py -q -m gpu -p no:cacheprovider \
   -k "paired and smap1" > gpurun_out/u_memcheck_paired.log 2>&1
echo "memcheck paired rc=$?"; tail -n 3 gpurun_out/u_memcheck_paired.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?"; tail -n 2 gpurun_out/u_memcheck_smoke.log
