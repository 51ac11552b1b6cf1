#!/bin/bash
# One gpurun call: parity of the SIFT stages with the quad-design descriptor / orientation
# kernels, a memcheck of the smoke pass, then A/B step timings (first designs via PANO_*_V1,
# build variants via PANO_B200_LIB), optionally an ncu capture of the two kernels.
tag=${1:-ab}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sift.py tests/test_gpu_golden.py -m gpu -x -q --timeout 600 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -4 gpurun_out/${tag}_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck exit $?"; tail -3 gpurun_out/${tag}_memcheck.log
{
  python tools/ab_value.py 40
  PANO_DESC_V1=1 PANO_ORI_V1=1 python tools/ab_value.py 40
  for v in openpano_b200/_variants/*.so; do PANO_B200_LIB=$v python tools/ab_value.py 40; done
} > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_descriptor|k_orientation" --launch-skip 2 -c 2 -o gpurun_out/${tag}_desc python tools/one_step.py 2 > gpurun_out/${tag}_desc.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/${tag}_desc.log
