#!/bin/bash
# One gpurun call: BA + SIFT parity tests, smoke, A/B of build variants.
tag=${1:-r3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ba.py tests/test_gpu_sift.py tests/test_gpu_golden.py -m gpu -x -q --timeout 600 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -6 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
{
  python tools/ab_value.py 40
  for v in openpano_b200/_variants/*.so; do PANO_B200_LIB=$v python tools/ab_value.py 40; done
} > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
