#!/bin/bash
# N-GPU call (default 4): the sharded path on an UNEVEN split (38 images over 4 ranks), checked bit for bit
# against one GPU, linear and multiband; then the 2-GPU tests of the suite.
tag=${1:-run4}
NG=${2:-4}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 tools/run_dist.py unordered38 4 0 > gpurun_out/${tag}_rundist_u38.json 2> gpurun_out/${tag}_rundist_u38.err
echo "run_dist unordered38 exit $?"; tail -c 1500 gpurun_out/${tag}_rundist_u38.json; tail -3 gpurun_out/${tag}_rundist_u38.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29543 tools/run_dist.py small 2 3 > gpurun_out/${tag}_rundist_small_mb3.json 2> gpurun_out/${tag}_rundist_small_mb3.err
echo "run_dist small mb3 exit $?"; tail -c 800 gpurun_out/${tag}_rundist_small_mb3.json
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_configs.py::test_two_devices_in_one_process -m gpu -q --timeout 500 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${tag}_pytest.log
