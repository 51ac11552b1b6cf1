#!/bin/bash
# N-GPU call (default 2): sharded path tests + the bench line at N (with the `sharded` object).
# usage: tools/gpu_round2.sh <tag> [N]
tag=${1:-run2}
NG=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_configs.py::test_two_devices_in_one_process -m gpu -q --timeout 600 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tools/run_dist.py unordered38 4 0 > gpurun_out/${tag}_rundist.json 2> gpurun_out/${tag}_rundist.err
echo "run_dist exit $?"; tail -2 gpurun_out/${tag}_rundist.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 30 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${tag}_bench.err; python -c "
import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print(d['value'],d['e2e']['value']);print(json.dumps(d['sharded'],indent=1))"
