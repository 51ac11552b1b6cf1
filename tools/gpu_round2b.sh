#!/bin/bash
# 2-GPU call: new SIFT window test + smoke (early signal), run_dist on config 3, the bench line at N=2.
tag=${1:-run2b}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sift.py -m gpu -q -x --timeout 200 -k "wide or other" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/run_dist.py unordered38 4 0 > gpurun_out/${tag}_rundist_u38.json 2> gpurun_out/${tag}_rundist_u38.err
echo "run_dist exit $?"; tail -c 1300 gpurun_out/${tag}_rundist_u38.json
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${tag}_bench.err; python -c "
import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print(d['value'],d['e2e']['value']);print(json.dumps(d['sharded'],indent=1))"
