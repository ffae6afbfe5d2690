# 2-GPU check of the round-end build: sharded parity worker (both exchanges) + c2 / c5 / c4 bench lines at N = 2
set -u
O=gpurun_out/r02_2gpu; mkdir -p $O
run() { name=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/$name.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","parity_checked","parity_errors")}, d["roofline"]["phase_ms_per_step"], 'e2e', d["e2e"]["value"])
except Exception as e:
    print("ERR", e); print(open("$O/$name.err").read()[-1200:])
PY
}
timeout 500 python -m pytest tests/test_multi_gpu.py -q -m gpu > $O/mgpu_tests.log 2>&1; tail -2 $O/mgpu_tests.log
run c2_n2 2 --config c2 --steps 20
run c5_n2 2 --config c5 --steps 3
run c4_n2 2 --config c4 --steps 20
