set -u
O=gpurun_out/r02_8gpu; mkdir -p $O
run() { name=$1; n=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/$name.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","parity_checked","parity_errors")}, d["roofline"]["phase_ms_per_step"], 'e2e', d["e2e"]["value"])
except Exception as e:
    print("ERR", e); print(open("$O/$name.err").read()[-1200:])
PY
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_worker.py > $O/mgpu_worker_8.log 2>&1; tail -1 $O/mgpu_worker_8.log
run c2_n8 8 --config c2 --steps 20
run c5_n8 8 --config c5 --steps 3
run c5_n4 4 --config c5 --steps 3
run c4_n4 4 --config c4 --steps 20
run c3_n8 8 --config c3 --steps 20
run c2_n4 4 --config c2 --steps 20
