set -u
O=gpurun_out/r02m; mkdir -p $O
run() { name=$1; n=$2; shift 2
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/$name.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","parity_checked","parity_errors")}, d["roofline"]["phase_ms_per_step"], d["roofline"]["per_plugin_kernel_ms"], 'frac', d["roofline"]["frac"], 'e2e', d["e2e"]["value"])
except Exception as e:
    print("ERR", e); print(open("$O/$name.err").read()[-1500:])
PY
}
timeout 900 python -m pytest tests/test_multi_gpu.py -q -m gpu 2>&1 | tail -3
run c2_n2 2 --config c2 --steps 20
run c2_n2_b 2 --config c2 --steps 50
python bench.py --config c2 --steps 20 --cycles 100 > $O/c2_n1.json 2>$O/c2_n1.err; python -c "
import json; d=json.loads(open('$O/c2_n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'], d['parity_checked'], d['parity_errors'])"
