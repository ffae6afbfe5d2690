# runs measure_configs c4 for every A/B library under scheduler-plugins_b200/lib/ab (same box, back to back, twice)
O=gpurun_out/${1:-ab}; mkdir -p $O
for rep in 1 2; do
for f in scheduler-plugins_b200/lib/ab/*.so; do
  n=$(basename $f .so)
  B200S_LIB=$PWD/$f timeout 300 python tools/measure_configs.py --configs c4 --steps 20 > $O/$n.$rep.log 2>&1
  python - <<PY
import json
for l in open("$O/$n.$rep.log"):
    if l.startswith("{"):
        d=json.loads(l)
        if "LeastAllocated" in d["plugin"] or "Balanced" in d["plugin"]:
            print("$n", $rep, d["plugin"].split("/")[1], d["out"], d["kernel_ms"])
PY
done
done
