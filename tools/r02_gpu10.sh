set -u
O=gpurun_out/r02i; mkdir -p $O
for c in c2 c3 c4; do
  timeout 900 python bench.py --config $c --steps 10 --cycles 200 > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?"; head -c 1500 $O/bench_$c.json; echo; tail -3 $O/bench_$c.err
done
timeout 1200 python bench.py --config c5 --steps 2 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 rc=$?"; head -c 2500 $O/bench_c5.json; echo; tail -5 $O/bench_c5.err
timeout 600 python bench.py --impl reference --config c2 --steps 3 --warmup 1 > $O/ref_c2.json 2> $O/ref_c2.err; head -c 600 $O/ref_c2.json; echo
