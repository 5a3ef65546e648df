#!/bin/bash
# One-GPU measurement pass of a round (run on the GPU box through gpurun): bench lines, ncu launch list, full captures.
#   tools/measure_round.sh <tag> [quick]     -> gpurun_out/<tag>_*.{json,csv,ncu-rep}   (quick: without the default bench and the reference arm)
# Every ncu pass runs after the plain command has exited 0; numbers printed under ncu are never bench values.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
if [ "$2" != quick ]; then
python bench.py > $out/${tag}_bench_default_c3.json 2> $out/${tag}_bench_default_c3.err || echo "bench failed"
echo "default bench: $(( $(date +%s) - t0 )) s"
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm_c3.json 2> $out/${tag}_bench_reference_arm_c3.err || echo "reference arm failed"
fi
for w in C1 C2 C5; do
  python bench.py --workload $w --no-dropin > $out/${tag}_bench_${w,,}.json 2> $out/${tag}_bench_${w,,}.err || echo "bench $w failed"
done
echo "benches: $(( $(date +%s) - t0 )) s"
small="bench.py --steps 2 --warmup 1 --no-sub --no-dropin --no-cpu-baseline --no-e2e"
python $small > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c3.csv python $small > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gravity_sym|k_spring_tiles" --launch-skip 6 -c 3 -o $out/${tag}_full_c3 -f python $small > $out/${tag}_ncu_full.log 2>&1
ncu -i $out/${tag}_full_c3.ncu-rep --page raw --csv > $out/${tag}_ncu_full_gravity_spring_tiles_c3.csv 2>/dev/null
echo "ncu: $(( $(date +%s) - t0 )) s"
