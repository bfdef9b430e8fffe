#!/bin/bash
# Round evidence on ONE B200 (run through gpurun): the GPU test suite, bench lines for every workload, the reference arm,
# the ncu launch lists and the two --set full captures of the group kernel.  tools/evidence_collect.py turns the output
# into the files under profiles/.   gpurun --timeout 2400 -- 'bash tools/evidence.sh'
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/ev_gputest.log 2>&1; echo "rc=$?" >> $O/ev_gputest.log
grep -E "^FAILED|passed|failed|rc=" $O/ev_gputest.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 > $O/ev_bench_c5.json 2> $O/ev.err || exit 1
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu > $O/ev_bench_c5_200.json 2>> $O/ev.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/ev_bench_reference_arm.json 2>> $O/ev.err
for w in c2 c3 c4 c5v c5p m64; do
  timeout 300 python bench.py --workload $w --no-cpu --steps 20 --warmup 5 > $O/ev_bench_$w.json 2>> $O/ev.err
done
Q="--steps 20 --warmup 3 --no-cpu --clock-seconds 0 --e2e-steps 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ev_launches_c5.csv python bench.py $Q > $O/ev_launches_c5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ev_launches_c4.csv python bench.py --workload c4 $Q > $O/ev_launches_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_group -s 4 -c 2 -f -o $O/ev_full_cold python tools/prof_c5.py > $O/ev_full_cold.log 2>&1
timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:step_group -s 3 -c 4 -f -o $O/ev_full_warm python tools/prof_c5.py > $O/ev_full_warm.log 2>&1
timeout 300 python tools/latency_c1.py > $O/ev_latency.txt 2>&1
timeout 600 python tools/soak.py c5 2000 > $O/ev_soak.txt 2>&1
tail -3 $O/ev.err
tail -2 $O/ev_soak.txt
ls -la $O/ev_* | awk '{print $5, $9}'
