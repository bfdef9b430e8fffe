python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/t40.log
python bench.py --steps 200 --warmup 5 > gpurun_out/b40_c5.json 2>gpurun_out/b40.err
python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/b40_ref.json 2>>gpurun_out/b40.err
for w in c2 c3 c4 c5v c5p; do python bench.py --workload $w --steps 200 --warmup 5 --clock-seconds 0.3 > gpurun_out/b40_$w.json 2>>gpurun_out/b40.err; done
BA="--steps 20 --warmup 3 --no-cpu --clock-seconds 0 --e2e-steps 2"
python bench.py $BA > gpurun_out/plain40.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_c5.csv python bench.py $BA > gpurun_out/ncu40.log 2>&1
python bench.py --workload c4 $BA > gpurun_out/plain40b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1_c4.csv python bench.py --workload c4 $BA > gpurun_out/ncu40b.log 2>&1
python bench.py $BA > gpurun_out/plain40c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_group -s 12 -c 2 -o gpurun_out/prof_r1_final python bench.py $BA > gpurun_out/ncu40c.log 2>&1
BW="--steps 100 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py $BW > gpurun_out/plain40d.log 2>&1 && ncu --set full --cache-control none --clock-control none -k regex:step_group -s 70 -c 2 -o gpurun_out/prof_r1_warm python bench.py $BW > gpurun_out/ncu40d.log 2>&1
cat gpurun_out/t40.log; tail -3 gpurun_out/b40.err
for f in gpurun_out/b40_c*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s | e2e %s | cpu %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value')))
"; done
