python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/t4.log
python bench.py --steps 200 --warmup 5 > gpurun_out/b4_c5.json 2>gpurun_out/b4.err
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
MRS_B200_LIB=$PWD/build_variants/lib_mb8.so python bench.py $Q > gpurun_out/b4_c5_mb8.json 2>>gpurun_out/b4.err
for w in c2 c3 c4; do python bench.py --workload $w --steps 200 --warmup 5 --clock-seconds 0.3 > gpurun_out/b4_$w.json 2>>gpurun_out/b4.err; done
BA="--steps 20 --warmup 3 --no-cpu --clock-seconds 0 --e2e-steps 2"
python bench.py $BA > gpurun_out/plain4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py $BA > gpurun_out/ncu4.log 2>&1
python bench.py --workload c4 $BA > gpurun_out/plain4b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_c4.csv python bench.py --workload c4 $BA > gpurun_out/ncu4b.log 2>&1
cat gpurun_out/t4.log; tail -3 gpurun_out/b4.err
for f in gpurun_out/b4_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s | e2e %s | cpu %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value')))
"; done
