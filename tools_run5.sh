python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/t5.log
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py $Q > gpurun_out/b5_c5_pf.json 2>gpurun_out/b5.err
MRS_B200_LIB=$PWD/build_variants/lib_nopf.so python bench.py $Q > gpurun_out/b5_c5_nopf.json 2>>gpurun_out/b5.err
python bench.py $Q > gpurun_out/b5_c5_pf2.json 2>>gpurun_out/b5.err
python bench.py --workload c4 $Q > gpurun_out/b5_c4.json 2>>gpurun_out/b5.err
python bench.py --workload c3 $Q > gpurun_out/b5_c3.json 2>>gpurun_out/b5.err
BA="--steps 20 --warmup 3 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py --workload c4 $BA > gpurun_out/plain5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1_c4b.csv python bench.py --workload c4 $BA > gpurun_out/ncu5.log 2>&1
python bench.py $BA > gpurun_out/plain5b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_group -s 30 -c 2 -o gpurun_out/prof_r1c python bench.py $BA > gpurun_out/ncu5b.log 2>&1
cat gpurun_out/t5.log; tail -3 gpurun_out/b5.err
for f in gpurun_out/b5_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))
"; done
