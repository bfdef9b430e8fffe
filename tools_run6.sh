python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/t6.log
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py $Q > gpurun_out/b6_c5.json 2>gpurun_out/b6.err
MRS_B200_LIB=$PWD/build_variants/lib_nopf.so python bench.py $Q > gpurun_out/b6_c5_old_nopf.json 2>>gpurun_out/b6.err
python bench.py $Q > gpurun_out/b6_c5_b.json 2>>gpurun_out/b6.err
python bench.py --workload c4 $Q > gpurun_out/b6_c4.json 2>>gpurun_out/b6.err
python bench.py --workload c3 $Q > gpurun_out/b6_c3.json 2>>gpurun_out/b6.err
python bench.py --workload c2 $Q > gpurun_out/b6_c2.json 2>>gpurun_out/b6.err
BA="--steps 20 --warmup 3 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py $BA > gpurun_out/plain6b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_group -s 30 -c 2 -o gpurun_out/prof_r1d python bench.py $BA > gpurun_out/ncu6b.log 2>&1
cat gpurun_out/t6.log; tail -3 gpurun_out/b6.err
for f in gpurun_out/b6_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))
"; done
