python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/t7.log
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py $Q > gpurun_out/b7_c5_pdl.json 2>gpurun_out/b7.err
MRS_B200_PDL=0 python bench.py $Q > gpurun_out/b7_c5_nopdl.json 2>>gpurun_out/b7.err
python bench.py $Q > gpurun_out/b7_c5_pdl2.json 2>>gpurun_out/b7.err
python bench.py --workload c3 $Q > gpurun_out/b7_c3_pdl.json 2>>gpurun_out/b7.err
MRS_B200_PDL=0 python bench.py --workload c3 $Q > gpurun_out/b7_c3_nopdl.json 2>>gpurun_out/b7.err
python bench.py --workload c2 $Q > gpurun_out/b7_c2_pdl.json 2>>gpurun_out/b7.err
python bench.py --workload c4 $Q > gpurun_out/b7_c4.json 2>>gpurun_out/b7.err
cat gpurun_out/t7.log; tail -3 gpurun_out/b7.err
for f in gpurun_out/b7_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))
"; done
