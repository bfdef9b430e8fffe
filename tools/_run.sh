python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/t14.log
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py $Q > gpurun_out/b14_big.json 2>gpurun_out/b14.err
MRS_B200_BIGCTA=0 python bench.py $Q > gpurun_out/b14_small.json 2>>gpurun_out/b14.err
python bench.py $Q > gpurun_out/b14_big2.json 2>>gpurun_out/b14.err
MRS_B200_PDL=0 python bench.py $Q > gpurun_out/b14_big_nopdl.json 2>>gpurun_out/b14.err
cat gpurun_out/t14.log; tail -3 gpurun_out/b14.err
for f in gpurun_out/b14_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))
"; done
MRS_B200_LIB=$PWD/build_variants/lib_trace.so python tools/trace_c5.py 2>&1 | tail -20
