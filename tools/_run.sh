Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py $Q > gpurun_out/b23_base.json 2>gpurun_out/b23.err
for ns in 1000 2000 3500; do MRS_B200_LIB=$PWD/build_variants/lib_st$ns.so python bench.py $Q > gpurun_out/b23_st$ns.json 2>>gpurun_out/b23.err; done
tail -2 gpurun_out/b23.err
for f in gpurun_out/b23_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))
"; done
