MRS_B200_LIB=$PWD/build_variants/lib_tma.so python -m pytest tests -m gpu -q -x 2>&1 | tail -6
Q="--steps 200 --warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py $Q > gpurun_out/b29_base.json 2>gpurun_out/b29.err
MRS_B200_LIB=$PWD/build_variants/lib_tma.so python bench.py $Q > gpurun_out/b29_tma.json 2>>gpurun_out/b29.err
MRS_B200_LIB=$PWD/build_variants/lib_tma.so python bench.py --workload c3 $Q > gpurun_out/b29_tma_c3.json 2>>gpurun_out/b29.err
python bench.py --workload c3 $Q > gpurun_out/b29_base_c3.json 2>>gpurun_out/b29.err
tail -3 gpurun_out/b29.err
for f in gpurun_out/b29_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))
"; done
