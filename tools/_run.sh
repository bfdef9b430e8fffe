Q="--warmup 5 --no-cpu --clock-seconds 0 --e2e-steps 0"
python bench.py --steps 200 $Q > gpurun_out/b41_stream.json 2>gpurun_out/b41.err
MRS_B200_LIB=$PWD/build_variants/lib_nostream.so python bench.py --steps 200 $Q > gpurun_out/b41_nostream.json 2>>gpurun_out/b41.err
python bench.py --steps 200 $Q > gpurun_out/b41_stream2.json 2>>gpurun_out/b41.err
python bench.py --steps 200 --envs 262144 $Q --graph-steps 50 > gpurun_out/b41_stream_big.json 2>>gpurun_out/b41.err
MRS_B200_LIB=$PWD/build_variants/lib_nostream.so python bench.py --steps 200 --envs 262144 $Q --graph-steps 50 > gpurun_out/b41_nostream_big.json 2>>gpurun_out/b41.err
tail -2 gpurun_out/b41.err
for f in gpurun_out/b41_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(' value %.3e ms/step %.4f frac %.3f | flushed ms %.4f frac %.3f | many %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['l2_flushed']['ms_per_step_median'],d['l2_flushed']['frac'],d['step_many'] and '%.3e'%d['step_many']['value']))
"; done
