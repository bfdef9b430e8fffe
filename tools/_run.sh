for args in "--steps 7 --warmup 2" "--steps 3 --warmup 1" "--steps 101 --warmup 3" "--steps 250 --warmup 5" "--steps 200 --warmup 5"; do
python bench.py $args --no-cpu > gpurun_out/b74.json 2>gpurun_out/b74.err || { echo FAIL $args; tail -5 gpurun_out/b74.err; }
python -c "
import json
d=json.load(open('gpurun_out/b74.json'))
print('$args', 'value %.3e ms/step %.4f steps %d warmup %d launches %d e2e %s launch %s'%(d['value'],d['ms_per_step'],d['steps'],d['warmup'],d['gpu_launches'],d['e2e']['value'], d['config']['launch']))"
done
