N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/ev_bench_c5_${N}gpu.json 2> gpurun_out/ev_mg_${N}.err
$TR bench.py --gpus $N --steps 200 --warmup 5 --scaling strong --e2e-steps 0 > gpurun_out/ev_bench_c5_${N}gpu_strong.json 2>> gpurun_out/ev_mg_${N}.err
tail -2 gpurun_out/ev_mg_${N}.err
for f in gpurun_out/ev_bench_c5_${N}gpu.json gpurun_out/ev_bench_c5_${N}gpu_strong.json; do python -c "
import json
d=json.load(open('$f'))
print('$f', 'value %.3e ms/step %.4f frac %.3f scaling %s e2e %s clocks %s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['scaling'],d['e2e']['value'],d['clocks']))"; done
