TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 4 2; do
$TR --nproc-per-node $N --master-port 2953$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_s_bench$N.json 2> gpurun_out/r02_s_bench$N.err; echo "rc $?"
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/r02_s_bench1.json 2> gpurun_out/r02_s_bench1.err
for f in gpurun_out/r02_s_bench*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('sm_reserve_backward'), d['config'].get('nccl_max_ctas'), round(d['e2e']['value'],1))"; done
grep -i "error\|warn" gpurun_out/r02_s_bench4.err | head -5
