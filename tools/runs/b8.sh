TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=${1:-8}
$TR --nproc-per-node $N --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_p_bench${N}_k16.json 2> gpurun_out/r02_p_bench${N}_k16.err; tail -c 600 gpurun_out/r02_p_bench${N}_k16.json | head -c 300; echo
FNM_DP_SM_RESERVE=12 $TR --nproc-per-node $N --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_p_bench${N}_k12.json 2> gpurun_out/r02_p_bench${N}_k12.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-comparator > gpurun_out/r02_p_bench1.json 2> gpurun_out/r02_p_bench1.err
for f in gpurun_out/r02_p_bench*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('sm_reserve_backward'), d['config'].get('nccl_max_ctas'))"; done
