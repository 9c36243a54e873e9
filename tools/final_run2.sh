# final bench lines of the small configurations on the final code (full default lines: CPU baseline + stock-GPU comparator)
for w in cfg1 cfg2 cfg3 cfg4; do python bench.py --workload $w > gpurun_out/r02_final2_bench_$w.json 2> gpurun_out/r02_final2_bench_$w.err; done
python bench.py --no-cpu-baseline > gpurun_out/r02_final2_bench_cfg5.json 2> gpurun_out/r02_final2_bench_cfg5.err
for f in gpurun_out/r02_final2_bench_cfg*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value'],1), d['roofline']['kernel'], d['roofline']['frac'], (d.get('gpu_comparator') or {}).get('fp32',{}).get('ms_per_step'), (d.get('cpu_baseline') or {}).get('value'))"; done
