# round-end evidence run on one B200: tests, bench lines of all five configs, ncu launch list, ncu --set full of the tcgen05 kernels
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/r02_final_tests.txt; cat gpurun_out/r02_final_tests.txt
python bench.py > gpurun_out/r02_final_bench_cfg5.json 2> gpurun_out/r02_final_bench_cfg5.err
for w in cfg1 cfg2 cfg3 cfg4; do python bench.py --workload $w > gpurun_out/r02_final_bench_$w.json 2> gpurun_out/r02_final_bench_$w.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_reference_cfg5.json 2> gpurun_out/r02_final_reference_cfg5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_final_launches_cfg5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-comparator --no-graph > gpurun_out/r02_final_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pw2_kernel|ydft_tma|tabgemm_tc|conv_wgrad_tma|gz_pass_tma' -s 8 -c 8 -o gpurun_out/r02_final_tc -f python tools/prof_kernels.py 2 > gpurun_out/r02_final_ncu_full.log 2>&1
for f in gpurun_out/r02_final_bench_cfg*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['roofline']['kernel'], d['roofline']['frac'], (d.get('gpu_comparator') or {}).get('fp32',{}).get('ms_per_step'), (d.get('cpu_baseline') or {}).get('value'))"; done
tail -c 400 gpurun_out/r02_final_reference_cfg5.json
