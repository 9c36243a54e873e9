timeout 120 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k pointwise 2>&1 | tail -2
for cfg in "1 48" "2 48" "2 64"; do set -- $cfg; echo "NLO=$1 NT=$2"; FNM_PW2_NLO=$1 FNM_PW2_NT=$2 timeout 100 python tools/bench_stage.py cfg5 10 2>&1 | sed -n '2,3p;5,5p;7,7p'; done
