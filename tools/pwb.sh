for cfg in "2 64" "2 48" "1 64" "1 48"; do set -- $cfg; echo "NLO=$1 NT=$2"; FNM_PW2_NLO=$1 FNM_PW2_NT=$2 timeout 100 python tools/bench_stage.py cfg5 10 2>&1 | sed -n '2,3p;5,5p;7,7p'; done
