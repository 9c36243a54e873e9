TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=${1:-2}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
k=16
FNM_DP_SM_RESERVE=$k NCCL_MAX_CTAS=$k NCCL_MAX_NCHANNELS=$k NCCL_NVLS_NCHANNELS=$k $TR --nproc-per-node $N --master-port 29513 tools/dp_timeline.py gpurun_out/tl5_n${N}_k$k.txt 2>&1 | grep "^#" | head -2
python tools/dp_timeline.py gpurun_out/tl5_n1.txt 2>&1 | grep "^#" | head -2
