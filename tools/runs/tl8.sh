TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
N=${1:-8}
port=29520
for k in 12 8 16; do
  port=$((port+1))
  FNM_DP_SM_RESERVE=$k NCCL_MAX_CTAS=$k NCCL_MAX_NCHANNELS=$k NCCL_NVLS_NCHANNELS=$k NCCL_DEBUG=INFO $TR --nproc-per-node $N --master-port $port tools/dp_timeline.py gpurun_out/tl4_n${N}_k$k.txt > gpurun_out/tl4_info.log 2>&1
  grep "^#" gpurun_out/tl4_info.log | head -2
  grep -i "coll channels\|error\|warn" gpurun_out/tl4_info.log | head -3 | cut -c1-200
done
rm -f gpurun_out/tl4_info.log
