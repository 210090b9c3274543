#!/bin/bash
# Round-2 multi-GPU runs (one 8xB200 node): the complete loops in the BASELINE config #3 / #5 shapes with the gradient mean over the ranks
# fused into the Adam kernels (peer memory) and replayed from CUDA graphs, beside the NCCL variants and the single-GPU share.
#   gpurun --gpus 8 -- bash profiles/run_r02_multi_gpu.sh
set -u
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# every command under its own timeout: a hung collective must cost seconds, not the box (round 2 lost its GPU budget to one that was not)
run() { echo "== $*" >> $OUT/r2_multi.log; timeout 150 "$@" 2>>$OUT/r2_multi.err | grep -v "^episode\|^iter\|^update" >> $OUT/r2_multi.log; }
: > $OUT/r2_multi.log; : > $OUT/r2_multi.err
N=$(nvidia-smi -L | wc -l)
# config #3: TD3, 1,048,576 reactors over N GPUs, 8 fused-rollout steps + 8 DP updates of batch 4096 per rank and iteration
for dp in peer nccl-eager; do
  run $TR --nproc-per-node $N --master-port 29601 examples/td3_fused_rollout.py --n-envs 1048576 --iters 200 --steps-per-iter 8 --updates-per-iter 8 --batch 4096 --dp $dp
done
# the single-GPU share of the same job (131,072 reactors)
CUDA_VISIBLE_DEVICES=0 run python examples/td3_fused_rollout.py --n-envs $((1048576 / N)) --iters 200 --steps-per-iter 8 --updates-per-iter 8 --batch 4096
# SAC in the same shape
run $TR --nproc-per-node $N --master-port 29602 examples/sac_fused_rollout.py --n-envs 1048576 --iters 200 --steps-per-iter 8 --updates-per-iter 8 --batch 4096
CUDA_VISIBLE_DEVICES=0 run python examples/sac_fused_rollout.py --n-envs $((1048576 / N)) --iters 200 --steps-per-iter 8 --updates-per-iter 8 --batch 4096
# config #5: MADDPG / IDDPG, the two reactors as two agents, 262,144 env copies over N GPUs
run $TR --nproc-per-node $N --master-port 29603 examples/maddpg_two_agents.py --n-envs 262144 --iters 400 --batch 1024
run $TR --nproc-per-node $N --master-port 29604 examples/maddpg_two_agents.py --n-envs 262144 --iters 400 --batch 1024 --iddpg
CUDA_VISIBLE_DEVICES=0 run python examples/maddpg_two_agents.py --n-envs $((262144 / N)) --iters 400 --batch 1024
# the 2-rank NCCL / peer tests and the bench at N
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -3 >> $OUT/r2_multi.log
timeout 300 $TR --nproc-per-node $N --master-port 29605 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/r2_bench_n$N.json 2> $OUT/r2_bench_n$N.err
echo "bench rc=$?" >> $OUT/r2_multi.log
tail -40 $OUT/r2_multi.log
