# 8-GPU A/B of the target workload (gpurun --gpus 8 -- bash tools/run_8gpu.sh [N])
N=${1:-8}
run() { name=$1; shift; env "$@" VB_BENCH_ALSO=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 30 --warmup 5 --skip-cpu-baseline --skip-roofline > gpurun_out/r02_ab_${N}gpu_$name.json 2> gpurun_out/r02_ab_${N}gpu_$name.err; echo $name $(grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/r02_ab_${N}gpu_$name.json | head -2 | tr '\n' ' ') $(grep -o "\"loss_first_last\": [^]]*]" gpurun_out/r02_ab_${N}gpu_$name.json) $(grep -o "\"replicas_identical\": [a-z]*" gpurun_out/r02_ab_${N}gpu_$name.json); }
run mc_bf16 VAULT_B200_COMM=multimem
run mc_fp32 VAULT_B200_COMM=multimem VB_COMM_DTYPE=fp32
run mc_bf16_b VAULT_B200_COMM=multimem
