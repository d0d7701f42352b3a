# 2-GPU A/B of the data-parallel gradient exchange (run under: gpurun --gpus 2 -- bash tools/run_2gpu_ab.sh)
timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -s -k "multimem and (bf16 or tiny)" > gpurun_out/pytest_multigpu_r02h.log 2>&1; tail -6 gpurun_out/pytest_multigpu_r02h.log
run() { name=$1; shift; env "$@" VB_BENCH_ALSO=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu-baseline --skip-roofline > gpurun_out/r02_bench_2gpu_h_$name.json 2> gpurun_out/r02_bench_2gpu_h_$name.err; echo $name $(grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/r02_bench_2gpu_h_$name.json | head -2 | tr '\n' ' ') $(grep -o "\"loss_first_last\": [^]]*]" gpurun_out/r02_bench_2gpu_h_$name.json); }
run mc64 VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=64
run mc32 VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=32
run mc128 VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=128
run mc64skip VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=64 VAULT_B200_MC_SKIP_KERNEL=1
run mc64repl VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=64 VAULT_B200_MC_SHARD_MASTER=0
run mc64fp32 VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=64 VB_COMM_DTYPE=fp32
run mc64dyn VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=64 VAULT_B200_DYNAMIC_TILES=1
