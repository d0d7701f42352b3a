# 2-GPU A/B of the data-parallel gradient exchange (run under: gpurun --gpus 2 -- bash tools/run_2gpu_ab.sh)
timeout 400 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -s -k "multimem" > gpurun_out/pytest_multigpu_r02i.log 2>&1; tail -8 gpurun_out/pytest_multigpu_r02i.log
run() { name=$1; shift; env "$@" VB_BENCH_ALSO=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu-baseline --skip-roofline > gpurun_out/r02_bench_2gpu_i_$name.json 2> gpurun_out/r02_bench_2gpu_i_$name.err; echo $name $(grep -o "\"ms_per_step\": [0-9.]*" gpurun_out/r02_bench_2gpu_i_$name.json | head -2 | tr '\n' ' ') $(grep -o "\"loss_first_last\": [^]]*]" gpurun_out/r02_bench_2gpu_i_$name.json); tail -c 300 gpurun_out/r02_bench_2gpu_i_$name.err | grep -i "error" ; }
run mc128 VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=128
run mc64 VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=64
run mc192 VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=192
run mc128seg VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=128 VAULT_B200_MC_IN_GRAPH=0
run mc128skip VAULT_B200_COMM=multimem VAULT_B200_MC_CTAS=128 VAULT_B200_MC_SKIP_KERNEL=1
