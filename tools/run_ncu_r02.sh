# Round-2 ncu evidence, one GPU (gpurun -- bash tools/run_ncu_r02.sh).  --set full captures of the kernels the bench names + in-graph wgrad A/B.
set -x
python tools/wgrad_colsum_time.py > gpurun_out/r02_wgrad_colsum_time.jsonl 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_bf16_kernel -s 3 -c 2 -f -o gpurun_out/r02_ncu_gemm_qkv_fwd python tools/one_gemm.py 11808 2304 768 0 > /dev/null 2>&1
$NCU -k regex:gemm_bf16_kernel -s 3 -c 2 -f -o gpurun_out/r02_ncu_gemm_mlp1_gelu python tools/one_gemm.py 11808 3072 768 1 > /dev/null 2>&1
$NCU -k regex:gemm_bf16_kernel -s 3 -c 2 -f -o gpurun_out/r02_ncu_gemm_mlp2_dgrad python tools/one_gemm.py 11808 768 3072 3 0 1 > /dev/null 2>&1
$NCU -k regex:gemm_bf16_kernel -s 3 -c 2 -f -o gpurun_out/r02_ncu_gemm_w1_wgrad_colsum python tools/one_gemm.py 3072 768 11808 5 1 1 2 256 1 > /dev/null 2>&1
$NCU -k regex:attn_sm100_kernel -s 3 -c 3 -f -o gpurun_out/r02_ncu_attn_sm100_s369 python tools/attn_ncu_case.py 32 369 12 > /dev/null 2>&1
$NCU -k regex:"ln_fwd_kernel|ln_bwd_kernel|adamw_kernel" -s 3 -c 3 -f -o gpurun_out/r02_ncu_hbm python tools/one_hbm.py > /dev/null 2>&1
ls -la gpurun_out/r02_ncu_*.ncu-rep
