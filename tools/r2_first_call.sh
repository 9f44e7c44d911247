#!/bin/bash
# First GPU call of round 2 (run as:  gpurun --timeout 2700 -- 'bash tools/r2_first_call.sh').
# Order: the regression gate first, then the experiments written blind at the end of round 1 (each under its own
# timeout: an untested tcgen05 pipeline can hang), then the bench lines with and without them.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log
# 1. does tcgen05.mma take A from tensor memory in the layouts linear_tc3_kernel assumes, and at what rate?
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I ai_education_generative_recommendation_b200/csrc \
     tools/tmem_a_probe.cu -o tools/tmem_a_probe > gpurun_out/r2_tmem_a_probe.txt 2>&1
timeout 60 tools/tmem_a_probe >> gpurun_out/r2_tmem_a_probe.txt 2>&1; echo "probe exit $?" >> gpurun_out/r2_tmem_a_probe.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I ai_education_generative_recommendation_b200/csrc \
     tools/tmem_a_probe2.cu -o tools/tmem_a_probe2 > gpurun_out/r2_tmem_a_probe2.txt 2>&1
timeout 60 tools/tmem_a_probe2 >> gpurun_out/r2_tmem_a_probe2.txt 2>&1; echo "probe2 exit $?" >> gpurun_out/r2_tmem_a_probe2.txt
# 2. linear_tc3_kernel: bit-identical to linear_tc2_kernel?  faster?
for cfg in c2_slice c5_slice; do
    timeout 240 python tools/check_tc3.py $cfg > gpurun_out/r2_check_tc3_$cfg.txt 2>&1; echo "check_tc3 exit $?" >> gpurun_out/r2_check_tc3_$cfg.txt
done
if grep -q "check_tc3 exit 0" gpurun_out/r2_check_tc3_c2_slice.txt; then
    timeout 300 python tools/ablate_tc3.py > gpurun_out/r2_ablation_linear_tc3.txt 2>&1
    # A/B: three K slabs of X in flight per producer thread instead of two
    bash tools/build_variant.sh p3 encode_tc3.cu -DT3_PREFETCH_N=3 > gpurun_out/r2_build_p3.log 2>&1 && \
        RQB200_LIB=$PWD/ai_education_generative_recommendation_b200/librqvae_b200_p3.so timeout 240 python tools/check_tc3.py c2_slice \
        > gpurun_out/r2_check_tc3_c2_slice_p3.txt 2>&1
    # A/B: eight epilogue warps (two per tensor-memory lane quarter) instead of four
    bash tools/build_variant.sh e8 encode_tc3.cu -DT3_EPI_WARPS=8 > gpurun_out/r2_build_e8.log 2>&1 && \
        RQB200_LIB=$PWD/ai_education_generative_recommendation_b200/librqvae_b200_e8.so timeout 240 python tools/check_tc3.py c2_slice \
        > gpurun_out/r2_check_tc3_c2_slice_e8.txt 2>&1
fi
# 2b. sort-free suffix dedup: identical ids / statistics?  faster?
timeout 300 python tools/check_dedup_list.py > gpurun_out/r2_check_dedup_list.txt 2>&1; echo "check_dedup_list exit $?" >> gpurun_out/r2_check_dedup_list.txt
# 2c. quantizer A/B: hi half of the residual tile from tensor memory, three row groups (quantize_tc.cu, -DQTC_A_HI_TMEM=1)
bash tools/build_variant.sh qts quantize_tc.cu -DQTC_A_HI_TMEM=1 -DQTC_GROUPS=3 > gpurun_out/r2_build_qts.log 2>&1 && {
    QLIB=$PWD/ai_education_generative_recommendation_b200/librqvae_b200_qts.so
    RQB200_LIB=$QLIB timeout 600 python -m pytest tests/test_gpu_tensorcore.py -x -q > gpurun_out/r2_qts_pytest.log 2>&1; echo "qts pytest exit $?" >> gpurun_out/r2_qts_pytest.log
    if grep -q "qts pytest exit 0" gpurun_out/r2_qts_pytest.log; then
        for cfg in c2 c5; do
            timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${cfg}_n1_base.json 2>/dev/null
            RQB200_LIB=$QLIB timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${cfg}_n1_qts.json 2>/dev/null
        done
    fi
}
# 3. bench lines: production kernels, then with linear_tc3_kernel in the step (only meaningful if step 2 said bit-identical)
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/r2_bench_c2_n1.err
RQB200_TC3=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c2_n1_tc3.json 2> gpurun_out/r2_bench_c2_n1_tc3.err
RQB200_DEDUP_LIST=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c2_n1_dedup_list.json 2> gpurun_out/r2_bench_c2_n1_dedup_list.err
# 4. one full ncu capture of the new first-layer kernel, only if it is correct (the run above exited 0 without ncu)
if grep -q "check_tc3 exit 0" gpurun_out/r2_check_tc3_c2_slice.txt; then
    RQB200_TC3=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tc3_kernel -c 1 \
        -o gpurun_out/r2_ncu_full_linear_tc3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_tc3.log 2>&1
fi
# 5. ncu evidence still missing from round 1: achieved DRAM throughput of the dedup kernels (sort path and list path)
timeout 600 ncu --set full --clock-control none -k "regex:radix_scatter_kernel|seg_rank_kernel|pack_keys_kernel" -c 5 \
    -o gpurun_out/r2_ncu_full_dedup_sort python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_dedup_sort.log 2>&1
if grep -q "check_dedup_list exit 0" gpurun_out/r2_check_dedup_list.txt; then
    RQB200_DEDUP_LIST=1 timeout 600 ncu --set full --clock-control none -k "regex:list_insert_kernel|list_rank_kernel" -c 2 \
        -o gpurun_out/r2_ncu_full_dedup_list python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_dedup_list.log 2>&1
fi
tail -n 3 gpurun_out/r2_pytest_gpu.log gpurun_out/r2_tmem_a_probe.txt gpurun_out/r2_tmem_a_probe2.txt gpurun_out/r2_check_tc3_c2_slice.txt gpurun_out/r2_check_dedup_list.txt
