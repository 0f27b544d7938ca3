#!/bin/bash
# Race / memory checks of the hand-written kernels (SURVEY.md section 5 "race detection"): compute-sanitizer over the small-shape
# GPU parity tests.  Run on the GPU box:   gpurun --timeout 1500 -- bash scripts/sanitize.sh
# Output: gpurun_out/sanitize_{memcheck,racecheck,synccheck}_*.log and gpurun_out/sanitize_summary.txt (copied to
# profiles/rNN_sanitize.md by hand).  racecheck only sees shared-memory hazards between threads (generic-proxy accesses); the
# async-proxy traffic of TMA / tcgen05 is ordered by the mbarrier protocols and is not modelled by the tool, so for the tcgen05
# kernels the signal is memcheck (out-of-bounds / misaligned global, shared and tensor-map accesses) + synccheck (barrier misuse).
set -u
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
SUM=gpurun_out/sanitize_summary.txt
: > $SUM
run() {   # tool, tag, pytest selection...
    local tool=$1 tag=$2; shift 2
    local log=gpurun_out/sanitize_${tool}_${tag}.log
    timeout ${SAN_TIMEOUT:-420} $CS --tool $tool --print-limit 20 --error-exitcode 0 \
        python -m pytest -x -q -m gpu "$@" > $log 2>&1
    local rc=$?
    local errs=$(grep -c "^========= .*\(Invalid\|Race\|hazard\|Error\|misaligned\|Barrier error\)" $log)
    local summ=$(grep "ERROR SUMMARY\|RACECHECK SUMMARY" $log | tail -1)
    local tests=$(grep -E "passed|failed" $log | tail -1)
    echo "$tool $tag: rc=$rc  reports=$errs  [$summ]  pytest: $tests" | tee -a $SUM
}
CONV="tests/test_conv_gpu.py::test_conv_matches_torch tests/test_conv_gpu.py::test_deconv8s4_matches_torch tests/test_conv_gpu.py::test_epilogue_variants tests/test_conv_gpu.py::test_f32_planar_output_and_class_bias tests/test_conv_gpu.py::test_deconv8s4_merged_subphases"
run memcheck conv $CONV
run memcheck kpred "tests/test_kpred_gpu.py::test_kpred_chains_vs_torch"
run memcheck metrics tests/test_metrics_gpu.py::test_golden_vectors_bit_exact tests/test_metrics_gpu.py::test_edge_cases tests/test_metrics_gpu.py::test_large_lists_and_sequential_replay_agree tests/test_metrics_gpu.py::test_degrade_matches_golden_and_oracle tests/test_metrics_gpu.py::test_psnr_ssim_kernel_vs_reference_golden_and_oracle
run memcheck losses tests/test_losses_gpu.py
run memcheck glue tests/test_glue_gpu.py::test_bilinear_fwd_bwd tests/test_glue_gpu.py::test_adaptive_avgpool_fwd_bwd tests/test_glue_gpu.py::test_patch_split_join_and_crop_flip_vs_oracle
run memcheck wgrad tests/test_train_gpu.py::test_conv_wgrad_vs_autograd tests/test_train_gpu.py::test_deconv_wgrad_vs_autograd tests/test_train_gpu.py::test_fused_adam_vs_torch_adam tests/test_train_gpu.py::test_batch_norm_fn_vs_torch
run racecheck metrics tests/test_metrics_gpu.py::test_golden_vectors_bit_exact tests/test_metrics_gpu.py::test_edge_cases tests/test_metrics_gpu.py::test_large_lists_and_sequential_replay_agree tests/test_metrics_gpu.py::test_degrade_matches_golden_and_oracle
run racecheck losses tests/test_losses_gpu.py
run racecheck conv "tests/test_conv_gpu.py::test_epilogue_variants" "tests/test_conv_gpu.py::test_deconv8s4_matches_torch" "tests/test_conv_gpu.py::test_deconv8s4_merged_subphases"
run synccheck conv "tests/test_conv_gpu.py::test_epilogue_variants" "tests/test_conv_gpu.py::test_deconv8s4_matches_torch" "tests/test_conv_gpu.py::test_deconv8s4_merged_subphases"
run synccheck kpred "tests/test_kpred_gpu.py::test_kpred_chains_vs_torch"
cat $SUM
