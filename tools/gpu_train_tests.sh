#!/bin/bash
# Run the training-step GPU tests in separate processes (a faulting kernel poisons the CUDA context of its process),
# one summary line per group into gpurun_out/$1.log
out=gpurun_out/${1:-train}.log
: > $out
run() {
  echo "=== $1" >> $out
  python -m pytest $2 -q --no-header -p no:cacheprovider -k "$3" 2>&1 | grep -E "^(FAILED|ERROR)|passed|failed|^E  |wgrad bench|rror:" | cut -c1-300 | head -40 >> $out
}
run "ops (no tensor cores)" tests/test_gpu_train_ops.py "not wgrad_tensor and not wgrad_bench and not batched_planes"
run "wgrad tc" tests/test_gpu_train_ops.py "wgrad_tensor"
run "wgrad bench" tests/test_gpu_train_ops.py "wgrad_bench"
run "batched planes" tests/test_gpu_train_ops.py "batched_planes"
run "e2e fp32" tests/test_gpu_train.py "matches_reference and fp32 and not tc_fwd"
run "e2e tc fwd fp32 bwd" tests/test_gpu_train.py "tc_fwd_fp32_bwd"
run "e2e tc" tests/test_gpu_train.py "matches_reference and tc and not fp32"
run "arena + dropout" tests/test_gpu_train.py "arena or dropout_training"
run "train graphs" tests/test_gpu_train.py "graph"
cat $out
