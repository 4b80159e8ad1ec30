#!/usr/bin/env bash
# Re-creates the evidence under profiles/ (one B200; run each block through
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/profile_round.sh <block>' ).
# Numbers printed by anything running under ncu are never bench values.
set -u
R=${ROUND:-r02}
O=gpurun_out
case "${1:-help}" in
  bench)      # headline line with extras (fp16) -> profiles/${R}_bench_1gpu*.json
    python bench.py > $O/${R}_bench_1gpu.json 2> $O/${R}_bench_1gpu.err ;;
  dtypes)     # same step, both operand types, per-shape GEMM rates
    for d in fp16 bf16; do
      python bench.py --no-extras --no-cpu-baseline --dtype $d \
        --profile-dump $O/${R}_gemm_shapes_$d.json > $O/${R}_bench_1gpu_${d}_no_extras.json
    done ;;
  launches)   # every launch of one step with its device time (compare SHARES, not absolutes)
    ncu --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file $O/${R}_launches_step.csv \
        python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --graph 0
    python tools/launch_summary.py $O/${R}_launches_step.csv $O/${R}_launch_summary.txt ;;
  kernels)    # ncu --set full of one launch of each hot kernel at its north-star / VAE shape
    DWM_NCU_DTYPE=${DTYPE:-fp16} ncu --set full --clock-control none --import-source on \
        --profile-from-start off -f -o $O/${R}_ncu_kernels python tools/ncu_targets.py
    python tools/ncu_summary.py $O/${R}_ncu_kernels.ncu-rep > $O/${R}_ncu_kernels_summary.txt ;;
  micro)      # CUDA-event micro-benchmarks and the per-rank emulation of the 8-GPU step
    python tools/kernel_bench.py
    python tools/shard_emulate.py 8 5 fp16
    python tools/vae_bench.py 2; python tools/vae2d_bench.py 6; python tools/unet_bench.py 10 ;;
  parity)     # full-size parity files (also written by the GPU tests)
    python -m pytest tests/test_northstar_parity_gpu.py tests/test_fullsize_parity_gpu.py -m gpu -q ;;
  sass)       # no GPU needed
    python tools/sass_summary.py > profiles/${R}_sass_summary.txt ;;
  *)
    echo "usage: $0 bench|dtypes|launches|kernels|micro|parity|sass   (multi-GPU: tools/multi_gpu_check.sh N)" ;;
esac
