#!/usr/bin/env bash
# First GPU call of the next round: the three pieces written after round 1's GPU budget was spent
# (default-off / opt-in code paths), then the regular suite.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/validate_pending.sh'
set -x
DWM_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_vae_gpu.py tests/test_model_gpu.py \
    -m gpu -q -k "encode or ring" 2>&1 | tail -15
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -5
