#!/bin/bash
# validation of the fan-out gradient slots + dependent launch: tests, then the bench with dependent launch off / on
tag=${1:-n}
bash tools/r2/run_tests.sh $tag tests/test_gpu_conv.py tests/test_gpu_pdl.py tests/test_gpu_model.py tests/test_gpu_train.py tests/test_gpu_fullwidth.py
BA="--no-lif --no-gpu-eager --no-cpu-baseline --no-cfg3"
SNN_DEPENDENT_LAUNCH=0 bash tools/r2/run_bench.sh ${tag}_pdl0 $BA 2>&1 | head -30
SNN_DEPENDENT_LAUNCH=1 bash tools/r2/run_bench.sh ${tag}_pdl1 $BA 2>&1 | head -30
