#!/bin/bash
tag=$1; shift
bash tools/r2/run_tests.sh $tag tests/test_gpu_strip.py tests/test_gpu_conv.py tests/test_gpu_bench_shapes.py tests/test_gpu_model.py tests/test_gpu_train.py
bash tools/r2/run_ab.sh ${tag}_ab "$@"
