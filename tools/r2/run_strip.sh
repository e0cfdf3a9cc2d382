#!/bin/bash
tag=${1:-s1}
bash tools/r2/run_tests.sh $tag tests/test_gpu_strip.py tests/test_gpu_conv.py tests/test_gpu_bench_shapes.py -x
