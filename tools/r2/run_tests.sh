#!/bin/bash
# usage: bash tools/r2/run_tests.sh <tag> <pytest args...>   -> gpurun_out/<tag>/pytest.log
tag=$1; shift
mkdir -p gpurun_out/$tag
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/$tag/build.log 2>&1 || { echo build failed; tail -20 gpurun_out/$tag/build.log; exit 1; }
timeout -s KILL 1500 python -m pytest "$@" -q -m gpu -s --timeout 900 > gpurun_out/$tag/pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/$tag/pytest.log | tail -n 30
