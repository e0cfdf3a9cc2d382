#!/bin/bash
# whole -m gpu suite + smoke + default bench line
tag=${1:-full}
bash tools/r2/run_tests.sh $tag tests
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/$tag/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/$tag/smoke.log
shift
bash tools/r2/run_bench.sh ${tag} "$@"
