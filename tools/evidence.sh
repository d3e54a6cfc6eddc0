#!/bin/bash
# Round evidence on the GPU box: final bench lines, ncu launch list + captures (tools/profile.sh), compute-sanitizer logs.
R=${ROUND:-r2}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_final_bench.json 2> gpurun_out/${R}_final_bench.err; echo bench rc=$?
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${R}_final_reference.json 2> gpurun_out/${R}_final_reference.err; echo reference rc=$?
ROUND=$R timeout 1500 bash tools/profile.sh; echo profile rc=$?
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/${R}_sanitizer_memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_sanitizer_memcheck.out 2>&1; echo memcheck rc=$?
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/${R}_sanitizer_racecheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_sanitizer_racecheck.out 2>&1; echo racecheck rc=$?
tail -3 gpurun_out/${R}_sanitizer_memcheck.log gpurun_out/${R}_sanitizer_racecheck.log
