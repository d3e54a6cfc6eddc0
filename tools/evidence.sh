#!/bin/bash
# Round evidence on the GPU box: ncu launch list + captures (tools/profile.sh) and the DRAM traffic file bench.py quotes, compute-sanitizer
# logs (the alignment path through smoke(), the 8(f) kernels through tools/smoke_f.py), then the final bench lines.
R=${ROUND:-r2b}
mkdir -p gpurun_out
ROUND=$R timeout 1500 bash tools/profile.sh; echo profile rc=$?
python tools/make_traffic.py $R 400000 > /dev/null; cp profiles/${R}_traffic.json gpurun_out/; echo traffic rc=$?
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/${R}_sanitizer_memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_sanitizer_memcheck.out 2>&1; echo memcheck rc=$?
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/${R}_sanitizer_memcheck_f.log python tools/smoke_f.py > gpurun_out/${R}_sanitizer_memcheck_f.out 2>&1; echo memcheck_f rc=$?
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/${R}_sanitizer_racecheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_sanitizer_racecheck.out 2>&1; echo racecheck rc=$?
tail -3 gpurun_out/${R}_sanitizer_memcheck.log gpurun_out/${R}_sanitizer_memcheck_f.log gpurun_out/${R}_sanitizer_racecheck.log
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${R}_final_reference.json 2> gpurun_out/${R}_final_reference.err; echo reference rc=$?
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_final_bench.json 2> gpurun_out/${R}_final_bench.err; echo bench rc=$?
