#!/bin/bash
# Round evidence on the GPU box, in order of importance (every step under its own timeout; the whole script fits a 6-minute call):
# parity tests of the kernels that changed + bit-identity with the committed result hash, the final bench line, the role timeline of
# the single-cursor schedule (variants/lib_single.so, built with -DSMR_SCHED_A=0), ncu launch list + full captures (tools/profile.sh),
# the DRAM traffic file bench.py quotes, compute-sanitizer memcheck over smoke().
R=${ROUND:-r2c}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py "tests/test_gpu_integration.py::test_reference_host_with_gpu_library" -x -q -m gpu > gpurun_out/${R}_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/${R}_parity.log
timeout 120 python tools/result_hash.py 200000 2> /dev/null | grep result_hash | tee gpurun_out/${R}_result_hash.txt
timeout 400 python bench.py --steps 20 --warmup 5 ${BENCH_EXTRA:---no-cpu-baseline} > gpurun_out/${R}_final_bench.json 2> gpurun_out/${R}_final_bench.err; echo "bench rc=$?"
if [ -f variants/lib_single.so ]; then
  SMR_TIMELINE=1 SMR_LIB_PATH=$PWD/variants/lib_single.so timeout 200 python bench.py --steps 6 --warmup 3 --reads 3000000 --no-cpu-baseline > gpurun_out/${R}_single_cursor_bench.json 2> gpurun_out/${R}_single_cursor_bench.err; echo "single-cursor bench rc=$?"
fi
SMR_TIMELINE=1 timeout 200 python bench.py --steps 6 --warmup 3 --reads 3000000 --no-cpu-baseline > gpurun_out/${R}_two_cursor_bench.json 2> gpurun_out/${R}_two_cursor_bench.err; echo "two-cursor bench rc=$?"
ROUND=$R timeout 400 bash tools/profile.sh; echo "profile rc=$?"
python tools/make_traffic.py $R 400000 > /dev/null; cp profiles/${R}_traffic.json gpurun_out/ 2>/dev/null; echo "traffic rc=$?"
timeout 150 compute-sanitizer --tool memcheck --log-file gpurun_out/${R}_sanitizer_memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_sanitizer_memcheck.out 2>&1; echo "memcheck rc=$?"
tail -2 gpurun_out/${R}_sanitizer_memcheck.log
