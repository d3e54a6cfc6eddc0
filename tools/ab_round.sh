#!/bin/bash
# One GPU call: (1) parity + fuzz tests on the default build, (2) bit-identity of the variants' results on 200 k reads of the bench
# workload (tools/result_hash.py; the first variant is the yardstick), (3) short A/B benches (tools/ab_variants.sh).
# VARIANTS = names of variants/lib_*.so ("default" = the tree's build); HASH_FIRST = the ones hashed unconditionally.
mkdir -p gpurun_out
V=${VARIANTS:-"base default"}
HF=${HASH_FIRST:-"base default"}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu --durations=5 > gpurun_out/ab_parity_default.log 2>&1; echo "parity default rc=$?"; tail -3 gpurun_out/ab_parity_default.log
hash_of() {
  if [ $1 = default ]; then unset SMR_LIB_PATH; else export SMR_LIB_PATH=$PWD/variants/lib_$1.so; fi
  timeout 300 python tools/result_hash.py ${HASH_READS:-200000} 2> gpurun_out/ab_hash_$1.err | grep result_hash > gpurun_out/ab_hash_$1.txt
  echo "$1: $(cat gpurun_out/ab_hash_$1.txt)"; unset SMR_LIB_PATH
}
for v in $HF; do hash_of $v; done
first=$(echo $HF | cut -d' ' -f1); same=1
for v in $HF; do cmp -s gpurun_out/ab_hash_$first.txt gpurun_out/ab_hash_$v.txt || same=0; done
echo "hashes identical: $same"
if [ $same = 0 ]; then for v in $V; do case " $HF " in *" $v "*) ;; *) hash_of $v;; esac; done; fi
VARIANTS="$V" STEPS=${STEPS:-6} READS=${READS:-3000000} bash tools/ab_variants.sh 2>&1 | tee gpurun_out/ab_summary.txt
