# A/B of kernel-variant builds (variants/lib_*.so, made by hand with -D flags): short bench + phase shares
for v in ${VARIANTS:-default}; do
  if [ $v = default ]; then unset SMR_LIB_PATH; else export SMR_LIB_PATH=$PWD/variants/lib_$v.so; fi
  timeout 400 python bench.py --steps ${STEPS:-10} --warmup 3 --reads ${READS:-5000000} --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; echo bench $v rc=$?
  python - <<PY
import json
j=json.load(open("gpurun_out/ab_$v.json"))
c=j["counters"]; s=c["dbg_sum_read_cycles"] or 1
print("$v", "value",round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms", {k: round(x,1) for k,x in j["kernel_ms_per_step"].items()}, "frac", round(j["roofline"]["frac"],3))
print("   planner shares", {k: round(c[k]/s,3) for k in ("cyc_vote","cyc_order","cyc_group","cyc_plan","cyc_wait","cyc_replay")}, "maxread_ms", round(c["dbg_max_read_cycles"]/j["steps"]/1.965e6,1), "spec", round(j["roofline"]["speculation_overhead"],4))
ss=sum(c[k] for k in ("sc_wait","sc_load","sc_sw","sc_pub"))
print("   scorer shares", {k: round(c[k]/max(ss,1),3) for k in ("sc_wait","sc_load","sc_sw","sc_pub")}, "cycles/pair in sw", round(c["sc_sw"]/max(1,c["spec_pairs"])), "pairs", c["spec_pairs"], "tasks", c["spec_calls"], "roundsA", c.get("rounds_a"), "roundsB", c.get("rounds_b"), "lis_calls", c["lis_calls"], "1-pair rounds", c.get("w1_cnt"), "avg wait cyc", round(c.get("w1_cyc",0)/max(1,c.get("w1_cnt",1))), "max read busy ms", round(c.get("dbg_max_read_busy_cycles",0)/j["steps"]/1.965e6,2))
PY
done
