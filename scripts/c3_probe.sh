# config-3 probe: one stream per batch against a single stream: bash scripts/c3_probe.sh
for m in 1 0; do
  LUX_C3_STREAMS=$m python - <<PY 2>/dev/null
import sys, json
sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
import bench_configs as B
r = B.config3()
print("streams=%s ms_per_step %.4f  match-turns/s %.3e" % ("$m", r["ms_per_step_all_sizes"], r["match_turns_per_s"]))
PY
done
