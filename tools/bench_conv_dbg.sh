#!/bin/bash
# Whole-step headroom of the headline bench with the conv kernel's operand fills removed (SR100_CONV_DBG, timing only:
# the images are wrong by construction).  Needs the development build of the library (make -C image-enhance-keras_b200/csrc
# DEV=1 into another directory, SR100_LIB=<that .so>, and bench.py --allow-dev-build); the default build ignores the variable.  Prints value (MP/s), ms/step, in-bench conv TFLOP/s, SM clock per setting.
for v in 0 1 3; do
  SR100_CONV_DBG=$v timeout 150 python bench.py --steps 5 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/bench_dbg_$v.json
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
d = json.load(open("/tmp/bench_dbg_%s.json" % v))
print(json.dumps({"SR100_CONV_DBG": int(v), "value": d["value"], "ms_per_step": d["ms_per_step"],
                  "conv_tflops": d["roofline"]["achieved"], "sm_mhz": d["clocks"]["sm_mhz"]}))
PY
done
