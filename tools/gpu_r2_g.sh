#!/bin/bash
mkdir -p gpurun_out
R56=$PWD/face-detection-and-tracking_b200/csrc/libfdt_b200_r56.so
FDT_B200_LIB=$R56 timeout 600 python -m pytest tests/test_detect_paths_gpu.py -m gpu -x -q > gpurun_out/r2g_pytest_r56.log 2>&1
tail -2 gpurun_out/r2g_pytest_r56.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --depth 4 --no-secondary --no-cpu-baseline > gpurun_out/r2g_$name.log 2> gpurun_out/r2g_$name.err
}
run fused64
run unfused64 FDT_DETECT_FUSED=0
run fused56 FDT_B200_LIB=$R56
run unfused56 FDT_B200_LIB=$R56 FDT_DETECT_FUSED=0
run unfused56_d3 FDT_B200_LIB=$R56 FDT_DETECT_FUSED=0 FDT_DETECT_DEPTH=3
for f in gpurun_out/r2g_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    u=r.get("unfused_path") or {}
    print("  value %.0f ms/step %.4f lat %.4f frac %.3f | unfused k3 b2b %s iso %s" % (d["value"], d["ms_per_step"], d["latency"]["ms_per_step"], r.get("frac",0), u.get("k_sort_nms_ms_back_to_back"), u.get("k_sort_nms_ms_isolated")))
except Exception as e:
    print("  parse error", e)
PY
done
