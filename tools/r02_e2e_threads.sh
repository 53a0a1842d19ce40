#!/bin/bash
# e2e (host-buffer) throughput of the resident path against the host-side narrowing settings
for cfg in "default:" "narrow16:CVB_NARROW_UPLOAD=1 CVB_UPLOAD_THREADS=16" "narrow32:CVB_NARROW_UPLOAD=1 CVB_UPLOAD_THREADS=32" "narrow64:CVB_NARROW_UPLOAD=1 CVB_UPLOAD_THREADS=64" "narrow128:CVB_NARROW_UPLOAD=1 CVB_UPLOAD_THREADS=128"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-scale24 > gpurun_out/r02_e2e_$name.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_e2e_$name.json").read().strip().splitlines()[-1])
print("$name", "e2e %.3f G" % (d["e2e"]["value"]/1e9), "value %.2f G" % (d["value"]/1e9), "parity", d["parity"]["ok"])
PY
done
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "c2 or c1 or hub" 2>&1 | tail -2
echo done
