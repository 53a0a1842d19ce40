#!/bin/bash
# resident drop-in (reference main.cpp on the outer boundary), upload seconds per level: narrowing (default for the
# reference's pageable vectors) vs the plain copy, and with the host's OpenMP threads sleeping instead of spinning
for cfg in "narrow_default:" "narrow_passive:OMP_WAIT_POLICY=passive" "plain:CVB_NARROW_UPLOAD=0" "narrow_default2:"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env CVB_ADAPTER_TIMING=1 $envs python tools/ref_cuda_baseline.py --scale 20 --only gc_gpu_resident > gpurun_out/r02_dropin_$name.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_dropin_$name.json"))
ph=d["gc_gpu_resident"].get("resident_phases") or []
print("$name", "upload_s per level:", [p["upload_s"] for p in ph][:3], "clustering_s:", [l[2] for l in d["gc_gpu_resident"]["levels"]][:3])
PY
done
