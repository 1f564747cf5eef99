#!/bin/bash
# bench line (default arguments) + the new GPU tests
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_splat_gpu.py -x -q -k "host or convert or golden" 2>&1 | tail -4
timeout 1200 python bench.py > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_now.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_now.json"))
ex = d.pop("extra", {})
print(json.dumps(d, indent=1)[:6000])
print(json.dumps(ex, indent=1)[:7000])
PY
