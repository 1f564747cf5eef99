#!/bin/bash
# round-2 profile artefacts: launch list of the bench command, full captures of the step kernel (both launch kinds, caches
# not flushed) and of the live backward-warp kernel
cd "$GRAFT_REPO_ROOT" || exit 1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-c5"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_r02.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python profiles/scripts/run_fwd.py 8 soft 3 > gpurun_out/plain_fwd.log 2>&1 &&
ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_splat_step -s 20 -c 4 -f -o gpurun_out/step_r02 python profiles/scripts/run_fwd.py 8 soft 3 > gpurun_out/ncu_step.log 2>&1
echo "step capture rc=$?"
python profiles/scripts/run_backwarp.py > gpurun_out/plain_backwarp.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_backwarp -s 2 -c 2 -f -o gpurun_out/backwarp_r02 python profiles/scripts/run_backwarp.py > gpurun_out/ncu_backwarp.log 2>&1
echo "backwarp capture rc=$?"
cat gpurun_out/plain_backwarp.log | head -4
