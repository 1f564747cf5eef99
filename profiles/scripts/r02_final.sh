#!/bin/bash
# round-2 FINAL profile artefacts: launch list of the bench command, full captures of the step kernels (scatter launch,
# epilogue launch; caches not flushed) and of the recipe's pass-B kernels
cd "$GRAFT_REPO_ROOT" || exit 1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-c5"
$CMD > gpurun_out/plain_bench_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_r02_final.csv $CMD > gpurun_out/ncu_launches_final.log 2>&1
echo "launch list rc=$?"
python profiles/scripts/run_fwd.py 8 soft 3 > gpurun_out/plain_fwd.log 2>&1 &&
ncu --set full --cache-control none --clock-control none --import-source on -k regex:"k_splat_(step|epilogue)" -s 20 -c 4 -f -o gpurun_out/step_r02_final python profiles/scripts/run_fwd.py 8 soft 3 > gpurun_out/ncu_step_final.log 2>&1
echo "step capture rc=$?"
python profiles/scripts/run_recipe.py 8 3 > gpurun_out/plain_recipe.log 2>&1 &&
ncu --set full --cache-control none --clock-control none --import-source on -k regex:"k_splat_(step|epilogue)" -s 48 -c 4 -f -o gpurun_out/recipe_r02_final python profiles/scripts/run_recipe.py 8 3 > gpurun_out/ncu_recipe_final.log 2>&1
echo "recipe capture rc=$?"
tail -2 gpurun_out/plain_recipe.log
