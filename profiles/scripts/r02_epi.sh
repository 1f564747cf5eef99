#!/bin/bash
# epilogue-only launches as CTAs of 1 / 2 / 4 / 8 warps: headline forward and the conditioning recipe
cd "$GRAFT_REPO_ROOT" || exit 1
for v in default build_variants/pipe_epi1.so build_variants/pipe_epi2.so build_variants/pipe_epi8.so; do
  if [ "$v" = default ]; then unset DCB_LIB_PATH; else export DCB_LIB_PATH=$PWD/$v; fi
  echo "== $v"
  python profiles/scripts/run_fwd.py 32 soft 6 | tail -3
  python profiles/scripts/run_recipe.py 16 6 | tail -4 | head -3
done
