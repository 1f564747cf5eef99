#!/bin/bash
# C4 forward + backward per build variant of splat_bwd.cu (registers per thread x channels in flight of k_bwd_source)
cd "$GRAFT_REPO_ROOT" || exit 1
echo "default:"; python profiles/scripts/run_c4.py bwd | tail -2
for v in build_variants/bwd_*.so; do
  echo "$v:"; DCB_LIB_PATH=$PWD/$v python profiles/scripts/run_c4.py bwd | tail -2
done
