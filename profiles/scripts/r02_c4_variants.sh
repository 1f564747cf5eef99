#!/bin/bash
# C4 (8 x 64 x 256 x 256 fp32 soft) forward per build variant of splat_lists.cu (build_variants/lists_*.so), plus the C4 parity tests on the default build
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests -x -q -m gpu -k "c4 or C4 or lists" 2>&1 | tail -3
echo "default:"; python profiles/scripts/run_c4.py | tail -2
for v in build_variants/lists_*.so; do
  echo "$v:"; DCB_LIB_PATH=$PWD/$v python profiles/scripts/run_c4.py | tail -2
done
