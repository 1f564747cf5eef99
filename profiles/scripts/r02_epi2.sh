#!/bin/bash
# epilogue launches: chunks per warp (DCB_NBATCH), pixels per lane in flight (DCB_NPER), warps per CTA -- headline on rough and smooth flow
cd "$GRAFT_REPO_ROOT" || exit 1
for v in default build_variants/pipe_*.so; do
  if [ "$v" = default ]; then unset DCB_LIB_PATH; else export DCB_LIB_PATH=$PWD/$v; fi
  echo "== $v"
  python profiles/scripts/run_fwd.py 32 soft 6 | tail -3 | head -2
done
