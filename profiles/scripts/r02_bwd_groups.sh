#!/bin/bash
# packed backward on 16 x 1080p: frames per group (packed cells L2-resident between the two passes) after the source pass rewrite
cd "$GRAFT_REPO_ROOT" || exit 1
for k in 0 1 2 3 4 8; do
  export DCB_BWD_GROUP_BYTES=$((k * 1080 * 1920 * 16))
  echo "frames per group $k (0 = all):"; python profiles/scripts/run_frames_bwd.py 16 | tail -1
done
