cd "$GRAFT_REPO_ROOT" || exit 1
echo default; python profiles/scripts/run_c4_nhwc.py | grep "channels_last  lists_nhwc=1"
for v in build_variants/nhwc_*.so; do echo $v; DCB_LIB_PATH=$PWD/$v python profiles/scripts/run_c4_nhwc.py | grep "channels_last  lists_nhwc=1"; done
