#!/bin/bash
# round 2, call X: the whole -m gpu suite under the bounds-asserting build (final sources)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ALACGPU_LIB=$PWD/alac/net_b200/libalacgpu_checked.so timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2x_checked_suite.log 2>&1; echo "checked suite rc=$?"; tail -15 gpurun_out/r2x_checked_suite.log
