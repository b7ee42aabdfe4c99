#!/bin/bash
# Build the library of git HEAD into profiles/ab/libhead.so (travels to the GPU box; select with HPFG_B200_LIB) so a working-tree
# change can be A/B-timed against the committed code inside ONE gpurun session (box-to-box variance is a few per cent).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
rm -rf /tmp/headbuild
git -C "$ROOT" worktree add -q /tmp/headbuild HEAD
make -s -j8 -C /tmp/headbuild/hpfg_b200/csrc > /dev/null
mkdir -p "$ROOT/profiles/ab"
cp /tmp/headbuild/hpfg_b200/libhpfg_b200.so "$ROOT/profiles/ab/libhead.so"
git -C "$ROOT" worktree remove --force /tmp/headbuild
echo "built $ROOT/profiles/ab/libhead.so"
