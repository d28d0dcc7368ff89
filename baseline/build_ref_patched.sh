#!/bin/bash
# Builds the reference WITH baseline/patches/*.patch applied (the multi-tile plugin seam) into
# baseline/_ref_patched/bin/.  The patch is applied to a scratch copy of /root/reference under baseline/_build/;
# /root/reference itself is never written.  The unmodified build (baseline/build_ref.sh -> baseline/_ref) stays the
# reference arm; this one is only the host that exercises the patched per-tile binding in tests.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC=/root/reference
[ -d "$SRC/src/lib/core" ] || { echo "no reference tree at $SRC"; exit 0; }
COPY="$HERE/_build/patched_src"
if [ ! -f "$COPY/.patched" ] || [ "$HERE/patches/0001-multi-tile-plugin-encode-decode.patch" -nt "$COPY/.patched" ]; then
  rm -rf "$COPY"; mkdir -p "$HERE/_build"
  cp -r "$SRC" "$COPY"; rm -rf "$COPY/.git"
  for p in "$HERE"/patches/*.patch; do patch -s -p1 -d "$COPY" < "$p"; done
  touch "$COPY/.patched"
fi
GROK_SRC="$COPY" GROK_BUILD_DIR="$HERE/_build/patched" GROK_OUT_DIR="$HERE/_ref_patched" bash "$HERE/build_ref.sh"
