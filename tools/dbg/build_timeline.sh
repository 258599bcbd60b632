#!/usr/bin/env bash
# debug variant of the library with per-warp timestamps (never shipped; loaded via AGX_LIB_PATH=tools/dbg/libagx_timeline.so).
# All translation units are built (tools/build_variant.py), so the variant exports the full ABI _lib.load() checks.
set -e
cd "$(dirname "$0")/../.."
python tools/build_variant.py timeline --hp1 "-DAGX_TIMELINE"
