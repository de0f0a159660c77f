#!/bin/bash
# usage: tools/probe_variants.sh <probe args...>  -- runs tools/probe.py with every library under build_variants/
for lib in build_variants/*.so; do
  echo "== $lib"
  JELLYFYSH_B200_LIBRARY=$PWD/$lib timeout 200 python tools/probe.py "$@" 2>&1 | grep -E "step [2-5]|rror" | cut -c1-120
done
