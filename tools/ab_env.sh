#!/bin/bash
# A/B on ONE box: tools/ab_env.sh VAR "<command>"  runs the command with VAR unset, then set to 1, twice each
v=$1; shift
for i in 1 2; do
  echo "--- $v unset"; "$@"
  echo "--- $v=1"; env $v=1 "$@"
done
