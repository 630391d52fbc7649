#!/bin/bash
# A/B timing on the GPU box: tools/ab.sh <variant>... ; "tree" = the in-tree library.
# Each variant: tools/time_cases.py (create_image wall + kernel times), best of 5.
for v in "$@"; do
  echo "== $v"
  if [ "$v" = tree ]; then python tools/time_cases.py 2>&1 | tail -3
  else RTB200_LIB=.variants/$v.so python tools/time_cases.py 2>&1 | tail -3; fi
done
