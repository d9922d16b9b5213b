#!/bin/bash
for args in "96 54 64 36" "384 216 256 144" "200 100 120 90" "3840 2160 2560 1440" "960 540 640 360 rgba 2" "960 540 640 360 rgba 2 text"; do
  echo "== $args"
  timeout 60 python tools/dbg_resize.py $args 2>&1 | tail -3
done
