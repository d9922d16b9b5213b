#!/bin/bash
for args in "3840 2160 2560 1440" "384 216 256 144" "256 144 384 216" "200 100 120 90" "128 72 192 108" "640 360 426 240" "1280 720 1920 1080" "1920 1080 1280 720" "3840 2160 2560 1440 rgba 4 text" "960 540 640 360 rgba 2 text"; do
  echo "== $args"
  timeout 120 python tools/dbg_resize.py $args 2>&1 | tail -4
done
