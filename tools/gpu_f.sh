#!/bin/bash
out=gpurun_out/r2k; mkdir -p $out
(timeout 240 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log)
tail -12 $out/pytest.log
