#!/bin/bash
echo "nproc $(nproc)"
(timeout 200 python -m pytest tests -m gpu -q -x -k "mux or shim or session" 2>&1 | tail -2)
for i in 1 2 3 4 5 6 7 8 9 10; do timeout 200 python tools/diag_mux.py --measure e2e --full 0 2>&1 | tail -1 | cut -c1-200; done
