#!/bin/bash
# round 2, call 40: full GPU suite + smoke on the final commit
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/tests40.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke40.txt 2>&1
tail -2 $O/tests40.txt; tail -1 $O/smoke40.txt
