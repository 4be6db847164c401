#!/bin/bash
# round 2, GPU call 34 (1 GPU): the committed end-of-round tree once more: smoke + GPU suite
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -q -x > $O/r2zh_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2zh_pytest.log
