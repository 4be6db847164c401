#!/bin/bash
# round 2, GPU call 36 (1 GPU): the final committed tree: smoke + compress parity subset
cd /root/repo
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bit_exact or adversarial or sweep" 2>&1 | tail -2
