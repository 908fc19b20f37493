#!/bin/bash
# round 2, call A: attention A/B in one run + the whole GPU suite + smoke
mkdir -p gpurun_out
echo "=== times: new"; timeout -s KILL 100 python tools/attn_times.py 2>&1 | head -2
echo "=== times: pair1"; ECADK_ATTN_MODE=pair1 timeout -s KILL 100 python tools/attn_times.py 2>&1 | head -2
echo "=== times: new (again)"; timeout -s KILL 100 python tools/attn_times.py 2>&1 | head -2
echo "=== GPU suite"
timeout -s KILL 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -30
echo "=== smoke"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
