#!/bin/bash
# Quick GPU visit: parity tests only.
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_tests.log) 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" || echo SMOKE_FAILED
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -40
