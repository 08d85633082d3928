#!/bin/bash
# Round-end regression on one B200: build check + smoke, then the whole GPU suite.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6
