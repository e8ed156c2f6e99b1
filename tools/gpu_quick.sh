#!/bin/bash
# usage: gpu_quick.sh "<pytest -k expr>"
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -s -k "$1" 2>&1 | tail -n 25
