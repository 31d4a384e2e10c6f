cd $GRAFT_REPO_ROOT
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long_simple" 2>&1 | tail -15
