timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "solve or damping or pipelining" 2>&1 | tail -2
echo "--- solve only: fused / unfused"
timeout 120 python scripts/solve_only.py
BALM_NO_FUSED_PANEL=1 timeout 120 python scripts/solve_only.py
timeout 120 python scripts/solve_only.py
BALM_NO_FUSED_PANEL=1 timeout 120 python scripts/solve_only.py
