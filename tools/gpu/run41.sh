timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
