python -m pytest tests/test_gpu_full_size.py -m gpu -q -x -k "small_box or oversized" 2>&1 | tail -15
