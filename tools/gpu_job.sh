python tools/ab_bench.py tools/ab/librpbmd_A.so reactive_pb_nn_md_b200/csrc/librpbmd.so c3 100 2 2>&1 | tail -3
python tools/diag_e2e.py c3 2>&1 | tail -1
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
