timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "other_decoders" 2>&1 | tail -8
