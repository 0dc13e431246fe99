timeout 600 python -m pytest tests/test_gpu_grad.py -x -q -k "regularisers" 2>&1 | tail -5
