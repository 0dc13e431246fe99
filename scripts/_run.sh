timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -s 2>&1 | grep -E "full size|passed|failed|Error|assert" | tail -8
