timeout 600 python -m pytest tests/test_gpu_train_psnr.py -x -q -s 2>&1 | grep -E "PSNR|passed|failed|Error|assert" | tail
