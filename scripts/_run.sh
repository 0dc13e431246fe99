timeout 600 python scripts/train_demo.py --steps 300 --batch 4096 2>&1 | tail -5
