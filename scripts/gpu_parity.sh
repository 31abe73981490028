set -x
nvidia-smi --query-gpu=name,memory.used,memory.total --format=csv
make -C oracle -s 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -60
nvidia-smi --query-gpu=memory.used --format=csv
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -5
