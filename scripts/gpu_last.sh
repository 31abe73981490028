#!/bin/bash
timeout 100 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "store_modes and (mixed or humanoid_random)" 2>&1 | tail -1
