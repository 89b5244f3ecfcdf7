#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_multi_device.py -m gpu -x -q 2>&1 | tail -40
