#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_device_shim_gpu.py tests/test_abi_cpu.py -q -s -k "registry or multi_device" 2>&1 | grep -E "available_devices|passed|failed|Error|assert" | cut -c1-250 | tail -20
