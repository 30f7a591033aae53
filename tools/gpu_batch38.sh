#!/bin/bash
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -v --timeout 90 --timeout-method=thread -p no:cacheprovider -k "parts or many_agents or large_map or config3_generated" > $O/r3k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r3k_pytest.log
