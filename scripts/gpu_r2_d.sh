#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/panel_check.py lin solve 2>&1 | tail -9
timeout 600 python scripts/panel_check.py time 2>&1 | grep panel
