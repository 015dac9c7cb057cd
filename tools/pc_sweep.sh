#!/bin/bash
# GPU tuning aid: tools/pc_probe.py once per PC kernel variant (LPC_PC_VARIANT).
for v in 0 1 2 3 4; do echo "variant $v"; LPC_PC_VARIANT=$v python tools/pc_probe.py 5 2>&1 | sed 's/"has_changed.*"device_ms"/"device_ms"/'; done
