#!/bin/bash
# host-buffer (e2e) call: PCIe ceilings of the box, then the call with 1..16 chunks (development build: MAGI_E2E_CHUNKS)
cd "$GRAFT_REPO_ROOT" || exit 1
python tools/e2e_probe.py 2>&1 | tail -5
for c in 2 4 8 16; do MAGI_LIB_NAME=libmagi_dev.so MAGI_E2E_CHUNKS=$c python tools/e2e_probe.py 2>&1 | tail -2; done
