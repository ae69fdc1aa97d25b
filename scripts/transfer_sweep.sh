#!/bin/bash
# Sweep of the staged-transfer parameters (mcba_transfer.cu) on the GPU box: workers x chunk size.
for w in 4 8 12 16 24; do for kb in 1024 2048 4096; do
  echo "workers=$w chunk_kb=$kb: $(MCBA_XFER_WORKERS=$w MCBA_XFER_CHUNK_KB=$kb python scripts/transfer_timing.py | tail -2 | tr '\n' '|')"
done; done
