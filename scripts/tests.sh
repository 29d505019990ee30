#!/usr/bin/env bash
# CPU suite everywhere; GPU suite when a device is visible (multi-GPU tests pick up every visible device).
set -euo pipefail
cd "$(dirname "$0")/.."
python -m pytest tests -q -m "not gpu"
if python -c "import torch, sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
  python -c "import __graft_entry__ as g; g.build()"
  python -m pytest tests -q -m gpu
fi
