#!/bin/bash
# Build the native modules and run the test-suite (CPU tier always; GPU tier when a GPU is visible).
set -euo pipefail
cd "$(dirname "$0")/.."
python -m torchft_b200._build
python -m pytest tests -x -q -m "not gpu"
if python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)"; then
  python -m pytest tests -x -q -m gpu
fi
