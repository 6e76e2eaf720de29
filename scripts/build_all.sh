#!/bin/bash
# builds the product library and the FB_PROFILE variant used by scripts/phase_profile.py
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd $ROOT
python -m freddi_b200.build --out build/prof_default.so -- -DFB_PROFILE "$@"
python -c "from freddi_b200 import build; build.build(force=True)"
echo built
