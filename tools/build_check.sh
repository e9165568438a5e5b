#!/bin/bash
# build libwat.so in-tree and check that it loads and exports the whole ABI (no GPU needed)
set -e
cd "$(dirname "$0")/.."
make -C whisper-at_b200 -j8 2>&1 | grep -E "error|warning" || true
PYTHONDONTWRITEBYTECODE=1 python -c "
import sys; sys.path.insert(0,'whisper-at_b200')
from whisper_at import _lib
l=_lib.lib(); print('libwat ok, abi', l.wat_abi_version(), len(_lib.EXPORTS), 'exports')"
