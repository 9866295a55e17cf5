#!/usr/bin/env bash
# Build the sm_100a C-ABI extension in-tree (counterpart of the reference's twig/ops/make.sh).
set -euo pipefail
cd "$(dirname "$0")"
python build.py "$@"
