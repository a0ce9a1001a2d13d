#!/usr/bin/env bash
# Real-time-factor loop (batch 1) over a folder of mel .npy files on the GPU.
#
#   bash test.sh <checkpoint> <folder of mel .npy> <model_name> <config.yaml>
set -euo pipefail
if [[ $# -ne 4 ]]; then
  sed -n '2,4p' "$0" >&2
  exit 2
fi
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
MODE=test exec python3 "$here/bin/launcher.py" --checkpoint_path "$1" --file_path "$2" --model_name "$3" --config "$4"
