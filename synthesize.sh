#!/usr/bin/env bash
# mel (.npy) -> wav on the GPU through the B200-native generator path.
#
#   bash synthesize.sh <checkpoint> <mel.npy> <out.wav> <model_name> <config.yaml>
#
# Five positional arguments in the order FastVocoder users already pass them; model_name is one of
# melgan | hifigan | multiband-hifigan | basis-melgan.  Needs a CUDA device (there is no CPU path here).
set -euo pipefail
if [[ $# -ne 5 ]]; then
  sed -n '2,7p' "$0" >&2
  exit 2
fi
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
names=(checkpoint_path mel_path wav_path model_name config)
args=()
for i in "${!names[@]}"; do
  j=$((i + 1))
  args+=("--${names[$i]}" "${!j}")
done
MODE=synthesize exec python3 "$here/bin/launcher.py" "${args[@]}"
