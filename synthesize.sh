#!/bin/bash
# usage: bash synthesize.sh <checkpoint> <mel.npy> <out.wav> <model_name> <config.yaml>
# Same positional contract as the reference's synthesize.sh, but runs on the GPU (the reference pins CPU).
checkpoint=$1
mel_path=$2
wav_path=$3
model_name=$4
config=$5

export MODE=synthesize

python3 bin/launcher.py \
    --checkpoint_path "$checkpoint" \
    --mel_path "$mel_path" \
    --wav_path "$wav_path" \
    --model_name "$model_name" \
    --config "$config"
