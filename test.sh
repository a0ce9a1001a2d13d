#!/bin/bash
# usage: bash test.sh <checkpoint> <folder of mel .npy> <model_name> <config.yaml>   (RTF loop, batch 1)
checkpoint=$1
filelist=$2
model_name=$3
config=$4

export MODE=test

python3 bin/launcher.py \
    --checkpoint_path "$checkpoint" \
    --model_name "$model_name" \
    --config "$config" \
    --file_path "$filelist"
