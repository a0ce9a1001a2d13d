"""MODE-dispatching launcher, same contract as the reference's bin/launcher.py (MODE=synthesize|test)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    mode = os.getenv("MODE")
    if mode == "synthesize":
        from fastvocoder_b200.synthesizer import run_synthesizer
        run_synthesizer()
    elif mode == "test":
        from fastvocoder_b200.synthesizer import run_test
        run_test()
    elif mode == "publish":
        from fastvocoder_b200.synthesizer import run_publisher
        run_publisher()
    else:
        raise SystemExit(f"MODE={mode!r}: only the inference modes (synthesize, test, publish) exist in this repo; "
                         "train / preprocess are outside the generator forward path")
