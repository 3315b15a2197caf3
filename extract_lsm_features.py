"""Entry script with the reference's name and flags (/root/reference/extract_lsm_features.py); the B200
implementation lives in lsm_speech_classifier_b200/extract_lsm_features.py."""
from lsm_speech_classifier_b200.extract_lsm_features import _cli

if __name__ == "__main__":
    _cli()
