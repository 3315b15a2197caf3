"""Entry script with the reference's name and flags (/root/reference/main.py); the B200
implementation lives in lsm_speech_classifier_b200/main.py."""
from lsm_speech_classifier_b200.main import _cli

if __name__ == "__main__":
    _cli()
