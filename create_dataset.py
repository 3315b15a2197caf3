"""Entry script with the reference's name and flags (/root/reference/create_dataset.py); the B200
implementation lives in lsm_speech_classifier_b200/create_dataset.py."""
from lsm_speech_classifier_b200.create_dataset import _cli

if __name__ == "__main__":
    _cli()
