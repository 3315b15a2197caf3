"""Entry script with the reference's name (/root/reference/train_classifier.py); see
lsm_speech_classifier_b200/train_classifier.py (the consumer of the path's output, scikit-learn on the host)."""
from lsm_speech_classifier_b200.train_classifier import train_and_evaluate_classifier

if __name__ == "__main__":
    train_and_evaluate_classifier()
