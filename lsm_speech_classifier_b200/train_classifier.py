"""The consumer of the path's output (/root/reference/train_classifier.py:7-53), kept only so the
pipeline can be checked end to end: it reads the same four arrays of lsm_features_larger.npz
(:27-31) and fits the same multinomial logistic regression (:36-41).  Two shims for the installed
scikit-learn 1.9 (the reference pins 1.4.2): `multi_class=` no longer exists (multinomial is the
default), and the report is restricted to the labels that occur (config 1 has 4 classes, not 12)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

CLASS_NAMES = ["yes", "no", "up", "visual", "backward", "stop", "bird", "cat", "nine", "eight", "zero", "follow"]


def train_and_evaluate_classifier(dataset_filename: str = "lsm_features_larger.npz", verbose: bool = True,
                                  readout: str = "sklearn"):
    """readout="sklearn": the reference's classifier; readout="device": the same multinomial objective fitted on the GPU
    (readout.LogisticRegression / lsm_logreg_fit; needs at least three classes)."""
    if readout == "device":
        from .readout import LogisticRegression
    else:
        from sklearn.linear_model import LogisticRegression
    from sklearn.metrics import accuracy_score, classification_report
    if not Path(dataset_filename).exists():
        print("Error: Dataset file not found. Please run 'extract_lsm_features.py' first.")
        return None
    data = np.load(dataset_filename)
    X_train, y_train = data['X_train_features'], data['y_train']
    X_test, y_test = data['X_test_features'], data['y_test']
    if verbose:
        print(f"Loaded {len(X_train)} training and {len(X_test)} test samples.")
        print("Training the Logistic Regression classifier...")
    clf = LogisticRegression(random_state=42, max_iter=1000)
    clf.fit(X_train, y_train)
    y_pred = clf.predict(X_test)
    accuracy = accuracy_score(y_test, y_pred)
    if verbose:
        labels = sorted(set(y_train.tolist()) | set(y_test.tolist()))
        names = [CLASS_NAMES[i] if i < len(CLASS_NAMES) else str(i) for i in labels]
        print("\n--- Final Results ---")
        print(f"Test Accuracy: {accuracy * 100:.2f}%\n")
        print("Classification Report:")
        print(classification_report(y_test, y_pred, labels=labels, target_names=names))
    return accuracy


if __name__ == "__main__":
    train_and_evaluate_classifier()
